"""Where a predict pass loses time when two processes share one host (run one copy per GPU at once)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200 import graphconvmodel as G
from deepchem_b200.data import PackedDataset
from deepchem_b200.synthetic import PackedMols, make_molecules
dev = torch.device("cuda", 0)
B, NB = 4096, int(os.environ.get("NB", 96))
shards = [make_molecules(B, seed=100 + i, shape="pcba") for i in range(4)]
big = PackedMols.concat([shards[i % 4] for i in range(NB)]).pin_memory()
ds = PackedDataset(big)
torch.manual_seed(0)
m = G.GraphConvModel(128, [64, 64], 128, mode="classification", n_classes=2, batch_size=B, device=dev, gemm_mode="tf32x3")
m.predict(PackedDataset(big.slice(0, 4 * B)))
torch.cuda.synchronize()
T = {"fwd_wall": 0.0, "add_wall": 0.0, "prep_wall": 0.0, "h2d_wall": 0.0, "n": 0}
gpu = []
ofwd = m._engine.forward
def fwd(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    r = ofwd(*a, **k)
    e1.record(); T["fwd_wall"] += time.perf_counter() - t0; T["n"] += 1
    gpu.append((e0, e1))
    return r
m._engine.forward = fwd
oadd = G._StagedResult.add
def add(self, vals):
    t0 = time.perf_counter(); r = oadd(self, vals); T["add_wall"] += time.perf_counter() - t0; return r
G._StagedResult.add = add
oprep = m._prepare_batch
def prep(batch, slot=None):
    t0 = time.perf_counter(); r = oprep(batch, slot); T["prep_wall"] += time.perf_counter() - t0; return r
m._prepare_batch = prep
och = G._chunked_h2d
def ch(dst, src):
    t0 = time.perf_counter(); r = och(dst, src); T["h2d_wall"] += time.perf_counter() - t0; return r
G._chunked_h2d = ch
t0 = time.perf_counter()
p = m.predict(ds)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
g = sum(a.elapsed_time(b) for a, b in gpu) / len(gpu)
print(json.dumps({"mol_per_s": NB * B / dt, "ms_per_batch": dt / NB * 1e3, "gpu_fwd_ms": g,
                  "main_fwd_wall_ms": T["fwd_wall"] / T["n"] * 1e3, "main_add_wall_ms": T["add_wall"] / T["n"] * 1e3,
                  "prefetch_prepare_ms": T["prep_wall"] / T["n"] * 1e3, "of_which_chunked_h2d_issue_ms": T["h2d_wall"] / T["n"] * 1e3}))
