"""Where does GraphConvModel.predict spend its time?  (BASELINE config 5 shape: PCBA-like molecules, 128 tasks x 2 classes.)"""
import json, os, sys, time
os.environ["DCGC_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import PackedMols, make_molecules
dev = torch.device("cuda", 0)
B, NB = 4096, int(os.environ.get("NB", 64))
shards = [make_molecules(B, seed=100 + i, shape="pcba") for i in range(4)]
big = PackedMols.concat([shards[i % 4] for i in range(NB)]).pin_memory()
ds = PackedDataset(big)
torch.manual_seed(0)
m = GraphConvModel(128, [64, 64], 128, mode="classification", n_classes=2, batch_size=B, device=dev, gemm_mode="tf32x3")
m.predict(PackedDataset(big.slice(0, 4 * B)))
torch.cuda.synchronize()
for rep in range(3):
    m._pipe_trace.clear()
    t0 = time.perf_counter()
    p = m.predict(ds)
    dt = time.perf_counter() - t0
    tr = dict(m._pipe_trace)
    nb = max(1.0, tr.get("pf_batches", 1.0))
    print(json.dumps({"molecules_per_s": NB * B / dt, "ms_per_batch": dt / NB * 1e3, "out_mb": p.nbytes / 1e6,
                      "trace_ms_per_batch": {k: round(v / nb * 1e3, 3) for k, v in tr.items() if k != "pf_batches"}}), flush=True)
