"""Per-stage host timers of the end-to-end fit pipeline (DCGC_PIPE_TRACE=1), same setup as bench.py's e2e leg."""
import faulthandler
import itertools
import os
import sys
import time

faulthandler.enable()
os.environ["DCGC_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules

dev = torch.device("cuda", 0)
B = 4096
pool = [make_molecules(B, seed=i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
for workers in [int(a) for a in sys.argv[1:]] or [4]:
    os.environ["DCGC_HOST_WORKERS"] = str(workers)
    m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
    m.model.train()
    m.log_frequency = 1
    m.fit_generator(itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True), 40), checkpoint_interval=0)
    torch.cuda.synchronize()
    m._pipe_trace.clear()
    K = 100
    t = time.perf_counter()
    m.fit_generator(itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True), K), checkpoint_interval=0)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) / K * 1e3
    tr = m._pipe_trace
    print("workers=%d: %.3f ms/step | fit thread: wait for batch %.3f, _train_step host %.3f | prefetch thread: wait "
          "generator %.3f, wait slot %.3f, _prepare_batch %.3f, wait queue %.3f  (ms per step)" % (
              workers, ms, tr["fit_wait_batch"] / K * 1e3, tr["fit_train_step_host"] / K * 1e3,
              tr["pf_wait_generator"] / K * 1e3, tr["pf_wait_slot"] / K * 1e3, tr["pf_prepare"] / K * 1e3,
              tr["pf_wait_queue"] / K * 1e3))
    del m
# per-scope device time INSIDE the end-to-end run (side-stream H2D / permute running concurrently) against
# the same scopes on resident batches: slower kernels (contention) or bubbles?
from deepchem_b200 import ops
os.environ["DCGC_HOST_WORKERS"] = "4"
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
m.model.train()
m.log_frequency = 1
m.fit_generator(itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True), 30), checkpoint_interval=0)
torch.cuda.synchronize()
K = 40
ops.profile_begin("*")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
m.fit_generator(itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True), K), checkpoint_interval=0)
e1.record()
torch.cuda.synchronize()
rep_e2e = ops.profile_report()
ms_e2e = e0.elapsed_time(e1) / K
gen = m.default_generator(ds, epochs=1, deterministic=True, workers=1)
prepared = [m._prepare_batch(b) for b in itertools.islice(gen, 4)]
for i in range(8):
    m._train_step(*prepared[i % 4])
torch.cuda.synchronize()
ops.profile_begin("*")
e0.record()
for i in range(K):
    m._train_step(*prepared[i % 4])
e1.record()
torch.cuda.synchronize()
rep_res = ops.profile_report()
ms_res = e0.elapsed_time(e1) / K
print("main-stream events: e2e %.3f ms/step, resident %.3f ms/step (both with scope events on)" % (ms_e2e, ms_res))
print("| scope | e2e us/step | resident us/step |")
tot = [0.0, 0.0]
for name in sorted(set(rep_e2e) | set(rep_res), key=lambda n: -rep_e2e.get(n, (0, 0))[0]):
    a, b = rep_e2e.get(name, (0.0, 0))[0] * 1e3 / K, rep_res.get(name, (0.0, 0))[0] * 1e3 / K
    tot[0] += a
    tot[1] += b
    print("| %s | %.1f | %.1f |" % (name, a, b))
print("| total | %.1f | %.1f |" % tuple(tot))
