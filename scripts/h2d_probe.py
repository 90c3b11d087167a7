"""H2D bandwidth from pinned memory and the cost of the pieces of _prepare_batch on the GPU box."""
import faulthandler
import itertools
import os
import sys
import time

faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel, _DeviceSlot
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules

dev = torch.device("cuda", 0)
for mb in (4, 32, 128):
    h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print("H2D pinned %d MiB: %.1f GB/s" % (mb, 10 * (mb << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9))
B = 4096
pool = [make_molecules(B, seed=i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
m.model.train()
gen = m.default_generator(ds, epochs=1000, deterministic=True, workers=2)
batches = [next(gen) for _ in range(8)]
slots = [_DeviceSlot(dev) for _ in range(4)]
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    for i in range(8):
        m._prepare_batch(batches[i], slots[i % 4])
    torch.cuda.synchronize()
    for rep in range(3):
        t = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for i in range(8):
            m._prepare_batch(batches[i], slots[i % 4])
        e1.record(side)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print("_prepare_batch x8: host %.3f ms/batch, device %.3f ms/batch, wall %.3f ms/batch" % (
            (t1 - t) / 8 * 1e3, e0.elapsed_time(e1) / 8, (time.perf_counter() - t) / 8 * 1e3))
# generator alone, by piece
t = time.perf_counter()
n = 0
for X_b, y_b, w_b, ids in ds.iterbatches(batch_size=B, epochs=10, deterministic=True, pad_batches=True):
    n += 1
print("iterbatches: %.3f ms/batch" % ((time.perf_counter() - t) / n * 1e3))
X_b = next(iter(ds.iterbatches(batch_size=B, epochs=1, deterministic=True, pad_batches=True)))[0]
t = time.perf_counter()
for _ in range(20):
    m.batch_inputs(X_b)
print("batch_inputs (1 thread): %.3f ms/batch" % ((time.perf_counter() - t) / 20 * 1e3))
for workers in (6, 8):
    gen = m.default_generator(ds, epochs=1000, deterministic=True, workers=workers)
    for _ in range(10):
        next(gen)
    t = time.perf_counter()
    for _ in range(40):
        next(gen)
    print("generator workers=%d: %.3f ms/batch" % (workers, (time.perf_counter() - t) / 40 * 1e3))
    gen.close()
