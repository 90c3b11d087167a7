"""Which stage bounds the end-to-end step?  Throughput of each stage of fit_generator ALONE on the GPU box:
(1) layout generator, (2) generator + prefetcher (H2D + device permute + records, no training), (3) training
steps on resident batches, (4) the full pipeline — for several worker counts."""
import itertools
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel, _Prefetcher
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules

dev = torch.device("cuda", 0)
B = 4096
pool = [make_molecules(B, seed=i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
m.model.train()
print("default host_workers", m.host_workers, "cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
N = 60
for workers in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 6, 8]:
    m.host_workers = workers
    gen = m.default_generator(ds, epochs=1000, deterministic=True, workers=workers)
    for _ in range(10):
        next(gen)
    t = time.perf_counter()
    for _ in range(N):
        next(gen)
    t_gen = (time.perf_counter() - t) / N * 1e3
    gen.close()
    gen = itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True, workers=workers), N + 10)
    k = 0
    for prepared in _Prefetcher(m, gen, 2):
        k += 1
        if k == 10:
            torch.cuda.synchronize()
            t = time.perf_counter()
    torch.cuda.synchronize()
    t_pre = (time.perf_counter() - t) / N * 1e3
    m.log_frequency = 1
    g2 = itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True, workers=workers), 10)
    m.fit_generator(g2, checkpoint_interval=0)
    g2 = itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True, workers=workers), N)
    torch.cuda.synchronize()
    t = time.perf_counter()
    m.fit_generator(g2, checkpoint_interval=0)
    torch.cuda.synchronize()
    t_full = (time.perf_counter() - t) / N * 1e3
    print("workers=%d: generator %.3f ms/batch | generator+prefetch %.3f | full fit %.3f" % (workers, t_gen, t_pre, t_full))
gen = m.default_generator(ds, epochs=1, deterministic=True, workers=1)
prepared = [m._prepare_batch(b) for b in itertools.islice(gen, 4)]
for rep in range(2):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(40):
        m._train_step(*prepared[i % 4])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("train only: host %.3f ms/step, device %.3f ms/step" % ((t1 - t) / 40 * 1e3, (time.perf_counter() - t) / 40 * 1e3))
