"""Where does the end-to-end step go?  Times the three stages of fit_generator separately on the GPU box."""
import sys, os, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.data import PackedDataset
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules
dev = torch.device("cuda", 0)
B = 4096
pool = [make_molecules(B, seed=i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
m.model.train()
print("host_workers", m.host_workers, "cpus", os.cpu_count())
for workers in (1, 2, 4, 8):
    gen = m.default_generator(ds, epochs=1000, deterministic=True, workers=workers)
    for _ in range(8): next(gen)
    t = time.perf_counter(); n = 40
    for _ in range(n): next(gen)
    print("generator only, workers=%d: %.3f ms/batch" % (workers, (time.perf_counter() - t) / n * 1e3))
    gen.close()
gen = m.default_generator(ds, epochs=1000, deterministic=True)
batches = [next(gen) for _ in range(12)]
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    prepared = [m._prepare_batch(b) for b in batches]
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("_prepare_batch: host %.3f ms/batch, with device sync %.3f ms/batch" % ((t1 - t) / 12 * 1e3, (t2 - t) / 12 * 1e3))
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    for p in prepared: loss = m._train_step(*p)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("_train_step: host %.3f ms/step, with device sync %.3f ms/step" % ((t1 - t) / 12 * 1e3, (t2 - t) / 12 * 1e3))
for lf in (1, 100):
    m.log_frequency = lf
    for rep in range(2):
        g2 = itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True), 60)
        torch.cuda.synchronize(); t = time.perf_counter()
        m.fit_generator(g2, checkpoint_interval=0)
        torch.cuda.synchronize()
        print("fit_generator log_frequency=%d: %.3f ms/step" % (lf, (time.perf_counter() - t) / 60 * 1e3))
