"""cProfile of GraphConvModel._prepare_batch (the prefetch thread's per-batch Python) on prepared layouts, B = 4096."""
import cProfile, os, pstats, sys, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel, _DeviceSlot
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules
dev = torch.device("cuda", 0)
B = int(os.environ.get("B", 4096))
pool = [make_molecules(B, seed=i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
batches = []
for i, b in enumerate(m.default_generator(ds, epochs=8, deterministic=True)):
    batches.append(b)
    if len(batches) == 24:
        break
slots = [_DeviceSlot(dev) for _ in range(4)]
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for i in range(8):
        m._prepare_batch(batches[i], slots[i % 4])
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for i in range(8, 24):
        m._prepare_batch(batches[i], slots[i % 4])
    pr.disable()
    dt = (time.perf_counter() - t0) / 16
    torch.cuda.synchronize()
print("_prepare_batch: %.3f ms per batch (under cProfile)" % (dt * 1e3))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:48]))
