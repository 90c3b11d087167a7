"""Debug helper: time breakdown of the end-to-end fit path (host build, H2D, permute, step)."""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200.synthetic import make_molecules, make_labels, PackedMols
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200 import ops

B = 4096
pool = [make_molecules(B, seed=i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression")
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B)
m.model.train()
sync = torch.cuda.synchronize
gen = m.default_generator(ds, epochs=100, deterministic=True)
T = {"gen": 0, "to_device": 0, "feat_h2d": 0, "permute": 0, "labels": 0, "step": 0}
for it in range(24):
    t0 = time.perf_counter(); batch = next(gen); t1 = time.perf_counter()
    inputs, labels, weights = batch
    topo = inputs.layout.to_device(m.device); sync(); t2 = time.perf_counter()
    ft = torch.from_numpy(np.ascontiguousarray(inputs.packed_features, dtype=np.float32))
    pinned = ft.is_pinned()
    feats = ft.to(m.device, non_blocking=True); sync(); t3 = time.perf_counter()
    x = ops.permute_rows(feats, topo.perm); sync(); t4 = time.perf_counter()
    x._dcgc_zero_padded = True
    dev_inputs = topo.model_inputs(x, n_samples=int(inputs[3]))
    yl = [torch.as_tensor(np.asarray(a), device=m.device) for a in labels]
    wl = [torch.as_tensor(np.asarray(a), device=m.device) for a in weights]; sync(); t5 = time.perf_counter()
    loss = m._train_step(dev_inputs, yl, wl); float(loss); t6 = time.perf_counter()
    if it >= 4:
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
            T[k] += v
print("pinned features:", pinned)
print({k: "%.3f ms" % (v / 20 * 1e3) for k, v in T.items()})
# threaded pipeline
for pf in (0, 2, 2):
    losses = []
    m.log_frequency = 1
    m.fit_generator(itertools.islice(m.default_generator(ds, epochs=100, deterministic=True), 5), checkpoint_interval=0, prefetch=pf)
    sync(); t0 = time.perf_counter()
    m.fit_generator(itertools.islice(m.default_generator(ds, epochs=100, deterministic=True), 40), checkpoint_interval=0, prefetch=pf, all_losses=losses)
    sync(); print("prefetch=%d: %.3f ms/step" % (pf, (time.perf_counter() - t0) / 40 * 1e3))
m.log_frequency = 1000
sync(); t0 = time.perf_counter()
m.fit_generator(itertools.islice(m.default_generator(ds, epochs=100, deterministic=True), 40), checkpoint_interval=0, prefetch=2)
sync(); print("prefetch=2, no per-step readback: %.3f ms/step" % ((time.perf_counter() - t0) / 40 * 1e3))
