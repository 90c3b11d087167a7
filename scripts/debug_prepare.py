import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200 import ops
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.data import PackedDataset
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules
dev = torch.device("cuda", 0)
B = 4096
big = PackedMols.concat([make_molecules(B, seed=i) for i in range(4)]).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
gen = m.default_generator(ds, epochs=1000, deterministic=True)
batches = [next(gen) for _ in range(12)]
side = torch.cuda.Stream()
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
for rep in range(3):
    T.clear()
    with torch.cuda.stream(side):
        for inputs, labels, weights in batches:
            t0 = time.perf_counter(); topo = inputs.layout.to_device(dev); tick("to_device", t0)
            t0 = time.perf_counter(); feats = inputs.packed_features_pinned; pinned = feats.is_pinned(); tick("is_pinned", t0)
            t0 = time.perf_counter(); fd = feats.to(dev, non_blocking=True); tick("feats.to", t0)
            t0 = time.perf_counter(); x = ops.permute_rows(fd, topo.perm); tick("permute", t0)
            t0 = time.perf_counter(); di = topo.model_inputs(x, n_samples=int(inputs[3])); tick("model_inputs", t0)
            t0 = time.perf_counter()
            t = torch.from_numpy(np.ascontiguousarray(labels[0])).pin_memory().to(dev, non_blocking=True); tick("labels", t0)
    torch.cuda.synchronize()
    print("rep", rep, "pinned", pinned, {k: "%.3f" % (v / 12 * 1e3) for k, v in T.items()})
