"""End-to-end fit rate with SHUFFLED epochs (the reference's default, deterministic=False) against contiguous batches:
GraphConv [128,128,128], B=4096, 8 batches per epoch from one pinned compact shard."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules
B = 4096
big = PackedMols.concat([make_molecules(B, seed=i, shape="zinc") for i in range(8)])
big.compact()
big.pin_memory()
y, w = make_labels(8 * B, 1, "regression", seed=0)
ds = PackedDataset(big, y, w)
for name, det, lazy in (("contiguous", True, "1"), ("shuffled, gathered by the layout workers (C, pinned, int8)", False, "1"),
                        ("shuffled, gathered by the iterator (numpy, pageable)", False, "0")):
    os.environ["DCGC_LAZY_TAKE"] = lazy
    torch.manual_seed(0)
    m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, gemm_mode="tf32x3")
    m.fit(ds, nb_epoch=2, deterministic=det)
    torch.cuda.synchronize()
    res = []
    for rep in range(3):
        t0 = time.perf_counter()
        m.fit(ds, nb_epoch=5, deterministic=det)
        torch.cuda.synchronize()
        res.append((time.perf_counter() - t0) / 40 * 1e3)
    print(json.dumps({"batches": name, "ms_per_step": [round(r, 3) for r in res], "molecules_per_s": B / min(res) * 1e3}), flush=True)
