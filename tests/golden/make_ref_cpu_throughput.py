"""SURVEY 8(d) CPU line (1): the UNMODIFIED reference GraphConvModel (torch port, [64,64], dense 128, CPU) timed in the
build container on the synthetic ZINC-shaped stream at B in {64 .. 1024}; B = 4096 is not attempted (the reference's
unsorted_segment_max materialises [B, N, F] and is OOM-killed there).  Its GraphConv output is detached
(SURVEY 0.3), so `fit` trains only the dense / head layers: the number is an upper bound of what the reference does per
molecule, on fewer FLOPs than the benchmark configuration.  Writes tests/golden/ref_cpu_throughput.json (the reference
does not exist on the GPU box; bench.py quotes this file).

    python tests/golden/make_ref_cpu_throughput.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _refimport import import_reference  # noqa: E402

dc = import_reference()
import torch  # noqa: E402
from deepchem.feat.mol_graphs import ConvMol  # noqa: E402
from deepchem.models.torch_models import GraphConvModel  # noqa: E402
from deepchem_b200.synthetic import make_labels, make_molecules  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
rows = []
for B in (64, 128, 256, 512, 1024):
    pm = make_molecules(B, seed=0, shape="zinc")
    y, w = make_labels(B, 1, "regression", seed=0)
    mols = np.empty(B, dtype=object)
    for i, (f, adj) in enumerate(pm.to_list()):
        mols[i] = ConvMol(f.astype(np.float64), adj)
    ds = dc.data.NumpyDataset(mols, y, w)
    m = GraphConvModel(1, [75, 64], mode="regression", batch_size=B, device=torch.device("cpu"))
    steps = 3 if B >= 512 else 6
    m.fit(ds, nb_epoch=1)                                     # warm-up (builds the model)
    t0 = time.perf_counter()
    m.fit(ds, nb_epoch=steps)
    dt = (time.perf_counter() - t0) / steps
    rows.append({"batch": B, "ms_per_step": dt * 1e3, "molecules_per_s": B / dt})
    print(rows[-1], flush=True)
best = max(rows, key=lambda r: r["molecules_per_s"])
out = {"what": "unmodified reference deepchem.models.torch_models.GraphConvModel(1, [75, 64], mode='regression') [64,64]+dense128 "
               "fit() on CPU: agglomerate_mols + forward + loss + backward (dense/head only: GraphConv is detached) + Adam",
       "where": "build container (no GPU), torch %s, %d threads" % (torch.__version__, torch.get_num_threads()),
       "cores": torch.get_num_threads(), "rows": rows, "best": best,
       "b4096": "not run: unsorted_segment_max is O(B*N*F) memory, OOM-killed at B=4096 (SURVEY 0.5)"}
json.dump(out, open(os.path.join(HERE, "ref_cpu_throughput.json"), "w"), indent=1)
print("best", best)
