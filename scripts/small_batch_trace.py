"""Where does a SMALL-batch training step go (the reference's own batch sizes: 100 on Delaney, 50 on Tox21)?
Host stage timers of the fit pipeline (DCGC_PIPE_TRACE) + the library's per-scope device times."""
import json, os, sys, time
os.environ["DCGC_PIPE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from deepchem_b200 import ops
from deepchem_b200.data import CSVLoader, PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import make_labels, make_molecules

ds1 = CSVLoader(["y"]).create_dataset(os.path.join(ROOT, "tests", "golden", "delaney.csv"))
pm = make_molecules(7831, seed=21, shape="tox21").pin_memory()
y, w = make_labels(7831, 12, "classification", seed=3, missing=0.25)
ds2 = PackedDataset(pm, y, w)
for name, ds, kw, batch in (("delaney b100", ds1, dict(n_tasks=1, mode="regression"), 100),
                            ("tox21 b50", ds2, dict(n_tasks=12, mode="classification", n_classes=2), 50),
                            ("tox21 b256", ds2, dict(n_tasks=12, mode="classification", n_classes=2), 256)):
    torch.manual_seed(0)
    n_tasks = kw.pop("n_tasks")
    m = GraphConvModel(n_tasks, [64, 64], 128, batch_size=batch, gemm_mode="tf32x3", **kw)
    m.fit(ds, nb_epoch=2, deterministic=True)
    torch.cuda.synchronize()
    m._pipe_trace.clear()
    t0 = time.perf_counter()
    m.fit(ds, nb_epoch=4, deterministic=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tr = dict(m._pipe_trace)
    steps = tr.get("fit_steps", 1.0)
    ops.profile_begin("*")
    m.fit(ds, nb_epoch=1, deterministic=True)
    torch.cuda.synchronize()
    rep = ops.profile_report()
    n_prof = -(-len(ds) // batch)
    dev_us = sum(v[0] for v in rep.values()) * 1e3 / n_prof
    print(json.dumps({"case": name, "ms_per_step_wall": dt / steps * 1e3, "molecules_per_s": len(ds) * 4 / dt,
                      "device_scopes_us_per_step": round(dev_us, 1),
                      "host_ms_per_step": {k: round(v / steps * 1e3, 3) for k, v in tr.items() if k not in ("fit_steps", "pf_batches")}}), flush=True)
