"""How stable is the D-MPNN end-to-end rate?  Three timed fit_generator passes per GEMM mode in one process, per-pass ms/step."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepchem_b200.dmpnn import DMPNNModel, GraphDataset
from deepchem_b200.dmpnn_data import make_graphs
dev = torch.device("cuda", 0)
B = 4096
big = make_graphs(4 * B, seed=1, shape="qm9").pin_memory()
y4 = np.random.default_rng(1).standard_normal((4 * B, 12)).astype(np.float32)
ds4 = GraphDataset(big, y4)
for mode in ("tf32x3", "bf16", "tf32x3"):
    torch.manual_seed(0)
    m = DMPNNModel(device=dev, n_tasks=12, batch_size=B, gemm_mode=mode)
    m.fit_generator(m.default_generator(ds4, epochs=3, deterministic=True))
    torch.cuda.synchronize()
    res = []
    for rep in range(4):
        t0 = time.perf_counter()
        m.fit_generator(m.default_generator(ds4, epochs=10, deterministic=True))
        torch.cuda.synchronize()
        res.append((time.perf_counter() - t0) / 40 * 1e3)
    print(json.dumps({"mode": mode, "engine": m._engine is not None, "host_workers": m.host_workers, "ms_per_step": [round(r, 3) for r in res]}), flush=True)
