"""Where does a D-MPNN training step go?  torch.profiler kernel table + wall/GPU time of one step (B=4096, QM9-shaped)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from deepchem_b200.dmpnn import DMPNNModel, GraphDataset
from deepchem_b200.dmpnn_data import make_graphs

dev = torch.device("cuda", 0)
B = 4096
pg = make_graphs(B, seed=0, shape="qm9")
y = np.random.default_rng(0).standard_normal((B, 12)).astype(np.float32)
torch.manual_seed(0)
m = DMPNNModel(device=dev, n_tasks=12, batch_size=B, gemm_mode="tf32x3")
ds = GraphDataset(pg, y)
batch = next(m.default_generator(ds, deterministic=True))
inputs, labels, weights = m._prepare_batch(batch)
m.model.train()


def step():
    m._train_step(inputs, labels, weights)


for _ in range(10):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("20 steps: host issue %.3f ms/step, wall %.3f ms/step" % (t_issue / 20 * 1e3, t_all / 20 * 1e3))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
