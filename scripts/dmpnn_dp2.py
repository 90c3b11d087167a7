"""Data-parallel D-MPNN through the fused engine on 2 GPUs (torchrun): each rank trains on its half of every global batch
with ONE all-reduce of the flat gradient slab per step; rank 0 also trains a single-GPU replica on the whole batch.  Equal
shards and a mean loss make the averaged gradient exact, so the parameters must agree after every step.  Also times the
data-parallel step (max over ranks, CUDA events).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/dmpnn_dp2.py
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from deepchem_b200 import parallel
from deepchem_b200.dmpnn import DMPNNModel, GraphDataset
from deepchem_b200.dmpnn_data import make_graphs

rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
B = 1024                                  # per rank
pg = make_graphs(world * B, seed=9, shape="qm9")
y = np.random.default_rng(9).standard_normal((world * B, 12)).astype(np.float32)
w = np.ones_like(y)
torch.manual_seed(0)
m = DMPNNModel(device=dev, n_tasks=12, batch_size=B, gemm_mode="tf32x3")
m.enable_data_parallel()
assert m._engine is not None
ref = None
if rank == 0:
    torch.manual_seed(0)
    ref = DMPNNModel(device=dev, n_tasks=12, batch_size=world * B, gemm_mode="tf32x3")
    ref.model.load_state_dict({k: v.clone() for k, v in m.model.state_dict().items()})
mine = GraphDataset(pg.slice(rank * B, (rank + 1) * B), y[rank * B:(rank + 1) * B], w[rank * B:(rank + 1) * B])
inputs, labels, weights = m._prepare_batch(next(m.default_generator(mine, deterministic=True)))
if rank == 0:
    rin, rl, rw = ref._prepare_batch(next(ref.default_generator(GraphDataset(pg, y, w), deterministic=True)))
worst = 0.0
for step in range(4):
    m._train_step(inputs, labels, weights)
    if rank == 0:
        ref._train_step(rin, rl, rw)
        sa, sb = m.model.state_dict(), ref.model.state_dict()
        for k in sa:
            d = float((sa[k] - sb[k]).abs().max() / max(float(sb[k].abs().max()), 1e-12))
            worst = max(worst, d)
torch.cuda.synchronize()
# timing: B = 4096 per rank
B2 = 4096
pg2 = make_graphs(B2, seed=20 + rank, shape="qm9")
y2 = np.random.default_rng(rank).standard_normal((B2, 12)).astype(np.float32)
m2 = DMPNNModel(device=dev, n_tasks=12, batch_size=B2, gemm_mode="tf32x3")
m2.enable_data_parallel()
i2, l2, w2 = m2._prepare_batch(next(m2.default_generator(GraphDataset(pg2, y2), deterministic=True)))
for _ in range(5):
    m2._train_step(i2, l2, w2)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    m2._train_step(i2, l2, w2)
e1.record()
dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"check": "2-rank data-parallel D-MPNN engine vs single-GPU replica on the whole batch, 4 Adam steps",
                      "max_rel_param_diff": worst, "ok": worst < 2e-5,
                      "dp_ms_per_step": float(t), "molecules_per_s": world * B2 / float(t) * 1e3, "n_gpus": world}), flush=True)
    assert worst < 2e-5, worst
dist.destroy_process_group()
