#!/usr/bin/env python
"""Two (or more) ranks under torchrun: the fused engine with synchronised BatchNorm (statistics exchanged through
peer-memory mailboxes inside the finalize kernels, dcgc_gcmodel_train_step_sync) against ONE process on the
concatenated batch — loss, every gradient after the averaging all-reduce, the running statistics — and its step time
next to the unsynchronised data-parallel step.
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/syncbn_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from deepchem_b200 import parallel
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules

rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
B = int(os.environ.get("B", 512))
LAYERS = [128, 128]
out = {"world": world, "B_per_rank": B}

# ---- the same global batch on every rank (seeded), each rank trains on its slice
parts = [make_molecules(B, seed=40 + r) for r in range(world)]
labels = [make_labels(B, 1, "regression", seed=40 + r) for r in range(world)]
torch.manual_seed(0)
m = GraphConvModel(1, LAYERS, 128, mode="regression", batch_size=B, device=dev, sync_batch_norm=True)
assert m._engine is not None, "the engine must take sync_batch_norm models on CUDA + NCCL"
m.enable_data_parallel()
m.model.train()
y, w = labels[rank]
batch = m._prepare_batch((m.batch_inputs(parts[rank].pin_memory()), [y], [w]))
eng = m._engine
loss = eng.train_step(batch[0][1]._dcgc_topology, batch[0][0], batch[1][0].contiguous(), batch[2][0].contiguous(), B)
g = eng.grads.clone()
dist.all_reduce(g, op=dist.ReduceOp.SUM)
g /= world
lt = loss.detach().clone().double().reshape(1)
dist.all_reduce(lt, op=dist.ReduceOp.SUM)
lt /= world
bn = eng.bn_running.clone()
# bit-identical statistics on every rank
ref_bn = bn.clone()
dist.broadcast(ref_bn, src=0)
same_stats = torch.tensor([1.0 if torch.equal(ref_bn, bn) else 0.0], device=dev)
dist.all_reduce(same_stats, op=dist.ReduceOp.MIN)
out["running_stats_bit_identical_across_ranks"] = bool(same_stats.item() == 1.0)

if rank == 0:
    # ---- one process, the concatenated batch, plain BatchNorm
    torch.manual_seed(0)
    s = GraphConvModel(1, LAYERS, 128, mode="regression", batch_size=B * world, device=dev)
    s.model.load_state_dict({k: v for k, v in m.model.state_dict().items() if "running" not in k and "num_batches" not in k},
                            strict=False)
    # (the sync model's running statistics moved in its step: start the reference from fresh ones, as m did)
    for b_ in s.model.batch_norms:
        b_.reset_running_stats()
    # parameters: m has not taken an optimizer step, so they are still the seeded initial values
    s._engine.adopt()
    s.model.train()
    ya = np.concatenate([l[0] for l in labels]); wa = np.concatenate([l[1] for l in labels])
    sb = s._prepare_batch((s.batch_inputs(PackedMols.concat(parts).pin_memory()), [ya], [wa]))
    se = s._engine
    ls = se.train_step(sb[0][1]._dcgc_topology, sb[0][0], sb[1][0].contiguous(), sb[2][0].contiguous(), B * world)
    torch.cuda.synchronize()
    out["loss_sync_mean"] = float(lt)
    out["loss_single_process"] = float(ls)
    worst = ("", 0.0)
    for (name, p), (_, q) in zip(m.model.named_parameters(), s.model.named_parameters()):
        off = p.grad.storage_offset() - eng.grads.storage_offset()
        gs = torch.as_strided(g, p.grad.shape, p.grad.stride(), off)
        ref = q.grad
        scale = float(ref.abs().max())
        if scale == 0.0:
            continue
        e = float((gs - ref).abs().max()) / scale
        if e > worst[1]:
            worst = (name, e)
    out["worst_gradient_rel_err"] = {"tensor": worst[0], "err": worst[1]}
    out["running_stats_rel_err"] = float((bn - se.bn_running).abs().max() / se.bn_running.abs().max())
dist.barrier()

# ---- step time: synchronised against plain data-parallel (B = 4096 per rank, the bench shape)
def timed(sync):
    torch.manual_seed(0)
    mm = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=4096, device=dev, sync_batch_norm=sync)
    mm.enable_data_parallel()
    mm.model.train()
    pool = []
    for i in range(3):
        pm = make_molecules(4096, seed=100 * rank + i).pin_memory()
        yy, ww = make_labels(4096, 1, "regression", seed=100 * rank + i)
        pool.append(mm._prepare_batch((mm.batch_inputs(pm), [yy], [ww])))
    for i in range(5):
        mm._train_step(*pool[i % 3]); mm._global_step += 1
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 40
    for i in range(K):
        mm._train_step(*pool[i % 3]); mm._global_step += 1
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

out["ms_per_step_plain_dp"] = timed(False)
out["ms_per_step_sync_bn"] = timed(True)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
