#!/usr/bin/env python
"""Two ranks (torchrun): the overlapped per-slice gradient all-reduce must give bit-identical parameters to the single
all-reduce after the step, and the step time of both.  torchrun --nproc-per-node 2 scripts/dp_overlap_check.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from deepchem_b200 import parallel
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import make_labels, make_molecules

rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
B = 4096
res = {}
for mode in ("1", "0"):
    os.environ["DCGC_NO_OVERLAP"] = mode
    torch.manual_seed(0)
    m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev)
    m.enable_data_parallel()
    m.model.train()
    pool = []
    for i in range(3):
        pm = make_molecules(B, seed=100 * rank + i).pin_memory()
        y, w = make_labels(B, 1, "regression", seed=100 * rank + i)
        pool.append(m._prepare_batch((m.batch_inputs(pm), [y], [w])))
    for i in range(5):
        m._train_step(*pool[i % 3]); m._global_step += 1
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 40
    for i in range(K):
        m._train_step(*pool[i % 3]); m._global_step += 1
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = (float(t), m._engine.params.clone())
    del m
same = bool(torch.equal(res["1"][1], res["0"][1]))
other = res["0"][1].clone(); dist.broadcast(other, src=0)
in_sync = bool(torch.equal(other, res["0"][1]))
if rank == 0:
    print(json.dumps({"world": world, "ms_per_step_single_allreduce": res["1"][0], "ms_per_step_overlapped": res["0"][0],
                      "params_bit_identical": same, "replicas_in_sync": in_sync}))
dist.destroy_process_group()
