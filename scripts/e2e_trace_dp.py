"""Per-stage host timers of the end-to-end DATA-PARALLEL fit (torchrun, one rank per GPU), optionally with every rank
pinned to CORES_PER_RANK cores (rank r -> cores [r * C, (r + 1) * C)) to reproduce the 4-cores-per-rank host of an
8-GPU box on fewer GPUs.  Rank 0 prints."""
import itertools
import os
import sys
import time

os.environ["DCGC_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", "0"))
cpr = int(os.environ.get("CORES_PER_RANK", "0"))
if cpr:
    os.sched_setaffinity(0, set(range(rank * cpr, (rank + 1) * cpr)))
import numpy as np
import torch
import torch.distributed as dist

from deepchem_b200 import parallel
from deepchem_b200.data import PackedDataset
from deepchem_b200.graphconvmodel import GraphConvModel
from deepchem_b200.synthetic import PackedMols, make_labels, make_molecules

rank, world, local = parallel.init_from_env("nccl")
dev = torch.device("cuda", local)
B = 4096
pool = [make_molecules(B, seed=4 * rank + i) for i in range(4)]
big = PackedMols.concat(pool).pin_memory()
y, w = make_labels(4 * B, 1, "regression", seed=rank)
ds = PackedDataset(big, y, w)
m = GraphConvModel(1, [128, 128, 128], 128, mode="regression", batch_size=B, device=dev, gemm_mode="tf32x3")
if world > 1:
    m.enable_data_parallel()
m.model.train()
m.log_frequency = 1
m.fit_generator(itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True), 40), checkpoint_interval=0)
torch.cuda.synchronize()
dist.barrier()
m._pipe_trace.clear()
K = 150
t = time.perf_counter()
m.fit_generator(itertools.islice(m.default_generator(ds, epochs=1000, deterministic=True), K), checkpoint_interval=0)
torch.cuda.synchronize()
dist.barrier()
ms = (time.perf_counter() - t) / K * 1e3
tr = m._pipe_trace
line = ("rank %d/%d cores=%s workers=%s: %.3f ms/step | fit thread: wait for batch %.3f, _train_step host %.3f | prefetch "
        "thread: wait generator %.3f, wait slot %.3f, _prepare_batch %.3f, wait queue %.3f (ms per step)" % (
            rank, world, sorted(os.sched_getaffinity(0)) if cpr else "all", m.host_workers, ms,
            tr["fit_wait_batch"] / K * 1e3, tr["fit_train_step_host"] / K * 1e3, tr["pf_wait_generator"] / K * 1e3,
            tr["pf_wait_slot"] / K * 1e3, tr["pf_prepare"] / K * 1e3, tr["pf_wait_queue"] / K * 1e3))
extra = {k: round(v / K * 1e3, 3) for k, v in tr.items() if k not in ("fit_wait_batch", "fit_train_step_host", "fit_steps",
         "pf_wait_generator", "pf_wait_slot", "pf_prepare", "pf_wait_queue")}
for r in range(world):
    if r == rank and (r == 0 or r == world - 1):
        print(line, extra, flush=True)
    dist.barrier()
dist.destroy_process_group()
