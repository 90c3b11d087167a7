"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

Molecules are independent units (no edge crosses a molecule, mol_graphs.py:318-344), so the path
shards by molecule with no data-path collective.  Training adds exactly one exchange per step:
the all-reduce of the flat fp32 gradient slab.  The reference has no collective call
site of its own (it delegates to Lightning DDP, deepchem/models/trainer.py:93-101).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment; returns
    (rank, world_size, local_rank).  Single process: (0, 1, 0) and no group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GradSlab(object):
    """All gradients of a model as views of one contiguous fp32 buffer, so that the per-step
    exchange is a single all-reduce (latency-bound at ~1-4 MB: one launch, NVLS/tree inside NCCL)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(v)
            off += p.numel()

    def attach(self):
        """Point every p.grad at its slice (autograd then accumulates in place)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()

    def collect(self):
        """Copy grads that autograd replaced (p.grad no longer aliasing the slab) back in."""
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v

    def all_reduce_mean(self):
        w = world_size()
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(w)


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n items for this rank (inference sharding, no communication)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)
