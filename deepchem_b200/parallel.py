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


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


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


class _SyncBatchNormFn(torch.autograd.Function):
    """Training-mode batch normalisation of [rows, C] over the rows of ALL ranks: two small all-reduces per call
    (forward: per-column sum, sum of squares and the row count as one float64 vector of 2C+1; backward: the two
    per-column sums of the input gradient, 2C).  Ranks may hold different numbers of rows (atoms per batch vary)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, group):
        c = x.shape[1]
        xd = x.double()
        stats = torch.cat([xd.sum(0), (xd * xd).sum(0), xd.new_tensor([float(x.shape[0])])])
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        n = stats[2 * c]
        mean = stats[:c] / n
        var = (stats[c:2 * c] / n - mean * mean).clamp_min_(0.0)          # biased, as F.batch_norm normalises
        invstd = torch.rsqrt(var + eps)
        mean32, invstd32 = mean.float(), invstd.float()
        xhat = (x - mean32) * invstd32
        ctx.save_for_backward(xhat, weight, invstd32, n)      # (n stays on the device: no host synchronisation)
        ctx.group = group
        ctx.mark_non_differentiable(mean, var, n)
        y = xhat * weight + bias if weight is not None else xhat
        return y, mean, var, n

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar, _dn):
        xhat, weight, invstd, n = ctx.saved_tensors
        c = dy.shape[1]
        dyd = dy.double()
        sums = torch.cat([dyd.sum(0), (dyd * xhat.double()).sum(0)])
        local = sums.clone()                 # parameter gradients stay local: the gradient exchange sums them over ranks
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
        mean_dy = (sums[:c] / n).float()
        mean_dy_xhat = (sums[c:] / n).float()
        scale = invstd * weight if weight is not None else invstd
        dx = (dy - mean_dy - xhat * mean_dy_xhat) * scale
        dw = local[c:].float() if weight is not None else None
        db = local[:c].float() if weight is not None else None
        return dx, dw, db, None, None


class SyncBatchNorm1d(torch.nn.BatchNorm1d):
    """``BatchNorm1d`` whose TRAINING statistics are taken over the rows of every rank of the process group, so that a
    data-parallel run normalises with the statistics a single process would see on the concatenated batch (SURVEY 8e:
    the reference's BatchNorm is per process; this is the opt-in alternative).  Same parameters, buffers and
    ``state_dict`` keys as ``BatchNorm1d`` (eps / momentum as the model sets them: 1e-3 / 0.99,
    ``torch_models/graphconvmodel.py:160-170``); eval mode and single-process runs take the stock code path.  Works on
    any backend that all-reduces float64 (NCCL, gloo), which is what lets ``tests/test_parallel_gloo.py`` check it on
    CPU against ``BatchNorm1d`` over the concatenated rows."""

    def __init__(self, *args, process_group=None, **kwargs):
        super(SyncBatchNorm1d, self).__init__(*args, **kwargs)
        self.process_group = process_group

    def forward(self, x):
        if not (self.training and dist.is_available() and dist.is_initialized()
                and dist.get_world_size(self.process_group) > 1):
            return super(SyncBatchNorm1d, self).forward(x)
        if x.dim() != 2:
            raise ValueError("SyncBatchNorm1d expects [rows, channels], got %s" % (tuple(x.shape),))
        y, mean, var, n = _SyncBatchNormFn.apply(x, self.weight, self.bias, self.eps, self.process_group)
        if self.track_running_stats:
            with torch.no_grad():
                self.num_batches_tracked += 1
                m = self.momentum if self.momentum is not None else 1.0 / float(self.num_batches_tracked)
                unbiased = var * (n / (n - 1.0).clamp_min(1.0))
                self.running_mean.mul_(1.0 - m).add_(mean.to(self.running_mean.dtype), alpha=m)
                self.running_var.mul_(1.0 - m).add_(unbiased.to(self.running_var.dtype), alpha=m)
        return y


class PeerMailboxes(object):
    """The peer-memory mailboxes of the synchronised BatchNorm inside the fused engine (include/dcgc.h, dcgc_bn_sync):
    every rank of the node allocates one IPC-shareable device buffer, the 64-byte handles travel once through the
    process group, every rank maps the buffers of its peers.  From then on the BatchNorm finalize kernels exchange
    their column sums with plain stores over NVLink — no collective call and no host round trip inside the step.
    ``struct(n_uses)`` hands out the dcgc_bn_sync of the next step and advances the sequence number."""

    def __init__(self, cap, group=None):
        import ctypes
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerMailboxes needs an initialised process group")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.SYNC_MAX_RANKS:
            raise ValueError("at most %d ranks (one node)" % _lib.SYNC_MAX_RANKS)
        self.cap = int(cap)
        L = _lib.lib()
        nbytes = int(L.dcgc_bn_sync_mailbox_bytes(self.world, self.cap))
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _lib.check(L.dcgc_p2p_alloc(nbytes, ctypes.byref(own), handle))
        self._own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._peers = []
        self.ptrs = []
        for r in range(self.world):
            if r == self.rank:
                self.ptrs.append(self._own)
                continue
            ptr = ctypes.c_void_p()
            _lib.check(L.dcgc_p2p_open(ctypes.create_string_buffer(handles[r], 64), ctypes.byref(ptr)))
            self._peers.append(ptr.value)
            self.ptrs.append(ptr.value)
        dist.barrier(group=group)          # every mailbox is zeroed and mapped before the first kernel writes one
        self.seq = 1
        self._struct = _lib.BnSync()
        self._struct.world, self._struct.rank, self._struct.cap = self.world, self.rank, self.cap
        for r, ptr in enumerate(self.ptrs):
            self._struct.mailbox[r] = ptr

    def struct(self, n_uses):
        self._struct.seq0 = self.seq
        self.seq += int(n_uses)
        return self._struct

    def close(self):
        from . import _lib
        L = _lib.lib()
        for ptr in self._peers:
            L.dcgc_p2p_close(ptr)
        self._peers = []
        if self._own:
            L.dcgc_p2p_free(self._own)
            self._own = None
