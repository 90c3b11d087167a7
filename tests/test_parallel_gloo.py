"""Multi-process host logic of the data-parallel path on CPU (gloo, world_size 2): the flat gradient slab
all-reduce, the equal-shard property that makes averaged gradients exact, and the
communication-free inference sharding.  The model here is the CPU oracle (test infrastructure): what is
under test is deepchem_b200.parallel, which is device-agnostic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(pm):
    from oracle.convmol_layout import OracleConvMol, agglomerate, model_inputs
    mm = agglomerate([OracleConvMol(f, a) for f, a in pm.to_list()])
    ins = [torch.from_numpy(np.asarray(a)) if isinstance(a, np.ndarray) else a for a in model_inputs(mm)]
    ins[0] = ins[0].float()
    return ins


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from deepchem_b200 import parallel
    from deepchem_b200.synthetic import make_labels, make_molecules
    from oracle import graphconv_torch as O
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world) and parallel.world_size() == world

    B = 24                                     # per-rank batch; the global batch is 2B molecules
    pm = make_molecules(world * B, seed=5, shape="zinc")
    y, wt = make_labels(world * B, 2, "regression", seed=2)
    torch.manual_seed(0)                       # identical replicas
    model = O.OracleGraphConvModel(2, [16, 16], 32, mode="regression", batch_size=B, batch_normalize=False)
    slab = parallel.GradSlab(model.parameters())
    slab.zero()
    slab.attach()
    lo, hi = parallel.shard_range(world * B, rank, world)
    assert hi - lo == B
    out = model(_inputs(pm.slice(lo, hi)))
    loss = O.standard_loss("regression", out, torch.from_numpy(y[lo:hi]), torch.from_numpy(wt[lo:hi]))
    loss.backward()
    slab.collect()
    slab.all_reduce_mean()                     # THE exchange of a training step
    # every p.grad is a view of the reduced slab
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(slab.params, slab.views))
    torch.save({"flat": slab.flat.clone(), "loss": float(loss)}, os.path.join(out_dir, "rank%d.pt" % rank))

    # inference sharding: contiguous ranges, no communication; rank order concatenation == unsharded result
    model.eval()
    with torch.no_grad():
        pred = model(_inputs(pm.slice(lo, hi)))[0]
    torch.save(pred, os.path.join(out_dir, "pred%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_slab_allreduce_and_inference_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from deepchem_b200 import parallel
    from deepchem_b200.synthetic import make_labels, make_molecules
    from oracle import graphconv_torch as O
    r0 = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    r1 = torch.load(os.path.join(str(tmp_path), "rank1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])              # both replicas hold the same averaged gradient

    # single-process gradient of the mean loss over the GLOBAL batch == the averaged per-rank gradients
    B = 24
    pm = make_molecules(2 * B, seed=5, shape="zinc")
    y, wt = make_labels(2 * B, 2, "regression", seed=2)
    torch.manual_seed(0)
    model = O.OracleGraphConvModel(2, [16, 16], 32, mode="regression", batch_size=2 * B, batch_normalize=False)
    out = model(_inputs(pm))
    loss = O.standard_loss("regression", out, torch.from_numpy(y), torch.from_numpy(wt))
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad])
    assert abs(0.5 * (r0["loss"] + r1["loss"]) - float(loss)) < 1e-6
    scale = float(flat.abs().max())
    assert float((flat - r0["flat"]).abs().max()) < 2e-6 * scale

    model.eval()
    with torch.no_grad():
        full = model(_inputs(pm))[0]
    sharded = torch.cat([torch.load(os.path.join(str(tmp_path), "pred%d.pt" % r)) for r in range(2)])
    assert float((full - sharded).abs().max()) < 1e-5 * float(full.abs().max())

    # shard_range covers [0, n) exactly once for ragged n
    for n in (0, 1, 7, 100):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


# ---------------------------------------------------------------------------------------------------------
# SyncBatchNorm1d ("offer SyncBN to reproduce single-GPU large-batch statistics")
# ---------------------------------------------------------------------------------------------------------
_SBN_ROWS = (37, 52)          # rows (atoms) held by rank 0 / rank 1: unequal on purpose
_SBN_C = 12


def _sbn_data():
    g = torch.Generator().manual_seed(11)
    x = torch.randn(sum(_SBN_ROWS), _SBN_C, generator=g) * 2.0 + 0.5
    t = torch.randn(sum(_SBN_ROWS), _SBN_C, generator=g)
    return x, t


def _sbn_module(cls):
    torch.manual_seed(4)
    bn = cls(_SBN_C, eps=1e-3, momentum=0.99)
    with torch.no_grad():
        bn.weight.add_(torch.randn(_SBN_C) * 0.3)
        bn.bias.add_(torch.randn(_SBN_C) * 0.3)
    return bn


def _sbn_worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(1)
    from deepchem_b200 import parallel
    parallel.init_from_env("gloo")
    x, t = _sbn_data()
    lo = sum(_SBN_ROWS[:rank])
    hi = lo + _SBN_ROWS[rank]
    bn = _sbn_module(parallel.SyncBatchNorm1d)
    bn.train()
    out = {}
    for step in range(2):                                     # two steps: the running statistics are updated twice
        xr = x[lo:hi].clone().requires_grad_(True)
        y = bn(xr)
        loss = (torch.tanh(y) * t[lo:hi]).sum() / 10.0        # a sum: the global loss is the sum of the rank losses
        bn.zero_grad()
        loss.backward()
        out["y%d" % step], out["dx%d" % step] = y.detach().clone(), xr.grad.clone()
        out["dw%d" % step], out["db%d" % step] = bn.weight.grad.clone(), bn.bias.grad.clone()
    out["running_mean"], out["running_var"] = bn.running_mean.clone(), bn.running_var.clone()
    out["num_batches_tracked"] = bn.num_batches_tracked.clone()
    bn.eval()                                                 # eval mode: stock BatchNorm1d path, no collective
    out["eval"] = bn(x[lo:hi]).detach().clone()
    torch.save(out, os.path.join(out_dir, "sbn%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_batch_norm_equals_batch_norm_on_the_concatenated_rows(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_sbn_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from deepchem_b200 import parallel
    parts = [torch.load(os.path.join(str(tmp_path), "sbn%d.pt" % r)) for r in range(world)]
    x, t = _sbn_data()
    ref = _sbn_module(torch.nn.BatchNorm1d)
    ref.train()
    for step in range(2):
        xr = x.clone().requires_grad_(True)
        y = ref(xr)
        loss = (torch.tanh(y) * t).sum() / 10.0
        ref.zero_grad()
        loss.backward()
        cat = lambda k: torch.cat([p["%s%d" % (k, step)] for p in parts])       # noqa: E731
        assert torch.allclose(cat("y"), y.detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(cat("dx"), xr.grad, rtol=1e-4, atol=1e-6)
        # parameter gradients are local sums: the gradient exchange of the training step adds them up
        assert torch.allclose(sum(p["dw%d" % step] for p in parts), ref.weight.grad, rtol=1e-4, atol=1e-6)
        assert torch.allclose(sum(p["db%d" % step] for p in parts), ref.bias.grad, rtol=1e-4, atol=1e-6)
    for p in parts:
        assert torch.allclose(p["running_mean"], ref.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(p["running_var"], ref.running_var, rtol=1e-5, atol=1e-6)
        assert int(p["num_batches_tracked"]) == 2
    ref.eval()
    assert torch.allclose(torch.cat([p["eval"] for p in parts]), ref(x).detach(), rtol=1e-5, atol=1e-6)
    # single process, no group: the stock path, same state_dict keys as BatchNorm1d
    solo = _sbn_module(parallel.SyncBatchNorm1d)
    solo.train()
    plain = _sbn_module(torch.nn.BatchNorm1d)
    plain.train()
    assert torch.equal(solo(x), plain(x))
    assert list(solo.state_dict().keys()) == list(plain.state_dict().keys())
