"""Data-parallel exactness on ONE device: molecules never interact and the loss is a mean
(`torch_model.py:1290-1291`), so with equal shards the averaged per-shard gradient slab IS the whole-batch gradient.
Two engines take one half of a batch each, the slabs are added by hand (what the NCCL all-reduce does, measured
bit-identical to it on two B200s: `scripts/dmpnn_dp2.py`, `profiles/r3l_dmpnn_dp2.json`) and compared with a third
engine that sees the whole batch.  GraphConv runs without BatchNorm here: its statistics are per process in the
reference too, so a BatchNorm model on two ranks is a different function of the batch, not a rounding difference."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5       # of the largest entry of the slab; summation order is the only difference


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def _rel(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


@pytest.mark.parametrize("mode,n_tasks", [("regression", 2), ("classification", 3)])
def test_graphconv_shard_gradients_average_to_the_whole_batch_gradient(mode, n_tasks):
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    dev = _cuda()
    world, B = 2, 96
    pm = make_molecules(world * B, seed=31, shape="zinc")
    y, w = make_labels(world * B, n_tasks, mode, seed=7)

    def build(bsz):
        torch.manual_seed(3)
        m = GraphConvModel(n_tasks, graph_conv_layers=[64, 64], dense_layer_size=128, mode=mode, batch_size=bsz,
                           batch_normalize=False, device=dev, gemm_mode="tf32x3")
        assert m._engine is not None
        return m

    def grads_of(m, pm_, y_, w_):
        batch = next(m.default_generator(PackedDataset(pm_, y_, w_), deterministic=True, pad_batches=False))
        inputs, labels, weights = m._prepare_batch(batch)
        loss = m._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0].contiguous(), weights[0].contiguous(),
                                    pm_.n_mols)
        return m._engine.grads.clone(), float(loss)

    whole = build(world * B)
    g_whole, loss_whole = grads_of(whole, pm, y, w)
    g_sum, loss_sum = torch.zeros_like(g_whole), 0.0
    for r in range(world):
        shard = build(B)
        shard.model.load_state_dict(whole.model.state_dict())
        g, l = grads_of(shard, pm.slice(r * B, (r + 1) * B), y[r * B:(r + 1) * B], w[r * B:(r + 1) * B])
        g_sum += g
        loss_sum += l
    assert float(g_whole.abs().max()) > 0
    assert _rel(g_sum / world, g_whole) < TOL
    assert abs(loss_sum / world - loss_whole) < 1e-5 * max(1.0, abs(loss_whole))


def test_dmpnn_shard_gradients_average_to_the_whole_batch_gradient():
    from deepchem_b200.dmpnn import DMPNNModel, GraphDataset
    from deepchem_b200.dmpnn_data import make_graphs
    dev = _cuda()
    world, B = 2, 128
    pg = make_graphs(world * B, seed=9, shape="qm9")
    y = np.random.default_rng(9).standard_normal((world * B, 12)).astype(np.float32)
    w = np.ones_like(y)

    def build(bsz):
        torch.manual_seed(0)
        m = DMPNNModel(device=dev, n_tasks=12, batch_size=bsz, gemm_mode="tf32x3")
        assert m._engine is not None
        return m

    def grads_of(m, pg_, y_, w_):
        inputs, labels, weights = m._prepare_batch(next(m.default_generator(GraphDataset(pg_, y_, w_),
                                                                            deterministic=True)))
        topo = inputs.topology
        yy = labels[0].reshape(topo.n_mols, -1).contiguous()
        ww = weights[0].reshape(topo.n_mols, -1).expand_as(yy).contiguous()
        loss = m._engine.train_step(topo, inputs['atom_features'], inputs['f_ini_atoms_bonds'], yy, ww)
        return m._engine.grads.clone(), float(loss)

    whole = build(world * B)
    g_whole, loss_whole = grads_of(whole, pg, y, w)
    g_sum, loss_sum = torch.zeros_like(g_whole), 0.0
    for r in range(world):
        shard = build(B)
        shard.model.load_state_dict(whole.model.state_dict())
        g, l = grads_of(shard, pg.slice(r * B, (r + 1) * B), y[r * B:(r + 1) * B], w[r * B:(r + 1) * B])
        g_sum += g
        loss_sum += l
    assert float(g_whole.abs().max()) > 0
    assert _rel(g_sum / world, g_whole) < TOL
    assert abs(loss_sum / world - loss_whole) < 1e-5 * max(1.0, abs(loss_whole))


def test_dmpnn_two_rank_exchange_is_bit_identical_to_the_hand_summed_slabs():
    """The real thing on two devices (skipped on a one-GPU box): torchrun, NCCL all-reduce of the flat gradient slab;
    scripts/dmpnn_dp2.py asserts the first-step gradient against the whole-batch one and bit-identity with the
    in-process emulation over 4 Adam steps."""
    import os
    import subprocess
    import sys
    _cuda()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DP2_NO_TIMING="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(root, "scripts", "dmpnn_dp2.py")], env=env, cwd=root, capture_output=True,
                       text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    verdict = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert verdict["ok"], verdict


def test_sync_batch_norm_model_on_one_process_equals_the_default_model():
    """sync_batch_norm=True selects parallel.SyncBatchNorm1d and the per-layer autograd path; with one process it must
    be the default model: same state_dict keys, same loss and gradients as the fused engine (the cross-rank arithmetic
    of the layer itself is checked on CPU under gloo, tests/test_parallel_gloo.py)."""
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.parallel import SyncBatchNorm1d
    from deepchem_b200.synthetic import make_labels, make_molecules
    dev = _cuda()
    B = 128
    pm = make_molecules(B, seed=33, shape="zinc")
    y, w = make_labels(B, 2, "regression", seed=8)
    torch.manual_seed(5)
    a = GraphConvModel(2, graph_conv_layers=[64, 64], dense_layer_size=128, mode="regression", batch_size=B, device=dev,
                       gemm_mode="tf32x3")
    b = GraphConvModel(2, graph_conv_layers=[64, 64], dense_layer_size=128, mode="regression", batch_size=B, device=dev,
                       gemm_mode="tf32x3", sync_batch_norm=True)
    assert a._engine is not None and b._engine is None
    assert all(isinstance(m, SyncBatchNorm1d) for m in b.model.batch_norms)
    assert list(a.model.state_dict().keys()) == list(b.model.state_dict().keys())
    b.model.load_state_dict(a.model.state_dict())
    batch = next(a.default_generator(PackedDataset(pm, y, w), deterministic=True, pad_batches=False))
    inputs, labels, weights = a._prepare_batch(batch)
    loss_a = float(a._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0].contiguous(),
                                        weights[0].contiguous(), B))
    g_a = {n: p.grad.clone() for n, p in a.model.named_parameters()}
    b.model.train()
    outs = b.model(inputs)
    loss_b = b._loss_fn([outs[0]], labels, weights)
    loss_b.backward()
    assert abs(loss_a - float(loss_b)) < 1e-5 * max(1.0, abs(loss_a))
    for n, p in b.model.named_parameters():
        scale = max(float(g_a[n].abs().max()), 1e-6)
        assert float((p.grad - g_a[n]).abs().max()) < 1e-4 * scale + 1e-9, n


def test_synchronised_batchnorm_in_the_engine_equals_one_process_on_the_whole_batch():
    """dcgc_gcmodel_train_step_sync (SURVEY 8e opt-in): two "ranks" — two engines on two streams of ONE device, their
    mailboxes two plain allocations of this process, so no IPC — exchange the BatchNorm column sums through the
    mailboxes from inside the finalize kernels.  With statistics over both shards the model IS the single-process
    function of the concatenated batch: averaged loss, averaged gradient slab and running statistics must match one
    engine on the whole batch, and both ranks must hold bit-identical running statistics.  (The two-GPU version over
    NVLink with IPC-mapped mailboxes: scripts/syncbn_check.py, profiles/r5s_syncbn_2gpu.json.)"""
    import ctypes
    from deepchem_b200 import _lib
    from deepchem_b200.data import PackedDataset
    from deepchem_b200.engine import _ws, topology_struct
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_labels, make_molecules
    dev = _cuda()
    world, B = 2, 160
    pm = make_molecules(world * B, seed=33, shape="zinc")
    y, w = make_labels(world * B, 2, "regression", seed=8)

    def build(bsz):
        torch.manual_seed(5)
        m = GraphConvModel(2, graph_conv_layers=[64, 128], dense_layer_size=128, mode="regression", batch_size=bsz,
                           device=dev, gemm_mode="tf32x3")
        assert m._engine is not None and m._engine.cfg.batch_norm == 1
        m.model.train()
        return m

    def prepared(m, pm_, y_, w_):
        batch = next(m.default_generator(PackedDataset(pm_, y_, w_), deterministic=True, pad_batches=False))
        return m._prepare_batch(batch)

    whole = build(world * B)
    inputs, labels, weights = prepared(whole, pm, y, w)
    loss_whole = float(whole._engine.train_step(inputs[1]._dcgc_topology, inputs[0], labels[0].contiguous(),
                                                weights[0].contiguous(), world * B))
    g_whole, bn_whole = whole._engine.grads.clone(), whole._engine.bn_running.clone()

    L = _lib.lib()
    cap = 128
    nbytes = int(L.dcgc_bn_sync_mailbox_bytes(world, cap))
    boxes, handle = [], ctypes.create_string_buffer(64)
    for _ in range(world):
        ptr = ctypes.c_void_p()
        _lib.check(L.dcgc_p2p_alloc(nbytes, ctypes.byref(ptr), handle))
        boxes.append(ptr.value)
    try:
        shards, preps, streams = [], [], [torch.cuda.Stream(device=dev) for _ in range(world)]
        for r in range(world):
            s = build(B)
            s.model.load_state_dict(whole.model.state_dict(), strict=False)
            for b_ in s.model.batch_norms:
                b_.reset_running_stats()
            s._engine.adopt()
            shards.append(s)
            preps.append(prepared(s, pm.slice(r * B, (r + 1) * B), y[r * B:(r + 1) * B], w[r * B:(r + 1) * B]))
            # one plain step first: every kernel of the step is loaded before a kernel of this device waits for a peer
            # (CUDA loads modules lazily, and a load behind a waiting kernel would block the host: see model.cu)
            ins, lab, wts = preps[r]
            s._engine.train_step(ins[1]._dcgc_topology, ins[0], lab[0].contiguous(), wts[0].contiguous(), B)
            for b_ in s.model.batch_norms:
                b_.reset_running_stats()
        torch.cuda.synchronize()
        for r in range(world):                    # enqueue rank r's whole step on its own stream; nothing blocks the host
            eng = shards[r]._engine
            ins, lab, wts = preps[r]
            topo = ins[1]._dcgc_topology
            sy = _lib.BnSync()
            sy.world, sy.rank, sy.cap, sy.seq0 = world, r, cap, 1
            for q in range(world):
                sy.mailbox[q] = boxes[q]
            ws = torch.empty(int(L.dcgc_gcmodel_workspace_bytes(ctypes.byref(eng.cfg), topo.n_atoms, topo.n_segments)) + 256,
                             dtype=torch.uint8, device=dev)
            eng._test_ws = ws
            yb, wb = lab[0].contiguous(), wts[0].contiguous()
            eng._test_keep = (yb, wb)
            with torch.cuda.stream(streams[r]):
                _lib.check(L.dcgc_gcmodel_train_step_sync(
                    ctypes.byref(eng.cfg), ctypes.byref(topology_struct(topo)), ins[0].data_ptr(), ins[0].stride(0),
                    yb.data_ptr(), wb.data_ptr(), B, eng.params.data_ptr(), eng.grads.data_ptr(), eng.bn_running.data_ptr(),
                    ws.data_ptr(), ws.numel(), eng.loss.data_ptr(), None, None, None, 0, ctypes.byref(sy),
                    ctypes.c_void_p(streams[r].cuda_stream)))
        torch.cuda.synchronize()
        g_avg = sum(s._engine.grads for s in shards) / world
        loss_avg = sum(float(s._engine.loss) for s in shards) / world
        assert torch.equal(shards[0]._engine.bn_running, shards[1]._engine.bn_running)
        assert _rel(shards[0]._engine.bn_running, bn_whole) < 1e-6
        assert abs(loss_avg - loss_whole) < 1e-5 * max(1.0, abs(loss_whole))
        assert float(g_whole.abs().max()) > 0
        assert _rel(g_avg, g_whole) < TOL
    finally:
        torch.cuda.synchronize()
        for ptr in boxes:
            L.dcgc_p2p_free(ptr)
