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
