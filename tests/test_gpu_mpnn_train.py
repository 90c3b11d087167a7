"""Training (backward) of the MPNN edge-network layers and MPNNModel against the float64 autograd of the oracle.

The reference's torch port of these layers is forward-only; the Keras originals (models/layers.py:3648-3887,
graph_models.py:1045-1247) train.  The oracle restates the forward formulas in torch-CPU (oracle/mpnn_torch.py, pinned
to reference outputs in tests/test_oracle_mpnn.py); its float64 autograd is the gradient reference here.  Bar: 1e-5 of
the tensor scale per gradient tensor for single layers (TF32x3 GEMMs), 1e-4 for chains of several steps (the forward
already carries a few 1e-6 per GEMM, cf. tests/test_gpu_mpnn.py)."""
import numpy as np
import pytest
import torch

from oracle import mpnn_torch as O

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def _rel(a, b):
    b = b.double()
    return float((a.detach().cpu().double() - b).abs().max() / max(float(b.abs().max()), 1e-30))


def _weave_batch(n_mols, n_atom_feat, n_pair_feat, seed, lo=3, hi=9):
    """Dense all-pairs molecules as WeaveFeaturizer yields them (every ordered pair of a molecule, row-major)."""
    rng = np.random.RandomState(seed)
    sizes = rng.randint(lo, hi + 1, size=n_mols)
    af = rng.randn(int(sizes.sum()), n_atom_feat).astype(np.float32)
    a2p, split, start = [], [], 0
    for im, n in enumerate(sizes):
        c0, c1 = np.meshgrid(np.arange(n), np.arange(n))
        a2p.append(np.transpose(np.array([c1.flatten() + start, c0.flatten() + start])))
        split.extend([im] * n)
        start += n
    a2p = np.concatenate(a2p, 0)
    pf = rng.randn(a2p.shape[0], n_pair_feat).astype(np.float32)
    return af, pf, np.array(split), a2p


def test_edge_network_gradients():
    from deepchem_b200.mpnn import EdgeNetwork
    dev = _cuda()
    P, h = 6, 32
    _, pf, _, a2p = _weave_batch(12, h, P, seed=1)
    n = int(a2p[:, 0].max()) + 1
    torch.manual_seed(0)
    layer = EdgeNetwork(P, h, trainable=True).to(dev)
    with torch.no_grad():
        layer.b.copy_(torch.randn(h * h) * 0.05)
    x = torch.randn(n, h, device=dev, requires_grad=True)
    go = torch.randn(n, h, device=dev)
    out = layer([torch.from_numpy(pf), x, torch.from_numpy(a2p)])
    out.backward(go)
    W64 = layer.W.detach().cpu().double().requires_grad_()
    b64 = layer.b.detach().cpu().double().requires_grad_()
    x64 = x.detach().cpu().double().requires_grad_()
    ref = O.edge_network(torch.from_numpy(pf).double(), x64, torch.from_numpy(a2p).long(), W64, b64, h)
    ref.backward(go.cpu().double())
    assert _rel(out, ref.detach()) < 1e-5
    assert _rel(x.grad, x64.grad) < 1e-5
    assert _rel(layer.W.grad, W64.grad) < 1e-5
    assert _rel(layer.b.grad, b64.grad) < 1e-5
    # the frozen layer (the torch port's semantics) computes the same forward
    frozen = EdgeNetwork(P, h)
    frozen.W, frozen.b = layer.W.detach().cpu(), layer.b.detach().cpu()
    assert torch.equal(frozen([torch.from_numpy(pf), x.detach(), torch.from_numpy(a2p)]), out.detach())


def test_gru_gradients():
    from deepchem_b200.mpnn import GatedRecurrentUnit
    dev = _cuda()
    n, h = 300, 64
    torch.manual_seed(1)
    layer = GatedRecurrentUnit(h, trainable=True).to(dev)
    names = ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh")
    with torch.no_grad():
        for nm in ("bz", "br", "bh"):
            getattr(layer, nm).copy_(torch.randn(h) * 0.1)
    hp = torch.randn(n, h, device=dev, requires_grad=True)
    x = torch.randn(n, h, device=dev, requires_grad=True)
    go = torch.randn(n, h, device=dev)
    out = layer([hp, x])
    out.backward(go)
    p64 = [getattr(layer, nm).detach().cpu().double().requires_grad_() for nm in names]
    hp64, x64 = hp.detach().cpu().double().requires_grad_(), x.detach().cpu().double().requires_grad_()
    ref = O.gated_recurrent_unit(hp64, x64, *p64)
    ref.backward(go.cpu().double())
    assert _rel(out, ref.detach()) < 1e-5
    assert _rel(hp.grad, hp64.grad) < 1e-5 and _rel(x.grad, x64.grad) < 1e-5
    for nm, q in zip(names, p64):
        assert _rel(getattr(layer, nm).grad, q.grad) < 1e-5, nm


def test_set_gather_gradients():
    from deepchem_b200.mpnn import SetGather
    dev = _cuda()
    h, B, M = 32, 10, 3
    rng = np.random.RandomState(3)
    sizes = rng.randint(0, 8, size=B)
    sizes[2] = 0                                               # an empty molecule
    split = np.repeat(np.arange(B), sizes)
    torch.manual_seed(2)
    layer = SetGather(M, B, n_hidden=h, trainable=True).to(dev)
    x = torch.randn(len(split), h, device=dev, requires_grad=True)
    go = torch.randn(B, 2 * h, device=dev)
    out = layer([x, split])
    out.backward(go)
    U64 = layer.U.detach().cpu().double().requires_grad_()
    b64 = layer.b.detach().cpu().double().requires_grad_()
    x64 = x.detach().cpu().double().requires_grad_()
    ref = O.set_gather_d(x64, split, U64, b64, M, B, h)
    ref.backward(go.cpu().double())
    assert _rel(out, ref.detach()) < 1e-5
    assert _rel(x.grad, x64.grad) < 2e-5
    assert _rel(layer.U.grad, U64.grad) < 2e-5 and _rel(layer.b.grad, b64.grad) < 2e-5


def _oracle_params(net, dtype=torch.float64):
    mp = net.message_passing
    c = lambda t: t.detach().cpu().to(dtype).requires_grad_()       # noqa: E731
    gru = mp.update_function
    return {"enn": (c(mp.message_function.W), c(mp.message_function.b)),
            "gru": tuple(c(getattr(gru, nm)) for nm in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh")),
            "atom_dense": (c(net.atom_dense_kernel), c(net.atom_dense_bias)),
            "set_gather": (c(net.set_gather.U), c(net.set_gather.b)),
            "dense1": (c(net.dense1_kernel), c(net.dense1_bias)), "out": (c(net.out_kernel), c(net.out_bias))}


class _Mol(object):
    def __init__(self, af, pf):
        self.af, self.pf = af, pf

    def get_num_atoms(self):
        return self.af.shape[0]

    def get_atom_features(self):
        return self.af

    def get_pair_features(self):
        return self.pf


def _dataset(n_mols, n_atom_feat, n_pair_feat, n_tasks, seed):
    from deepchem_b200.data import NumpyDataset
    rng = np.random.RandomState(seed)
    mols = np.empty(n_mols, dtype=object)
    for i in range(n_mols):
        n = int(rng.randint(3, 9))
        mols[i] = _Mol(rng.randn(n, n_atom_feat).astype(np.float32), rng.randn(n, n, n_pair_feat).astype(np.float32))
    y = rng.randn(n_mols, n_tasks).astype(np.float32)
    w = (rng.rand(n_mols, n_tasks) > 0.2).astype(np.float32)
    return NumpyDataset(mols, y, w)


@pytest.mark.parametrize("mode", ["regression", "classification"])
def test_mpnn_model_step_against_oracle(mode):
    """One batch through MPNNModel: outputs, loss and every parameter gradient against the float64 oracle; then the
    model trains (the loss of the same batches falls)."""
    from deepchem_b200.mpnn import MPNNModel
    _cuda()
    n_tasks, h, T, M, B = 2, 32, 2, 2, 8
    ds = _dataset(13, 16, 5, n_tasks, seed=4)
    if mode == "classification":
        ds = type(ds)(ds.X, (ds.y > 0).astype(np.float32), ds.w)
    torch.manual_seed(0)
    m = MPNNModel(n_tasks, n_atom_feat=16, n_pair_feat=5, n_hidden=h, T=T, M=M, mode=mode, batch_size=B,
                  learning_rate=1e-3)
    inputs, labels, weights = next(m.default_generator(ds, deterministic=True))
    ins = m._to_inputs(inputs)
    k = ins[4]
    m.grads.zero_()
    outs = m.model(ins)
    loss = m._loss(outs, np.asarray(labels[0])[:k], np.asarray(weights[0])[:k])
    loss.backward()
    params = _oracle_params(m.model)
    oin = [torch.from_numpy(np.asarray(inputs[0])).double(), torch.from_numpy(np.asarray(inputs[1])).double(),
           inputs[2], inputs[3], inputs[4]]
    oo = O.mpnn_model(params, oin, T, M, h, B, mode=mode, n_tasks=n_tasks, n_classes=2)
    y64 = torch.from_numpy(np.asarray(labels[0])[:k]).double()
    w64 = torch.from_numpy(np.asarray(weights[0])[:k]).double()
    if mode == "classification":
        per = -(y64 * torch.log_softmax(oo[1], dim=-1)).sum(-1)
    else:
        per = (oo[0] - y64.reshape(oo[0].shape)) ** 2
    while w64.dim() < per.dim():
        w64 = w64.unsqueeze(-1)
    l64 = (per * w64).mean()
    l64.backward()
    assert _rel(outs[0], oo[0].detach()) < 1e-4
    assert abs(float(loss) - float(l64)) < 1e-4 * max(1.0, abs(float(l64)))
    net, gru = m.model, m.model.message_passing.update_function
    pairs = [(net.message_passing.message_function.W, params["enn"][0]), (net.message_passing.message_function.b, params["enn"][1]),
             (net.atom_dense_kernel, params["atom_dense"][0]), (net.atom_dense_bias, params["atom_dense"][1]),
             (net.set_gather.U, params["set_gather"][0]), (net.set_gather.b, params["set_gather"][1]),
             (net.dense1_kernel, params["dense1"][0]), (net.dense1_bias, params["dense1"][1]),
             (net.out_kernel, params["out"][0]), (net.out_bias, params["out"][1])]
    pairs += [(getattr(gru, nm), q) for nm, q in zip(("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh"), params["gru"])]
    for p, q in pairs:
        assert p.grad is not None and _rel(p.grad, q.grad) < 2e-4
    first = m.fit(ds, nb_epoch=1, deterministic=True)
    for _ in range(15):
        last = m.fit(ds, nb_epoch=1, deterministic=True)
    assert last < first
    pred = m.predict(ds)
    assert pred.shape[0] == 13 and np.isfinite(pred).all()
