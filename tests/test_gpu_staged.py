"""Molecule-group staged kernels (csrc/molgroup_kernels.cu: shared-memory staging by bulk copies) against the
one-thread-per-16-bytes CSR kernels and the oracle.  Same arithmetic in the same order => bit-identical."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import oracle_batch, rel_err
from oracle import graphconv_torch as O

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def _topo(shape, n, seed, segs=None, group_rows=None):
    import os
    from deepchem_b200 import mol_graphs as MG
    from deepchem_b200.synthetic import make_molecules
    pm = make_molecules(n, seed=seed, shape=shape)
    old = os.environ.get("DCGC_GROUP_ROWS")
    if group_rows:                      # wide rows: fewer rows per group so that two stages fit in shared memory
        os.environ["DCGC_GROUP_ROWS"] = str(group_rows)
    try:
        lay = MG.BatchLayout.build(pm, n_segments=segs or n)
    finally:
        if group_rows:
            if old is None:
                os.environ.pop("DCGC_GROUP_ROWS", None)
            else:
                os.environ["DCGC_GROUP_ROWS"] = old
    return pm, lay, lay.to_device(_cuda())


CASES = [("zinc", 700, 1, 128), ("stress", 300, 2, 128), ("stress", 300, 3, 76), ("delaney", 64, 4, 64),
         ("zinc", 3, 5, 32), ("qm9", 500, 6, 300)]


@pytest.mark.parametrize("shape,n,seed,width", CASES)
def test_staged_gather_sum_is_bit_identical(shape, n, seed, width):
    from deepchem_b200 import ops
    pm, lay, topo = _topo(shape, n, seed, segs=n + 3, group_rows=48 if width > 256 else None)
    assert topo.n_groups > 0 and ops.mg_supported(topo, width)
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(topo.n_atoms, width, generator=g).cuda()
    add = torch.randn(topo.n_atoms, width, generator=g).cuda()
    ref = ops.gather_sum(x, topo.row_ptr, topo.col_idx, topo.n_atoms)
    out = ops.neighbor_sum(x, topo)
    assert torch.equal(out, ref)
    ref_t = ops.gather_sum(x, topo.t_row_ptr, topo.t_src, topo.n_atoms, addend=add.clone())
    out_t = ops.neighbor_sum(x, topo, transposed=True, addend=add.clone())
    assert torch.equal(out_t, ref_t)
    # against the oracle's index arithmetic (float64 sum of the same rows)
    xs = x.double().cpu().numpy()
    want = np.zeros_like(xs)
    rows = np.repeat(np.arange(lay.n_atoms), np.diff(lay.row_ptr))
    np.add.at(want, rows, xs[lay.col_idx])
    assert rel_err(out.cpu().numpy(), want) < 1e-6


@pytest.mark.parametrize("shape,n,seed,width", CASES)
@pytest.mark.parametrize("affine", [False, True])
def test_staged_pool_is_bit_identical(shape, n, seed, width, affine):
    from deepchem_b200 import _lib, ops
    from deepchem_b200.engine import topology_struct
    pm, lay, topo = _topo(shape, n, seed)
    if width % 16:
        pytest.skip("the staged pool backward needs argmax rows that are multiples of 16 bytes")
    L = _lib.lib()
    N = topo.n_atoms
    g = torch.Generator(device="cpu").manual_seed(100 + seed)
    # many exact ties (quantised values) exercise the first-slot rule
    x = (torch.randn(N, width, generator=g) * 2).round().div(2).cuda()
    scale = (torch.rand(width, generator=g) - 0.3).cuda() if affine else None   # negative scales flip the order
    shift = torch.randn(width, generator=g).cuda() if affine else None
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    out_r, out_s = torch.empty_like(x), torch.empty_like(x)
    arg_r = torch.empty(N, width, dtype=torch.uint8, device="cuda")
    arg_s = torch.empty_like(arg_r)
    _lib.check(L.dcgc_pool_fwd(p(x), width, p(scale), p(shift), p(topo.row_ptr), p(topo.col_idx), N, width,
                               p(out_r), width, p(arg_r), width, st))
    _lib.check(L.dcgc_mg_pool_fwd(p(x), width, p(scale), p(shift), ctypes.byref(topology_struct(topo)), width,
                                  p(out_s), width, p(arg_s), width, st))
    assert torch.equal(out_s, out_r) and torch.equal(arg_s, arg_r)
    dy = torch.randn(N, width, generator=g).cuda()
    dx_r, dx_s = torch.empty_like(x), torch.empty_like(x)
    _lib.check(L.dcgc_pool_bwd(p(dy), width, p(arg_r), width, p(scale), p(topo.t_row_ptr), p(topo.t_src),
                               p(topo.t_slot), N, width, p(dx_r), width, st))
    _lib.check(L.dcgc_mg_pool_bwd(p(dy), width, p(arg_s), width, p(scale), ctypes.byref(topology_struct(topo)),
                                  width, p(dx_s), width, st))
    assert torch.equal(dx_s, dx_r)
    if not affine:
        # oracle: GraphPool forward and its autograd on the same layout
        _, mm = oracle_batch(pm.to_list())
        xo = x.cpu().clone().requires_grad_(True)
        args = [xo, torch.from_numpy(mm.deg_slice), torch.from_numpy(mm.membership)] + \
            [torch.from_numpy(a) for a in mm.get_deg_adjacency_lists()[1:]]
        po = O.graph_pool(args[0], args[1], args[3:])
        assert torch.equal(po.detach(), out_s.cpu())
        po.backward(dy.cpu())
        assert rel_err(dx_s.cpu().numpy(), xo.grad.numpy()) < 1e-6


def test_staged_kernels_refuse_what_they_cannot_do():
    from deepchem_b200 import _lib, ops
    from deepchem_b200.engine import topology_struct
    pm, lay, topo = _topo("zinc", 50, 9)
    assert not ops.mg_supported(topo, 75)          # rows not 16-byte multiples
    assert not ops.mg_supported(topo, 128, 100)    # argmax rows not 16-byte multiples
    assert not ops.mg_supported(topo, 1 << 16)     # a group does not fit in shared memory
    x = torch.randn(topo.n_atoms, 75, device="cuda")
    out = torch.empty_like(x)
    L = _lib.lib()
    rc = L.dcgc_mg_gather_sum(ctypes.c_void_p(x.data_ptr()), 75, ctypes.byref(topology_struct(topo)), 0, 75, None, 0,
                              ctypes.c_void_p(out.data_ptr()), 75, None)
    assert rc == _lib.DCGC_ERR_INVALID and b"not supported" in L.dcgc_last_error()
    # and the dispatching op still answers through the CSR kernel
    ref = ops.gather_sum(x, topo.row_ptr, topo.col_idx, topo.n_atoms)
    assert torch.equal(ops.neighbor_sum(x, topo), ref)


def test_engine_step_is_identical_with_and_without_staging():
    """The fused engine with the staged kernels (default) and with DCGC_NO_STAGED=1 semantics (a topology whose
    group table is hidden) produces bit-identical losses and outputs, and the same gradients."""
    from deepchem_b200 import mol_graphs as MG
    from deepchem_b200.graphconvmodel import GraphConvModel
    from deepchem_b200.synthetic import make_molecules, make_labels
    dev = _cuda()
    pm = make_molecules(256, seed=21, shape="zinc")
    y, w = make_labels(256, 2, seed=1)
    res = []
    for hide in (False, True):
        torch.manual_seed(0)
        m = GraphConvModel(2, graph_conv_layers=[64, 64], dense_layer_size=128, mode="regression", batch_size=256,
                           device=dev, gemm_mode="tf32x3")
        lay = MG.BatchLayout.build(pm, n_segments=256)
        topo = lay.to_device(dev)
        if hide:
            topo.n_groups, topo._c_struct = 0, None
        eng = m._engine
        x = torch.zeros(topo.n_atoms, 76, device=dev)
        x[:, :75] = torch.from_numpy(pm.features).to(dev)[topo.perm.long()]
        loss = eng.train_step(topo, x, torch.from_numpy(y).to(dev), torch.from_numpy(w).to(dev), 256)
        res.append((float(loss), eng.grads.clone()))
    # the forward pass is bit-identical; in the backward pass the staged pool kernel also produces the BatchNorm
    # column sums (dcgc_mg_pool_bwd_stats), whose float summation order differs from the separate statistics pass
    assert res[0][0] == res[1][0]
    scale = float(res[1][1].abs().max())
    assert float((res[0][1] - res[1][1]).abs().max()) < 2e-6 * scale


@pytest.mark.parametrize("shape,n,seed,width", [("zinc", 700, 1, 128), ("stress", 300, 2, 128), ("delaney", 64, 4, 64),
                                                ("zinc", 3, 5, 32), ("zinc", 4096, 6, 128)])
def test_staged_pool_bwd_with_fused_batchnorm_sums(shape, n, seed, width):
    """dcgc_mg_pool_bwd_stats: dx bit-identical to dcgc_mg_pool_bwd, and the per-CTA partials add up to the float64
    column sums of dx and dx * y (the stage-1 input of bn_bwd_finalize)."""
    from deepchem_b200 import _lib
    from deepchem_b200.engine import topology_struct
    pm, lay, topo = _topo(shape, n, seed)
    L = _lib.lib()
    N = topo.n_atoms
    g = torch.Generator(device="cpu").manual_seed(300 + seed)
    x = (torch.randn(N, width, generator=g) * 2).round().div(2).cuda()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    out = torch.empty_like(x)
    arg = torch.empty(N, width, dtype=torch.uint8, device="cuda")
    ts = ctypes.byref(topology_struct(topo))
    _lib.check(L.dcgc_mg_pool_fwd(p(x), width, None, None, ts, width, p(out), width, p(arg), width, st))
    dy = torch.randn(N, width, generator=g).cuda()
    y = (torch.randn(N, width, generator=g) * 3 + 5).cuda()          # a mean far from zero: centring matters
    mean = y.double().mean(0).float().contiguous()
    dx_r, dx_s = torch.empty_like(x), torch.empty_like(x)
    _lib.check(L.dcgc_mg_pool_bwd(p(dy), width, p(arg), width, None, ts, width, p(dx_r), width, st))
    part = torch.full((256, 2, width), float("nan"), dtype=torch.float64, device="cuda")
    chunks = ctypes.c_int32(-1)
    _lib.check(L.dcgc_mg_pool_bwd_stats(p(dy), width, p(arg), width, ts, width, p(dx_s), width, p(y), width, p(mean),
                                        p(part), ctypes.byref(chunks), st))
    torch.cuda.synchronize()
    assert torch.equal(dx_s, dx_r)
    k = chunks.value
    assert k == min(topo.n_groups, torch.cuda.get_device_properties(0).multi_processor_count)
    assert not torch.isnan(part[:k]).any() and torch.isnan(part[k:]).all()
    got = part[:k].sum(0).cpu().numpy()
    want_a = dx_r.double().sum(0).cpu().numpy()
    want_ab = (dx_r.double() * y.double()).sum(0).cpu().numpy()
    centred = (dx_r.double() * (y.double() - mean.double())).sum(0).cpu().numpy()
    tol_a = 2e-6 * float(dx_r.abs().double().sum(0).max())
    assert np.abs(got[0] - want_a).max() < tol_a
    # the centred part carries the rounding error; the mean * sum part is exact in float64
    tol_c = 2e-6 * float((dx_r.double() * (y.double() - mean.double())).abs().sum(0).max())
    assert np.abs((got[1] - mean.double().cpu().numpy() * got[0]) - centred).max() < tol_c
    assert np.abs(got[1] - want_ab).max() < tol_c + 10 * tol_a
