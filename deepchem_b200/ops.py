"""PyTorch custom ops (autograd Functions) over the C ABI of libdcgc.

PyTorch owns device memory and the stream; every compute call goes through ctypes into
hand-written sm_100a kernels (include/dcgc.h).  There is no fallback: tensors must be CUDA
float32 tensors and the library must load.
"""
import ctypes

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_TANH, GEMM_FP32, check

_ACT_CODES = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "tanh": ACT_TANH}


def act_code(act):
    """Map an activation spec (None / 'relu' / 'tanh' / torch function) to a fused epilogue code;
    returns None if the activation cannot be fused (caller applies it with torch)."""
    if act in _ACT_CODES:
        return _ACT_CODES[act]
    if act in (torch.relu, torch.nn.functional.relu):
        return ACT_RELU
    if act in (torch.tanh, torch.nn.functional.tanh):
        return ACT_TANH
    if isinstance(act, torch.nn.ReLU):
        return ACT_RELU
    if isinstance(act, torch.nn.Tanh):
        return ACT_TANH
    return None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_dev(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the deepchem_b200 ops have no CPU path" % name)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))


def _rowmajor(t):
    """Return a tensor whose rows are contiguous (stride(1) == 1); views with a padded leading
    dimension are used as they are."""
    if t.dim() != 2:
        raise ValueError("expected a 2-D tensor")
    if t.shape[1] > 1 and t.stride(1) != 1:
        return t.contiguous()
    if t.shape[1] <= 1 and t.shape[0] > 1 and t.stride(0) < 1:
        return t.contiguous()
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def padded_empty(n_rows, width, device, dtype=torch.float32):
    """[n_rows, width] view of a buffer whose leading dimension is a multiple of 4 floats, so
    that every row starts 16-byte aligned (pad columns are zero)."""
    ld = (width + 3) // 4 * 4
    if ld == width:
        return torch.empty(n_rows, width, device=device, dtype=dtype)
    buf = torch.zeros(n_rows, ld, device=device, dtype=dtype)
    return buf[:, :width]


# ------------------------------------------------------------------------------------------
# launch accounting (bench.py: gpu_launches, per-entry-point CUDA-event timing for the roofline);
# both live inside the library so that launches made by the fused engine are seen too.
# ------------------------------------------------------------------------------------------
def launch_count():
    """Number of libdcgc kernels launched by this process so far."""
    return int(_lib.lib().dcgc_launch_count())


def profile_begin(name):
    """Bracket every call of the C entry point `name` with CUDA events on its launching stream."""
    check(_lib.lib().dcgc_profile_begin(name.encode()))


def profile_end():
    """-> {'ms': summed event durations, 'launches': bracketed calls}"""
    ms, n = ctypes.c_double(), ctypes.c_int64()
    check(_lib.lib().dcgc_profile_end(ctypes.byref(ms), ctypes.byref(n)))
    return {"ms": ms.value, "launches": n.value}


def profile_report():
    """Ends a profile_begin("*") session -> {scope: (total_ms, calls)}."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(_lib.lib().dcgc_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, n = line.rsplit(" ", 2)
        out[name] = (float(ms), int(n))
    return out


def _count(n=1):
    pass


# ------------------------------------------------------------------------------------------
# raw launchers (no autograd)
# ------------------------------------------------------------------------------------------
def gather_sum(x, row_ptr, idx, n_rows_out, addend=None, out=None):
    x = _rowmajor(x)
    width = x.shape[1]
    if out is None:
        out = addend if addend is not None else padded_empty(n_rows_out, width, x.device)
    check(_lib.lib().dcgc_gather_sum(_p(x), _ld(x), _p(row_ptr), _p(idx), n_rows_out, width,
                                     _p(addend), _ld(addend) if addend is not None else 0,
                                     _p(out), _ld(out), _stream()))
    return out


def _topo_struct(topo):
    from .engine import topology_struct
    return ctypes.byref(topology_struct(topo))


def mg_supported(topo, ld_floats, ld_arg_bytes=0):
    """True if the molecule-group staged kernels (csrc/molgroup_kernels.cu) can run on this topology with rows
    of ld_floats floats (+ ld_arg_bytes argmax bytes): valid group table, symmetric adjacency, rows fit in
    shared memory."""
    if topo is None or not getattr(topo, "n_groups", 0) or not getattr(topo, "symmetric", False):
        return False
    return bool(_lib.lib().dcgc_mg_supported(_topo_struct(topo), int(ld_floats), int(ld_arg_bytes)))


def _al16(*ts):
    return all(t is None or t.data_ptr() % 16 == 0 for t in ts)


def neighbor_sum(x, topo, transposed=False, addend=None):
    """K1 / K5 over a DeviceTopology: the staged kernel when the layout allows it, else the CSR gather."""
    x = _rowmajor(x)
    n, width = topo.n_atoms, x.shape[1]
    if n and width % 4 == 0 and _ld(x) % 4 == 0 and mg_supported(topo, _ld(x)) and \
            (addend is None or _ld(addend) % 4 == 0):
        out = addend if addend is not None else padded_empty(n, width, x.device)
        if _al16(x, out):
            check(_lib.lib().dcgc_mg_gather_sum(_p(x), _ld(x), _topo_struct(topo), 1 if transposed else 0, width,
                                                _p(addend), _ld(addend) if addend is not None else 0,
                                                _p(out), _ld(out), _stream()))
            return out
    if transposed:
        return gather_sum(x, topo.t_row_ptr, topo.t_src, n, addend=addend)
    return gather_sum(x, topo.row_ptr, topo.col_idx, n, addend=addend)


def _permute_rows_i8(src, perm, n_feat=None, ld_out=None, out=None):
    """int8 feature rows -> degree-major fp32 rows (exact), pad columns zeroed."""
    if not src.is_cuda:
        raise RuntimeError("src must be a CUDA tensor: the deepchem_b200 ops have no CPU path")
    if src.dim() != 2 or src.stride(1) != 1:
        src = src.contiguous()
    n_feat = src.shape[1] if n_feat is None else n_feat
    n = perm.shape[0]
    ld_out = ld_out or (n_feat + 3) // 4 * 4
    if out is None:
        out = torch.empty(n, ld_out, device=src.device, dtype=torch.float32)
    else:
        out = out[:n * ld_out].view(n, ld_out)
    check(_lib.lib().dcgc_permute_rows_i8(_p(src), src.stride(0), _p(perm), n, n_feat, _p(out), ld_out, _stream()))
    _count()
    return out[:, :n_feat]


def permute_rows(src, perm, n_feat=None, ld_out=None, out=None):
    if src.dtype == torch.int8:
        return _permute_rows_i8(src, perm, n_feat, ld_out, out)
    src = _rowmajor(src)
    n_feat = src.shape[1] if n_feat is None else n_feat
    n = perm.shape[0]
    ld_out = ld_out or (n_feat + 3) // 4 * 4
    if out is None:
        out = torch.empty(n, ld_out, device=src.device, dtype=torch.float32)
    else:
        out = out[:n * ld_out].view(n, ld_out)
    check(_lib.lib().dcgc_permute_rows(_p(src), _ld(src), _p(perm), n, n_feat, _p(out), ld_out, _stream()))
    _count()
    return out[:, :n_feat]


def group_gemm_fwd(a1, a2, w, bias, topo, act, mode=GEMM_FP32):
    """y = act([a1|a2] . w[g] + bias[g]); topo=None means a single group over all rows."""
    n_rows, k1 = a1.shape
    k2 = a2.shape[1] if a2 is not None else 0
    n = w.shape[-1]
    y = padded_empty(n_rows, n, a1.device)
    tiles = topo.tiles if topo is not None else None
    check(_lib.lib().dcgc_group_gemm_fwd(
        mode, _p(a1), _ld(a1), k1, _p(a2), _ld(a2) if a2 is not None else 0, k2, _p(w), _p(bias), n,
        _p(tiles), topo.n_tiles if topo is not None else 0, _lib.TILE_ROWS, n_rows, act, _p(y), _ld(y),
        _stream()))
    _count()
    return y


def group_gemm_dgrad(g, w, k1, k2, topo, want1=True, want2=True, mode=GEMM_FP32):
    n_rows, n = g.shape
    d1 = padded_empty(n_rows, k1, g.device) if want1 else None
    d2 = padded_empty(n_rows, k2, g.device) if (want2 and k2) else None
    tiles = topo.tiles if topo is not None else None
    check(_lib.lib().dcgc_group_gemm_dgrad(
        mode, _p(g), _ld(g), n, _p(w), k1, k2, _p(tiles), topo.n_tiles if topo is not None else 0,
        _lib.TILE_ROWS, n_rows, _p(d1), _ld(d1) if d1 is not None else 0, _p(d2),
        _ld(d2) if d2 is not None else 0, _stream()))
    _count()
    return d1, d2


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.type, device.index)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def group_gemm_wgrad(a1, a2, g, topo, n_groups, mode=GEMM_FP32):
    """dW [n_groups, k1+k2, n], dbias [n_groups, n]; deterministic."""
    n_rows, k1 = a1.shape
    k2 = a2.shape[1] if a2 is not None else 0
    n = g.shape[1]
    L = _lib.lib()
    dw = torch.empty(n_groups, k1 + k2, n, device=g.device, dtype=torch.float32)
    db = torch.empty(n_groups, n, device=g.device, dtype=torch.float32)
    nbytes = int(L.dcgc_group_gemm_wgrad_workspace(k1, k2, n, n_groups))
    ws = _workspace(nbytes, g.device)
    if topo is not None:
        counts = topo._deg_count_c
    else:
        counts = (ctypes.c_int64 * 1)(n_rows)
    check(L.dcgc_group_gemm_wgrad(mode, _p(a1), _ld(a1), k1, _p(a2), _ld(a2) if a2 is not None else 0, k2,
                                  _p(g), _ld(g), n, counts, n_groups, _p(dw), _p(db), _p(ws), nbytes,
                                  _stream()))
    _count(2)
    return dw, db


# ------------------------------------------------------------------------------------------
# autograd Functions
# ------------------------------------------------------------------------------------------
class NeighborSum(torch.autograd.Function):
    """S[i] = sum_{j in N(i)} X[j]  (K1) with the transposed gather as backward (K5)."""

    @staticmethod
    def forward(ctx, x, topo):
        _check_dev(x, "atom_features")
        ctx.topo = topo
        return neighbor_sum(x, topo)

    @staticmethod
    def backward(ctx, ds):
        topo = ctx.topo
        return neighbor_sum(_rowmajor(ds), topo, transposed=True), None


class GraphConvFn(torch.autograd.Function):
    """Fused GraphConv (K1 + K2 forward, K5 + K6 backward).

    w: [11, 2*Fp, C] packed weights (rows 0:Fp self, Fp:2Fp neighbour; Fp = padded input width),
    bias: [11, C].  x may be a [N, F] view of a zero-padded [N, Fp] buffer."""

    @staticmethod
    def forward(ctx, x, w, bias, topo, act, mode):
        _check_dev(x, "atom_features")
        x = _rowmajor(x)
        fp = w.shape[1] // 2
        xk = _widen(x, fp)
        s = neighbor_sum(xk, topo)
        y = group_gemm_fwd(xk, s, w, bias, topo, act, mode)
        ctx.topo, ctx.act, ctx.mode, ctx.f = topo, act, mode, x.shape[1]
        ctx.save_for_backward(xk, s, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xk, s, w, y = ctx.saved_tensors
        topo, mode = ctx.topo, ctx.mode
        fp = w.shape[1] // 2
        g = _rowmajor(dy)
        if ctx.act == ACT_RELU:
            g = g * (y > 0)
        elif ctx.act == ACT_TANH:
            g = g * (1 - y * y)
        dw = db = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = group_gemm_wgrad(xk, s, g, topo, 11, mode)
        if ctx.needs_input_grad[0]:
            d1, d2 = group_gemm_dgrad(g, w, fp, fp, topo, True, True, mode)
            dx = neighbor_sum(d2, topo, transposed=True, addend=d1)
            dx = dx[:, :ctx.f]
        return dx, dw, db, None, None, None


def _widen(x, fp):
    """[N,F] -> [N,fp] sharing storage when x is a view of a zero-padded buffer with ld >= fp."""
    n, f = x.shape
    if f == fp:
        return x
    if n > 0 and x.stride(0) >= fp and x.stride(1) == 1 and getattr(x, "_dcgc_zero_padded", False):
        return torch.as_strided(x, (n, fp), (x.stride(0), 1))
    out = torch.zeros(n, fp, device=x.device, dtype=x.dtype)
    out[:, :f] = x
    return out


class GroupLinearFn(torch.autograd.Function):
    """y = act(x . w + b) with w [K, n] (single group): the atom-level Dense layer."""

    @staticmethod
    def forward(ctx, x, w, bias, act, mode):
        _check_dev(x, "input")
        x = _rowmajor(x)
        y = group_gemm_fwd(x, None, w, bias, None, act, mode)
        ctx.act, ctx.mode = act, mode
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        g = _rowmajor(dy)
        if ctx.act == ACT_RELU:
            g = g * (y > 0)
        elif ctx.act == ACT_TANH:
            g = g * (1 - y * y)
        dw = db = dx = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = group_gemm_wgrad(x, None, g, None, 1, ctx.mode)
            dw, db = dw[0], db[0]
        if ctx.needs_input_grad[0]:
            dx, _ = group_gemm_dgrad(g, w, x.shape[1], 0, None, True, False, ctx.mode)
        return dx, (dw if ctx.needs_input_grad[1] else None), (db if ctx.needs_input_grad[2] else None), None, None


class GraphPoolFn(torch.autograd.Function):
    """P[i] = max(X[i], max_{j in N(i)} X[j]) (K3); backward gathers over CSR^T (K7).
    (The C entry point can also fold a per-channel affine; the fused model engine uses that.)"""

    @staticmethod
    def forward(ctx, x, topo):
        _check_dev(x, "atom_features")
        x = _rowmajor(x)
        n, c = x.shape
        out = padded_empty(n, c, x.device)
        ld_arg = (c + 3) // 4 * 4
        arg = torch.empty(n, ld_arg, dtype=torch.uint8, device=x.device) if x.requires_grad else None
        if n and c % 4 == 0 and _ld(x) % 4 == 0 and _al16(x, out) and mg_supported(topo, _ld(x)):
            check(_lib.lib().dcgc_mg_pool_fwd(_p(x), _ld(x), None, None, _topo_struct(topo), c, _p(out), _ld(out),
                                              _p(arg), ld_arg, _stream()))
        else:
            check(_lib.lib().dcgc_pool_fwd(_p(x), _ld(x), None, None, _p(topo.row_ptr), _p(topo.col_idx),
                                           n, c, _p(out), _ld(out), _p(arg), ld_arg, _stream()))
        _count()
        ctx.topo = topo
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        topo = ctx.topo
        dy = _rowmajor(dy)
        n, c = dy.shape
        dx = padded_empty(n, c, dy.device)
        if n and c % 4 == 0 and _ld(dy) % 4 == 0 and _al16(dy, dx, arg) and \
                mg_supported(topo, _ld(dy), arg.stride(0)):
            check(_lib.lib().dcgc_mg_pool_bwd(_p(dy), _ld(dy), _p(arg), arg.stride(0), None, _topo_struct(topo), c,
                                              _p(dx), _ld(dx), _stream()))
        else:
            check(_lib.lib().dcgc_pool_bwd(_p(dy), _ld(dy), _p(arg), arg.stride(0), None, _p(topo.t_row_ptr),
                                           _p(topo.t_src), _p(topo.t_slot), n, c, _p(dx), _ld(dx), _stream()))
        _count()
        return dx, None


class GraphGatherFn(torch.autograd.Function):
    """Z[g] = act([sum_{i in g} X[i] | max_{i in g} X[i]]) (K4) and its backward (K7)."""

    @staticmethod
    def forward(ctx, x, topo, n_segments, act):
        _check_dev(x, "atom_features")
        x = _rowmajor(x)
        n, d = x.shape
        if n_segments > topo.n_segments:
            raise ValueError("layout was built for %d segments, GraphGather asks for %d"
                             % (topo.n_segments, n_segments))
        out = torch.empty(n_segments, 2 * d, device=x.device, dtype=torch.float32)
        argrow = torch.empty(n_segments, d, device=x.device, dtype=torch.int32) if x.requires_grad else None
        check(_lib.lib().dcgc_gather_fwd(_p(x), _ld(x), None, None, _p(topo.mol_ptr), _p(topo.mol_atoms),
                                         n_segments, d, act, _p(out), 2 * d, _p(argrow), _stream()))
        _count()
        ctx.topo, ctx.act, ctx.shape = topo, act, (n, d)
        ctx.save_for_backward(out, argrow)
        return out

    @staticmethod
    def backward(ctx, dout):
        out, argrow = ctx.saved_tensors
        n, d = ctx.shape
        dout = _rowmajor(dout)
        dx = padded_empty(n, d, dout.device)
        check(_lib.lib().dcgc_gather_bwd(_p(dout), _ld(dout), _p(out), 2 * d, _p(argrow),
                                         _p(ctx.topo.membership), n, d, ctx.act, _p(dx), _ld(dx), _stream()))
        _count()
        return dx, None, None, None


def pack_graphconv_weights(W_list, b_list, fp):
    """21 reference parameters -> packed [11, 2*fp, C] / [11, C] (differentiable).

    Reference order (torch_models/layers.py:6189-6226): W[2(d-1)] neighbour weight of degree d,
    W[2(d-1)+1] self weight of degree d, W[20] degree-0 self weight.  Packed group d holds
    [self ; neighbour]; group 0's neighbour half is zero."""
    f, c = W_list[0].shape
    dev = W_list[0].device
    wst = torch.stack(list(W_list) + [torch.zeros(f, c, device=dev, dtype=W_list[0].dtype)])   # [22,F,C]
    if fp != f:
        wst = torch.nn.functional.pad(wst, (0, 0, 0, fp - f))
    idx = _pack_index(dev)
    w = wst[idx].reshape(11, 2 * fp, c)
    bst = torch.stack(list(b_list) + [torch.zeros(c, device=dev, dtype=b_list[0].dtype)])      # [22,C]
    bias = bst[idx].sum(1)
    return w.contiguous(), bias.contiguous()


_pack_idx_cache = {}


def _pack_index(device):
    key = str(device)
    if key not in _pack_idx_cache:
        rows = [[20, 21]] + [[2 * (d - 1) + 1, 2 * (d - 1)] for d in range(1, 11)]
        _pack_idx_cache[key] = torch.tensor(rows, dtype=torch.long, device=device)
    return _pack_idx_cache[key]


# ------------------------------------------------------------------------------------------
# D-MPNN path: generic CSR gather-sum, two-operand linear, per-molecule readout
# ------------------------------------------------------------------------------------------
class CsrPair(object):
    """A gather pattern and its transpose: out[i] = sum_{e in [row_ptr[i], row_ptr[i+1])} x[idx[e]]
    (n_out rows from n_in rows); (t_row_ptr, t_idx) lists, for every input row, the output rows reading it."""

    def __init__(self, row_ptr, idx, t_row_ptr, t_idx, n_out, n_in):
        self.row_ptr, self.idx, self.t_row_ptr, self.t_idx = row_ptr, idx, t_row_ptr, t_idx
        self.n_out, self.n_in = int(n_out), int(n_in)


class CsrGatherSumFn(torch.autograd.Function):
    """K8: ``message[mapping].sum(1)`` / ``h_message[atom_to_incoming_bonds].sum(1)``
    (torch_models/layers.py:1629, 1539); backward = gather over the transposed pattern (no atomics)."""

    @staticmethod
    def forward(ctx, x, csr):
        _check_dev(x, "message")
        ctx.csr = csr
        return gather_sum(_rowmajor(x), csr.row_ptr, csr.idx, csr.n_out)

    @staticmethod
    def backward(ctx, dy):
        csr = ctx.csr
        return gather_sum(_rowmajor(dy), csr.t_row_ptr, csr.t_idx, csr.n_in), None


class GroupLinear2Fn(torch.autograd.Function):
    """y = act([a1 | a2] . w + b) with w [k1+k2, n] (one group).  K9: W_i / W_h (a2 = None) and
    W_o(cat(atom_features, messages)) (layers.py:1541) without materialising the concatenation."""

    @staticmethod
    def forward(ctx, a1, a2, w, bias, act, mode, fwd_mode=None):
        _check_dev(a1, "input")
        a1 = _rowmajor(a1)
        a2 = _rowmajor(a2) if a2 is not None else None
        w = w.contiguous()
        # fwd_mode: arithmetic of the forward product alone (the D-MPNN layers pass fp16x3 where the fused engine does)
        y = group_gemm_fwd(a1, a2, w, bias, None, act, mode if fwd_mode is None else fwd_mode)
        ctx.act, ctx.mode = act, mode
        ctx.has2 = a2 is not None
        ctx.save_for_backward(a1, a2, w, y if act != ACT_NONE else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        a1, a2, w, y = ctx.saved_tensors
        g = _rowmajor(dy)
        if ctx.act == ACT_RELU:
            g = g * (y > 0)
        elif ctx.act == ACT_TANH:
            g = g * (1 - y * y)
        k1 = a1.shape[1]
        k2 = a2.shape[1] if ctx.has2 else 0
        dw = db = d1 = d2 = None
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[3]:
            dw, db = group_gemm_wgrad(a1, a2, g, None, 1, ctx.mode)
            dw, db = dw[0], db[0]
        need1, need2 = ctx.needs_input_grad[0], ctx.has2 and ctx.needs_input_grad[1]
        if need1 or need2:
            d1, d2 = group_gemm_dgrad(g, w, k1, k2, None, need1, need2, ctx.mode)
        return d1, d2, dw, (db if ctx.needs_input_grad[3] else None), None, None, None


def forward_gemm_mode(mode):
    """The arithmetic the fused engines use for FORWARD products of the tf32x3 mode: fp16 operand halves
    (csrc/gemm_tc.cu tc_gemm_kernel_v6) unless DCGC_FWD_F16X3=0 — the rule of csrc/dmpnn_model.cu forward_mode."""
    import os
    if mode == _lib.GEMM_TF32X3 and os.environ.get("DCGC_FWD_F16X3", "1")[:1] != "0":
        return _lib.GEMM_F16X3
    return mode


_READOUT_MODES = {"mean": 0, "sum": 1, "norm": 2}


class SegmentReadoutFn(torch.autograd.Function):
    """Per-molecule mean / sum / sum-over-norm of contiguous atom rows (layers.py:1550-1583)."""

    @staticmethod
    def forward(ctx, x, mol_ptr, n_mols, mode, norm):
        _check_dev(x, "atoms_hidden_states")
        x = _rowmajor(x)
        if mode not in _READOUT_MODES:
            raise Exception("Invalid aggregation")
        out = torch.empty(n_mols, x.shape[1], device=x.device, dtype=torch.float32)
        check(_lib.lib().dcgc_segment_readout_fwd(_p(x), _ld(x), _p(mol_ptr), n_mols, x.shape[1],
                                                  _READOUT_MODES[mode], float(norm), _p(out), _ld(out), _stream()))
        ctx.args = (mol_ptr, n_mols, x.shape[0], mode, float(norm))
        return out

    @staticmethod
    def backward(ctx, dout):
        mol_ptr, n_mols, n_atoms, mode, norm = ctx.args
        dout = _rowmajor(dout)
        dx = torch.empty(n_atoms, dout.shape[1], device=dout.device, dtype=torch.float32)
        check(_lib.lib().dcgc_segment_readout_bwd(_p(dout), _ld(dout), _p(mol_ptr), n_mols, n_atoms, dout.shape[1],
                                                  _READOUT_MODES[mode], norm, _p(dx), _ld(dx), _stream()))
        return dx, None, None, None, None


def dmpnn_concat_rows(atom_feat, bond_feat, bond_src, bond_edge, n_rows, ld_out=None):
    """f_ini_atoms_bonds on the device (dmpnn.py:183-188), zero rows on the pad rows; returns a [n_rows, fa+fb]
    view of a buffer whose rows are padded to a 16-byte multiple."""
    atom_feat, bond_feat = _rowmajor(atom_feat), _rowmajor(bond_feat)
    fa, fb = atom_feat.shape[1], bond_feat.shape[1]
    ld_out = ld_out or (fa + fb + 3) // 4 * 4
    out = torch.empty(n_rows, ld_out, device=atom_feat.device, dtype=torch.float32)
    check(_lib.lib().dcgc_dmpnn_concat_rows(_p(atom_feat), _ld(atom_feat), fa, _p(bond_feat), _ld(bond_feat), fb,
                                            _p(bond_src), _p(bond_edge), n_rows, _p(out), ld_out, _stream()))
    v = out[:, :fa + fb]
    v._dcgc_zero_padded = True
    return v
