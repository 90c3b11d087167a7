"""MPNN edge-network message passing on B200: ``EdgeNetwork``, ``GatedRecurrentUnit``, ``MessagePassing`` and
``SetGather`` with the reference's constructor arguments, attribute names and ``forward(inputs)`` list contracts.

Reference: deepchem/models/torch_models/layers.py:4006-4088 (EdgeNetwork), :2884-2919 (GatedRecurrentUnit),
:2976-3138 (SetGather); deepchem/models/layers.py:3648-3710 (MessagePassing, Keras only); MPNNModel
deepchem/models/graph_models.py:1045-1247 (Keras).  The torch port of these layers is forward-only (weights are plain
tensors, the LSTM step detaches); constructed as the reference constructs them, so are these (same attributes, a test
may assign ``layer.W = tensor``).  With ``trainable=True`` the weights are ``nn.Parameter`` s and every layer has a
backward pass (autograd Functions over the backward kernels of ``csrc/mpnn_kernels.cu`` and the tcgen05 dgrad / wgrad
GEMMs): the gradients of the Keras originals, which is what ``MPNNModel`` trains with.

What runs where: every contraction is a tcgen05 GEMM through ``dcgc_group_gemm_fwd`` (TF32x3: fp32-grade) and
everything else a hand-written kernel of ``csrc/mpnn_kernels.cu`` behind the C ABI; torch only owns memory and
repacks the (small) weight tensors.  EdgeNetwork does NOT materialise the reference's h x h matrix per atom pair:
the map is bilinear in (pair features, neighbour state), so it is computed as a per-destination contraction
(``dcgc_pair_contract_fwd``) followed by one dense ``[n_atoms, (P + 1) h] x [(P + 1) h, h]`` GEMM — about h / 2 times
fewer flops and no 4 h^2-byte intermediate per pair (see the header of mpnn_kernels.cu).  There is no CPU path.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, GEMM_FP32, GEMM_TF32X3
from .engine import AdamSlabState
from .ops import (GroupLinear2Fn, GroupLinearFn, _ld, _p, _rowmajor, _stream, check, group_gemm_dgrad, group_gemm_fwd,
                  group_gemm_wgrad, padded_empty)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("deepchem_b200.mpnn needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _f32(t, dev):
    """float32 CUDA tensor with contiguous rows from a numpy array / tensor on any device."""
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    t = t.detach().to(device=dev, dtype=torch.float32)
    return t if t.dim() < 2 or t.stride(-1) == 1 else t.contiguous()


def _f32g(t, dev):
    """_f32 that keeps a float32 CUDA tensor inside the autograd graph (layer inputs of a training model)."""
    if torch.is_tensor(t) and t.requires_grad and torch.is_grad_enabled() and t.is_cuda and t.dtype == torch.float32:
        return t if t.dim() < 2 or t.stride(-1) == 1 else t.contiguous()
    return _f32(t, dev)


def _init(name, shape):
    return getattr(nn.init, name)(torch.empty(*shape))


class _WeightCache(object):
    """Device copies / repackings of plain-tensor weights, rebuilt when a weight is replaced or modified in place
    (the reference's tests assign ``layer.W = ...`` after construction)."""

    def __init__(self):
        self.key, self.value = None, None

    def get(self, tensors, dev, build):
        key = tuple((id(t), t._version, t.data_ptr()) for t in tensors) + (str(dev),)
        if key != self.key:
            self.value = build([_f32(t, dev) for t in tensors])
            self.key = key
        return self.value



class _PairCsr(object):
    """Pairs grouped by destination (forward) and by source (backward), on the device."""
    __slots__ = ("ptr", "pid", "src", "n_dst", "max_src", "dst", "t_ptr", "t_pair", "n_src_rows")


class PairContractFn(torch.autograd.Function):
    """z[i] = per-destination contraction of (pair features, neighbour states), see dcgc_pair_contract_fwd; the gradient
    flows to the atom states only (the pair features are data)."""

    @staticmethod
    def forward(ctx, x, pf, csr, n_pf, h):
        x = _rowmajor(x)
        z = padded_empty(csr.n_dst, (n_pf + 1) * h, x.device)
        check(_lib.lib().dcgc_pair_contract_fwd(_p(x), _ld(x), _p(pf), _ld(pf), _p(csr.ptr), _p(csr.pid), _p(csr.src),
                                                csr.n_dst, n_pf, h, _p(z), _ld(z), _stream()))
        ctx.csr, ctx.n_pf, ctx.h, ctx.n_x = csr, n_pf, h, x.shape[0]
        ctx.save_for_backward(pf)
        return z

    @staticmethod
    def backward(ctx, dz):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        (pf,) = ctx.saved_tensors
        csr, h = ctx.csr, ctx.h
        dz = _rowmajor(dz)
        dx = torch.zeros(ctx.n_x, h, device=dz.device)          # atoms that no pair reads keep a zero gradient
        check(_lib.lib().dcgc_pair_contract_bwd_x(_p(dz), _ld(dz), _p(pf), _ld(pf), _p(csr.t_ptr), _p(csr.t_pair),
                                                  _p(csr.dst), csr.n_src_rows, ctx.n_pf, h, _p(dx), _ld(dx), _stream()))
        return dx, None, None, None, None


class GruFn(torch.autograd.Function):
    """One GatedRecurrentUnit step (two GEMMs + the gate / output kernels) with its hand-written backward."""

    @staticmethod
    def forward(ctx, hp, xm, w1, uh, bz, br, bh, mode):
        hp, xm = _rowmajor(hp), _rowmajor(xm)
        n, h = hp.shape
        dev = hp.device
        L = _lib.lib()
        g = group_gemm_fwd(xm, hp, w1, None, None, ACT_NONE, mode)            # [n, 3h]
        z, hr = padded_empty(n, h, dev), padded_empty(n, h, dev)
        check(L.dcgc_gru_gates_fwd(_p(g), _ld(g), _p(bz), _p(br), _p(hp), _ld(hp), n, h, _p(z), _ld(z), _p(hr), _ld(hr),
                                   _stream()))
        u = group_gemm_fwd(hr, None, uh, None, None, ACT_NONE, mode)
        out = padded_empty(n, h, dev)
        check(L.dcgc_gru_out_fwd(_p(g), _ld(g), _p(u), _ld(u), _p(bh), _p(z), _ld(z), _p(xm), _ld(xm), n, h, _p(out),
                                 _ld(out), _stream()))
        ctx.mode = mode
        ctx.save_for_backward(hp, xm, w1, uh, bz, br, bh, g, z, hr, u)
        return out

    @staticmethod
    def backward(ctx, dout):
        hp, xm, w1, uh, bz, br, bh, g, z, hr, u = ctx.saved_tensors
        n, h = hp.shape
        dev, mode = hp.device, ctx.mode
        L = _lib.lib()
        dout = _rowmajor(dout)
        dzg, dx1, dpre = padded_empty(n, h, dev), padded_empty(n, h, dev), padded_empty(n, h, dev)
        check(L.dcgc_gru_out_bwd(_p(dout), _ld(dout), _p(g), _ld(g), _p(u), _ld(u), _p(bh), _p(z), _ld(z), _p(xm), _ld(xm),
                                 n, h, _p(dzg), _ld(dzg), _p(dx1), _ld(dx1), _p(dpre), _ld(dpre), _stream()))
        duh, dbh = group_gemm_wgrad(hr, None, dpre, None, 1, mode)           # u = hr . Uh; column sums of dpre = d bh
        dhr, _ = group_gemm_dgrad(dpre, uh, h, 0, None, True, False, mode)
        dg, dh1 = padded_empty(n, 3 * h, dev), padded_empty(n, h, dev)
        check(L.dcgc_gru_gates_bwd(_p(dzg), _ld(dzg), _p(dhr), _ld(dhr), _p(dpre), _ld(dpre), _p(g), _ld(g), _p(bz), _p(br),
                                   _p(hp), _ld(hp), n, h, _p(dg), _ld(dg), _p(dh1), _ld(dh1), _stream()))
        dw1, dbg = group_gemm_wgrad(xm, hp, dg, None, 1, mode)               # g = [x | h_prev] . w1
        dx2, dh2 = group_gemm_dgrad(dg, w1, h, h, None, True, True, mode)
        return (dh1 + dh2, dx1 + dx2, dw1[0], duh[0], dbg[0][:h].contiguous(), dbg[0][h:2 * h].contiguous(), dbh[0],
                None)


class AttendFn(torch.autograd.Function):
    """One set2set attention read (dcgc_setgather_attend_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, q, ptr, atoms, n_mols, max_atoms):
        x, q = _rowmajor(x), _rowmajor(q)
        h = x.shape[1]
        qs = torch.empty(n_mols, 2 * h, device=x.device)
        check(_lib.lib().dcgc_setgather_attend_fwd(_p(x), _ld(x), _p(q), _ld(q), _p(ptr), _p(atoms), n_mols, h, max_atoms,
                                                   _p(qs), _ld(qs), _stream()))
        ctx.n_mols, ctx.max_atoms = n_mols, max_atoms
        ctx.save_for_backward(x, q, ptr, atoms)
        return qs

    @staticmethod
    def backward(ctx, dqs):
        x, q, ptr, atoms = ctx.saved_tensors
        dqs = _rowmajor(dqs)
        h = x.shape[1]
        dx = torch.zeros_like(x)                                  # atoms outside every molecule keep zero
        dq = torch.empty(ctx.n_mols, h, device=x.device)
        check(_lib.lib().dcgc_setgather_attend_bwd(_p(x), _ld(x), _p(q), _ld(q), _p(dqs), _ld(dqs), _p(ptr), _p(atoms),
                                                   ctx.n_mols, h, ctx.max_atoms, _p(dx), _ld(dx), _p(dq), _ld(dq),
                                                   _stream()))
        return dx, dq, None, None, None, None


class LstmStepFn(torch.autograd.Function):
    """The LSTM cell of set2set on pre-activations z = q_star . U + b (dcgc_lstm_step_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, z, c):
        z, c = _rowmajor(z), c.contiguous()
        n, h = c.shape
        h_new, c_new = torch.empty(n, h, device=z.device), torch.empty(n, h, device=z.device)
        check(_lib.lib().dcgc_lstm_step_fwd(_p(z), _ld(z), _p(c), n, h, _p(h_new), _p(c_new), _stream()))
        ctx.save_for_backward(z, c)
        return h_new, c_new

    @staticmethod
    def backward(ctx, dh, dc):
        z, c = ctx.saved_tensors
        n, h = c.shape
        dh = dh.contiguous() if dh is not None else None
        dc = dc.contiguous() if dc is not None else None
        dz = padded_empty(n, 4 * h, z.device)
        dc_in = torch.empty(n, h, device=z.device)
        check(_lib.lib().dcgc_lstm_step_bwd(_p(z), _ld(z), _p(c), _p(dh) if dh is not None else None,
                                            _p(dc) if dc is not None else None, n, h, _p(dz), _ld(dz), _p(dc_in), _stream()))
        return dz, dc_in


def _linear(x, w, b, act, mode):
    """act(x . w + b), differentiable when anything requires a gradient."""
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or (b is not None and b.requires_grad)):
        return GroupLinearFn.apply(x, w, b, act, mode)
    return group_gemm_fwd(x, None, w, b, None, act, mode)


class EdgeNetwork(nn.Module):
    """layers.py:4006-4088.  ``forward([pair_features [Np, P], atom_features [Na, h], atom_to_pair [Np, 2]])`` ->
    ``[n_dst, h]`` with ``out[i] = sum_{p: atom_to_pair[p,0] == i} reshape(pair_features[p] . W + b, [h, h]) .
    atom_features[atom_to_pair[p,1]]``; ``atom_to_pair[:, 0]`` must be sorted (the reference's segment_sum)."""

    def __init__(self, n_pair_features=8, n_hidden=100, init='xavier_uniform_', gemm_mode=GEMM_TF32X3, trainable=False,
                 **kwargs):
        super(EdgeNetwork, self).__init__(**kwargs)
        self.n_pair_features, self.n_hidden, self.init = n_pair_features, n_hidden, init
        self.W = _init(init, (n_pair_features, n_hidden * n_hidden))
        self.b = torch.zeros((n_hidden * n_hidden,))
        self.trainable = bool(trainable)
        if self.trainable:          # the Keras original's trainable weights (models/layers.py:3712-3753)
            self.W, self.b = nn.Parameter(self.W), nn.Parameter(self.b)
        self.built = True
        self.gemm_mode = gemm_mode
        self._w = _WeightCache()
        self._pairs = (None, None)

    K_SLICE = 320       # longest single tensor-core contraction (columns of z), see forward

    def __repr__(self):
        return '%s(n_pair_features:%s,n_hidden:%s,init:%s)' % (self.__class__.__name__, self.n_pair_features,
                                                               self.n_hidden, self.init)

    def _w_ext(self, dev):
        P, h = self.n_pair_features, self.n_hidden

        def build(ts):
            W, b = ts
            # W_ext[f*h + b, a] = W[f, a*h + b];  W_ext[P*h + b, a] = bias[a*h + b]
            return torch.cat([W.view(P, h, h).permute(0, 2, 1).reshape(P * h, h), b.view(h, h).t()], 0).contiguous()
        if self.trainable and torch.is_grad_enabled():
            return build((self.W, self.b))            # inside the autograd graph: the gradient reaches W and b
        return self._w.get((self.W, self.b), dev, build)

    def _pair_csr(self, atom_to_pair, dev):
        """Pairs grouped by destination (host, cached per index tensor): (pair_ptr, pair_id, pair_src, n_dst)."""
        # cached per index OBJECT (MessagePassing calls the layer T times with the same one); the cache keeps a reference
        # to it, so its id cannot be handed to another array while the entry lives
        key = (atom_to_pair, getattr(atom_to_pair, "_version", 0))
        if self._pairs[0] is not None and self._pairs[0][0] is atom_to_pair and self._pairs[0][1] == key[1]:
            return self._pairs[1]
        a2p = atom_to_pair.detach().cpu().numpy() if torch.is_tensor(atom_to_pair) else np.asarray(atom_to_pair)
        if a2p.ndim != 2 or a2p.shape[1] != 2:
            raise ValueError("atom_to_pair must be [n_pairs, 2]")
        dst, src = a2p[:, 0].astype(np.int64), a2p[:, 1].astype(np.int64)
        if dst.size and not bool(np.all(dst[1:] >= dst[:-1])):
            raise AssertionError("elements of segment_ids must be sorted")     # pytorch_utils.py:109-111
        n_dst = int(dst[-1]) + 1 if dst.size else 0
        if dst.size and (dst[0] < 0 or len(np.unique(dst)) != n_dst):
            raise ValueError("atom_to_pair[:, 0] must cover 0 .. n-1 without gaps")   # the reference indexes out of range
        ptr = np.searchsorted(dst, np.arange(n_dst + 1)).astype(np.int32)
        csr = _PairCsr()
        csr.ptr = torch.from_numpy(ptr).to(dev)
        csr.pid = torch.arange(dst.size, dtype=torch.int32, device=dev)
        csr.src = torch.from_numpy(src.astype(np.int32)).to(dev)
        csr.n_dst, csr.max_src = n_dst, (int(src.max()) if src.size else -1)
        # the same pairs grouped by SOURCE atom, pair ids ascending inside a group (the backward pass gathers over them)
        csr.dst = torch.from_numpy(dst.astype(np.int32)).to(dev)
        order = np.argsort(src, kind="stable").astype(np.int32)
        csr.n_src_rows = csr.max_src + 1
        csr.t_ptr = torch.from_numpy(np.searchsorted(src[order], np.arange(csr.n_src_rows + 1)).astype(np.int32)).to(dev)
        csr.t_pair = torch.from_numpy(order).to(dev)
        self._pairs = (key, csr)
        return csr

    def forward(self, inputs):
        pair_features, atom_features, atom_to_pair = inputs
        dev = _device()
        P, h = self.n_pair_features, self.n_hidden
        pf, x = _f32(pair_features, dev), _f32g(atom_features, dev)
        if pf.dim() != 2 or pf.shape[1] != P or x.dim() != 2 or x.shape[1] != h:
            raise ValueError("EdgeNetwork: pair_features must be [n_pairs, %d] and atom_features [n_atoms, %d]" % (P, h))
        csr = self._pair_csr(atom_to_pair, dev)
        if csr.pid.shape[0] != pf.shape[0] or csr.max_src >= x.shape[0]:
            raise IndexError("EdgeNetwork: atom_to_pair does not match pair_features / atom_features")
        grad = torch.is_grad_enabled() and (x.requires_grad or self.trainable)
        w_ext = self._w_ext(dev)
        if grad:
            z = PairContractFn.apply(x, pf, csr, P, h)
            if self.gemm_mode == GEMM_FP32 or (P + 1) * h <= self.K_SLICE:
                return _linear(z, w_ext, None, ACT_NONE, self.gemm_mode)
            fc = max(1, self.K_SLICE // h) * h            # (see below)
            out = None
            for k0 in range(0, (P + 1) * h, fc):
                k1 = min((P + 1) * h, k0 + fc)
                part = _linear(z[:, k0:k1], w_ext[k0:k1], None, ACT_NONE, self.gemm_mode)
                out = part if out is None else out + part
            return out
        z = padded_empty(csr.n_dst, (P + 1) * h, dev)
        check(_lib.lib().dcgc_pair_contract_fwd(_p(x), _ld(x), _p(pf), _ld(pf), _p(csr.ptr), _p(csr.pid), _p(csr.src),
                                                csr.n_dst, P, h, _p(z), _ld(z), _stream()))
        if self.gemm_mode == GEMM_FP32 or (P + 1) * h <= self.K_SLICE:
            return group_gemm_fwd(z, None, w_ext, None, None, ACT_NONE, self.gemm_mode)
        # The tensor core adds every MMA into its fp32 accumulator rounding toward zero, so the error of one long
        # contraction grows with its length (measured 1.1e-5 of the output scale at K = 1500).  Slices of whole
        # pair-feature blocks (<= K_SLICE columns of z) are contracted separately and their results added in
        # fp32 round-to-nearest, in slice order: the bias of each slice is K_SLICE / K of that, with random signs.
        fc = max(1, self.K_SLICE // h) * h
        out = None
        for k0 in range(0, (P + 1) * h, fc):
            k1 = min((P + 1) * h, k0 + fc)
            part = group_gemm_fwd(z[:, k0:k1], None, w_ext[k0:k1], None, None, ACT_NONE, self.gemm_mode)
            out = part if out is None else out.add_(part)
        return out


class GatedRecurrentUnit(nn.Module):
    """layers.py:2884-2919.  ``forward([h_tm1, x])``."""

    def __init__(self, n_hidden=100, init='xavier_uniform_', gemm_mode=GEMM_TF32X3, trainable=False, **kwargs):
        super(GatedRecurrentUnit, self).__init__(**kwargs)
        self.n_hidden, self.init = n_hidden, init
        self.trainable = bool(trainable)
        wrap = nn.Parameter if self.trainable else (lambda t: t)    # Keras original: trainable (models/layers.py:3755-3800)
        for name in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh"):
            setattr(self, name, wrap(_init(init, (n_hidden, n_hidden))))
        for name in ("bz", "br", "bh"):
            setattr(self, name, wrap(torch.zeros((n_hidden,))))
        self.gemm_mode = gemm_mode
        self._w = _WeightCache()

    def _packed(self, dev):
        def build(ts):
            Wz, Wr, Wh, Uz, Ur, Uh, bz, br, bh = ts
            top = torch.cat([Wz, Wr, Wh], 1)                                  # x . [Wz Wr Wh]
            bot = torch.cat([Uz, Ur, torch.zeros_like(Uh)], 1)                # h . [Uz Ur 0]
            return torch.cat([top, bot], 0).contiguous(), Uh.contiguous(), bz.contiguous(), br.contiguous(), bh.contiguous()
        ts = tuple(getattr(self, n) for n in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh"))
        if self.trainable and torch.is_grad_enabled():
            return build(ts)                          # inside the autograd graph
        return self._w.get(ts, dev, build)

    def forward(self, inputs):
        h_tm1, x = inputs
        dev = _device()
        h = self.n_hidden
        hp, xm = _f32g(h_tm1, dev), _f32g(x, dev)
        if hp.shape != xm.shape or hp.dim() != 2 or hp.shape[1] != h:
            raise ValueError("GatedRecurrentUnit: both inputs must be [n, %d]" % h)
        n = hp.shape[0]
        w1, uh, bz, br, bh = self._packed(dev)
        if torch.is_grad_enabled() and (self.trainable or hp.requires_grad or xm.requires_grad):
            return GruFn.apply(hp, xm, w1, uh, bz, br, bh, self.gemm_mode)
        g = group_gemm_fwd(xm, hp, w1, None, None, ACT_NONE, self.gemm_mode)            # [n, 3h]
        z, hr = padded_empty(n, h, dev), padded_empty(n, h, dev)
        L = _lib.lib()
        check(L.dcgc_gru_gates_fwd(_p(g), _ld(g), _p(bz), _p(br), _p(hp), _ld(hp), n, h, _p(z), _ld(z), _p(hr), _ld(hr),
                                   _stream()))
        u = group_gemm_fwd(hr, None, uh, None, None, ACT_NONE, self.gemm_mode)
        out = padded_empty(n, h, dev)
        check(L.dcgc_gru_out_fwd(_p(g), _ld(g), _p(u), _ld(u), _p(bh), _p(z), _ld(z), _p(xm), _ld(xm), n, h, _p(out),
                                 _ld(out), _stream()))
        return out


class MessagePassing(nn.Module):
    """models/layers.py:3648-3710: pad the atom features to n_hidden, then T x (EdgeNetwork message, GRU update).
    ``forward([atom_features, pair_features, atom_to_pair])``.  The sub-layers are built on first use from the width
    of the pair features (Keras ``build``)."""

    def __init__(self, T, message_fn='enn', update_fn='gru', n_hidden=100, gemm_mode=GEMM_TF32X3, trainable=False,
                 **kwargs):
        super(MessagePassing, self).__init__(**kwargs)
        self.T, self.message_fn, self.update_fn, self.n_hidden = T, message_fn, update_fn, n_hidden
        self.gemm_mode, self.trainable = gemm_mode, bool(trainable)
        self.message_function = None
        self.update_function = None
        self.built = False

    def build(self, n_pair_features):
        if self.message_fn == 'enn':
            self.message_function = EdgeNetwork(n_pair_features, self.n_hidden, gemm_mode=self.gemm_mode,
                                                trainable=self.trainable)
        if self.update_fn == 'gru':
            self.update_function = GatedRecurrentUnit(self.n_hidden, gemm_mode=self.gemm_mode, trainable=self.trainable)
        self.built = True

    def forward(self, inputs):
        atom_features, pair_features, atom_to_pair = inputs
        dev = _device()
        x, pf = _f32(atom_features, dev), _f32(pair_features, dev)
        if not self.built:
            self.build(int(pf.shape[-1]))
        n_feat = x.shape[-1]
        if n_feat > self.n_hidden:
            raise ValueError("Too large initial feature vector")
        out = x
        if n_feat < self.n_hidden:
            out = torch.zeros(x.shape[0], self.n_hidden, device=dev)
            out[:, :n_feat] = x
        for _ in range(self.T):
            message = self.message_function([pf, out, atom_to_pair])
            out = self.update_function([out, message])
        return out


class SetGather(nn.Module):
    """layers.py:2976-3138: M steps of set2set.  ``forward([atom_features [N, n_hidden], atom_split [N]])`` ->
    ``q_star [batch_size, 2 n_hidden]`` (float32 on the device; the reference returns float64 for float64 input)."""

    def __init__(self, M, batch_size, n_hidden=100, init='orthogonal', gemm_mode=GEMM_TF32X3, trainable=False, **kwargs):
        super(SetGather, self).__init__(**kwargs)
        self.M, self.batch_size, self.n_hidden, self.init = M, batch_size, n_hidden, init
        self.trainable = bool(trainable)     # False: the torch port (its LSTM step detaches); True: the Keras gradient
        self.U = nn.Parameter(torch.Tensor(2 * n_hidden, 4 * n_hidden).normal_(mean=0.0, std=0.1))
        self.b = nn.Parameter(torch.cat((torch.zeros(n_hidden), torch.ones(n_hidden), torch.zeros(n_hidden),
                                         torch.zeros(n_hidden))))
        self.built = True
        self.gemm_mode = gemm_mode
        self._w = _WeightCache()

    def __repr__(self):
        return '%s(M=%s, batch_size=%s, n_hidden=%s, init=%s)' % (self.__class__.__name__, self.M, self.batch_size,
                                                                  self.n_hidden, self.init)

    def forward(self, inputs):
        atom_features, atom_split = inputs
        dev = _device()
        h, B = self.n_hidden, self.batch_size
        x = _f32g(atom_features, dev) if self.trainable else _f32(atom_features, dev)
        split = atom_split.detach().cpu().numpy() if torch.is_tensor(atom_split) else np.asarray(atom_split)
        split = split.astype(np.int64).reshape(-1)
        if x.dim() != 2 or x.shape[1] != h or split.shape[0] != x.shape[0]:
            raise ValueError("SetGather: atom_features must be [N, %d] and atom_split [N]" % h)
        if split.size and (split.min() < 0 or split.max() >= B):
            raise IndexError("SetGather: atom_split outside [0, batch_size)")
        # molecule -> atoms in ascending order (stable counting sort on the host)
        counts = np.bincount(split, minlength=B)
        mol_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        mol_atoms = np.argsort(split, kind="stable").astype(np.int32)
        ptr_d, atoms_d = torch.from_numpy(mol_ptr).to(dev), torch.from_numpy(mol_atoms).to(dev)
        max_atoms = int(counts.max()) if counts.size else 0
        if self.trainable and torch.is_grad_enabled():
            # models/layers.py:3802-3887 with its gradient: attention read, q_star . U + b, LSTM cell, M times
            Up, bp = self.U, self.b
            if Up.device != dev:
                raise RuntimeError("SetGather(trainable=True): move the module to the CUDA device first")
            c = torch.zeros(B, h, device=dev)
            hq = torch.zeros(B, h, device=dev)
            q_star = torch.zeros(B, 2 * h, device=dev)
            for _ in range(self.M):
                q_star = AttendFn.apply(x, hq, ptr_d, atoms_d, B, max_atoms)
                z = _linear(q_star, Up, bp, ACT_NONE, self.gemm_mode)
                hq, c = LstmStepFn.apply(z, c)
            return q_star
        U, b = self._w.get((self.U, self.b), dev, lambda ts: (ts[0].contiguous(), ts[1].contiguous()))
        L = _lib.lib()
        c = torch.zeros(B, h, device=dev)
        hq = torch.zeros(B, h, device=dev)
        q_star = torch.zeros(B, 2 * h, device=dev)
        for _ in range(self.M):
            q_star = torch.empty(B, 2 * h, device=dev)
            check(L.dcgc_setgather_attend_fwd(_p(x), _ld(x), _p(hq), _ld(hq), _p(ptr_d), _p(atoms_d), B, h, max_atoms,
                                              _p(q_star), _ld(q_star), _stream()))
            z = group_gemm_fwd(q_star, None, U, b, None, ACT_NONE, self.gemm_mode)       # [B, 4h]
            h_new, c_new = torch.empty(B, h, device=dev), torch.empty(B, h, device=dev)
            check(L.dcgc_lstm_step_fwd(_p(z), _ld(z), _p(c), B, h, _p(h_new), _p(c_new), _stream()))
            hq, c = h_new, c_new
        return q_star


# ------------------------------------------------------------------------------------------------------------------
# MPNNModel (deepchem/models/graph_models.py:1045-1247, Keras): MessagePassing(T) -> Dense(n_hidden) -> SetGather(M)
# -> Dense(2 n_hidden, relu) -> Dense(n_tasks [* n_classes]), trained with Adam on L2 / softmax cross-entropy.
# ------------------------------------------------------------------------------------------------------------------
def _glorot(fan_in, fan_out):
    return nn.Parameter(nn.init.xavier_uniform_(torch.empty(fan_in, fan_out)))          # Keras Dense default


class _MPNNTorchModel(nn.Module):
    """The network of MPNNModel.  Dense kernels are stored [in, out] as Keras stores them."""

    def __init__(self, n_tasks, n_atom_feat, n_pair_feat, n_hidden, T, M, mode, n_classes, batch_size, gemm_mode):
        super(_MPNNTorchModel, self).__init__()
        self.n_tasks, self.n_classes, self.mode, self.n_hidden = n_tasks, n_classes, mode, n_hidden
        self.gemm_mode = gemm_mode
        self.message_passing = MessagePassing(T, message_fn='enn', update_fn='gru', n_hidden=n_hidden,
                                              gemm_mode=gemm_mode, trainable=True)
        self.message_passing.build(n_pair_feat)
        self.atom_dense_kernel, self.atom_dense_bias = _glorot(n_hidden, n_hidden), nn.Parameter(torch.zeros(n_hidden))
        self.set_gather = SetGather(M, batch_size, n_hidden=n_hidden, gemm_mode=gemm_mode, trainable=True)
        self.dense1_kernel, self.dense1_bias = _glorot(2 * n_hidden, 2 * n_hidden), nn.Parameter(torch.zeros(2 * n_hidden))
        n_out = n_tasks * n_classes if mode == 'classification' else n_tasks
        self.out_kernel, self.out_bias = _glorot(2 * n_hidden, n_out), nn.Parameter(torch.zeros(n_out))

    def forward(self, inputs):
        atom_features, pair_features, atom_split, atom_to_pair, n_samples = inputs
        n_samples = int(n_samples)
        mode = self.gemm_mode
        h = self.message_passing([atom_features, pair_features, atom_to_pair])
        emb = _linear(h, self.atom_dense_kernel, self.atom_dense_bias, ACT_NONE, mode)
        mol = self.set_gather([emb, atom_split])
        d1 = _linear(mol, self.dense1_kernel, self.dense1_bias, ACT_RELU, mode)
        out = _linear(d1, self.out_kernel, self.out_bias, ACT_NONE, mode)
        if self.mode == 'classification':
            logits = out.reshape(-1, self.n_tasks, self.n_classes)[:n_samples]
            return [torch.softmax(logits, dim=2), logits]
        return [out[:n_samples]]


def weave_batch_inputs(X_b, n_pair_feat):
    """The collation of graph_models.py:1214-1246: atom features, pair features (all n x n ordered pairs of every
    molecule, row-major), atom_split (molecule of every atom) and atom_to_pair ([destination, source] atom of every
    pair, destinations ascending) of a batch of Weave-style molecules."""
    atom_feat, pair_feat, atom_split, atom_to_pair = [], [], [], []
    start = 0
    for im, mol in enumerate(X_b):
        n_atoms = mol.get_num_atoms()
        atom_split.extend([im] * n_atoms)
        C0, C1 = np.meshgrid(np.arange(n_atoms), np.arange(n_atoms))
        atom_to_pair.append(np.transpose(np.array([C1.flatten() + start, C0.flatten() + start])))
        start = start + n_atoms
        atom_feat.append(mol.get_atom_features())
        pair_feat.append(np.reshape(mol.get_pair_features(), (n_atoms * n_atoms, n_pair_feat)))
    return [np.concatenate(atom_feat, axis=0), np.concatenate(pair_feat, axis=0), np.array(atom_split),
            np.concatenate(atom_to_pair, axis=0)]


class MPNNModel(AdamSlabState):
    """graph_models.py:1045-1247.  ``fit`` / ``predict`` over a dataset whose X holds Weave-style molecules (objects
    with ``get_num_atoms()``, ``get_atom_features()`` [n, n_atom_feat] and ``get_pair_features()`` [n, n, n_pair_feat]
    — ``feat.mol_graphs.WeaveMol`` in the reference); ``default_generator`` yields the reference's five input arrays.
    ``dropout`` is accepted and unused exactly as in the reference network; ``uncertainty`` needs it and is not
    supported here."""

    def __init__(self, n_tasks, n_atom_feat=70, n_pair_feat=8, n_hidden=100, T=5, M=10, mode="regression", dropout=0.0,
                 n_classes=2, uncertainty=False, batch_size=100, learning_rate=1e-3, device=None,
                 gemm_mode=GEMM_TF32X3, **kwargs):
        if mode not in ['classification', 'regression']:
            raise ValueError("mode must be either 'classification' or 'regression'")
        if uncertainty:
            if mode != "regression":
                raise ValueError("Uncertainty is only supported in regression mode")
            if dropout == 0.0:
                raise ValueError('Dropout must be included to predict uncertainty')
            raise NotImplementedError("MPNNModel(uncertainty=True) is not available on this path")
        self.n_tasks, self.n_atom_feat, self.n_pair_feat, self.n_hidden = n_tasks, n_atom_feat, n_pair_feat, n_hidden
        self.T, self.M, self.mode, self.n_classes, self.batch_size = T, M, mode, n_classes, batch_size
        self.device = torch.device(device) if device is not None else _device()
        self.model = _MPNNTorchModel(n_tasks, n_atom_feat, n_pair_feat, n_hidden, T, M, mode, n_classes, batch_size,
                                     gemm_mode).to(self.device)
        self.lr, self.betas, self.eps = float(learning_rate), (0.9, 0.999), 1e-8
        self.step_count = 0
        self._global_step = 0
        self._build_slab()

    # ---- every parameter a view of one slab: one fused Adam launch (dcgc_adam_step), one memset for the gradients
    def _build_slab(self):
        ps = list(self.model.parameters())
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        dev = self.device
        self.params = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self._slots = []
        with torch.no_grad():
            for p, off in zip(ps, offs):
                view = self.params[off:off + p.numel()].view(p.shape)
                gview = self.grads[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = gview
                self._slots.append((p, view, gview))

    def default_generator(self, dataset, epochs=1, mode='fit', deterministic=True, pad_batches=True):
        """graph_models.py:1196-1247 (same arrays in the same order)."""
        from .data import pad_features
        for _ in range(epochs):
            for (X_b, y_b, w_b, ids_b) in dataset.iterbatches(batch_size=self.batch_size, deterministic=deterministic,
                                                              pad_batches=pad_batches):
                n_samples = np.array(X_b.shape[0])
                X_b = pad_features(self.batch_size, X_b)
                if y_b is not None and self.mode == 'classification':
                    y_b = np.eye(self.n_classes)[np.asarray(y_b).flatten().astype(int)].reshape(-1, self.n_tasks,
                                                                                              self.n_classes)
                yield (weave_batch_inputs(X_b, self.n_pair_feat) + [n_samples], [y_b], [w_b])

    def _loss(self, outputs, y, w):
        """L2Loss / SoftmaxCrossEntropy under KerasModel's weighted mean (models/losses.py:76-94, 236-259)."""
        dev = self.device
        y = torch.as_tensor(np.asarray(y), dtype=torch.float32, device=dev)
        w = torch.as_tensor(np.asarray(w), dtype=torch.float32, device=dev)
        if self.mode == 'classification':
            per = -(y * torch.log_softmax(outputs[1], dim=-1)).sum(-1)
        else:
            per = (outputs[0] - y.reshape(outputs[0].shape)) ** 2
        while w.dim() < per.dim():
            w = w.unsqueeze(-1)
        return (per * w).mean()

    def _to_inputs(self, inputs):
        dev = self.device
        af, pf, split, a2p, n_samples = inputs
        return [torch.as_tensor(np.asarray(af), dtype=torch.float32, device=dev),
                torch.as_tensor(np.asarray(pf), dtype=torch.float32, device=dev), np.asarray(split), np.asarray(a2p),
                int(n_samples)]

    def fit_generator(self, generator, **kwargs):
        self.model.train()
        last, total, n = 0.0, 0.0, 0
        L = _lib.lib()
        for inputs, labels, weights in generator:
            ins = self._to_inputs(inputs)
            k = ins[4]
            self.grads.zero_()
            outputs = self.model(ins)
            loss = self._loss(outputs, np.asarray(labels[0])[:k], np.asarray(weights[0])[:k])
            loss.backward()
            self.step_count += 1
            self._global_step += 1
            check(L.dcgc_adam_step(_p(self.params), _p(self.grads), _p(self.exp_avg), _p(self.exp_avg_sq),
                                   self.params.numel(), ctypes.c_float(self.lr), ctypes.c_float(self.betas[0]),
                                   ctypes.c_float(self.betas[1]), ctypes.c_float(self.eps), self.step_count,
                                   ctypes.c_float(1.0), _stream()))
            last = float(loss.detach())
            total += last
            n += 1
        return total / n if n else 0.0

    def fit(self, dataset, nb_epoch=10, deterministic=False, **kwargs):
        return self.fit_generator(self.default_generator(dataset, epochs=nb_epoch, deterministic=deterministic))

    def predict_on_generator(self, generator, **kwargs):
        self.model.eval()
        outs = []
        with torch.no_grad():
            for inputs, _, _ in generator:
                outs.append(self.model(self._to_inputs(inputs))[0].cpu().numpy())
        return np.concatenate(outs, axis=0) if outs else np.zeros((0,))

    def predict(self, dataset, **kwargs):
        return self.predict_on_generator(self.default_generator(dataset, mode='predict', deterministic=True,
                                                                pad_batches=False))
