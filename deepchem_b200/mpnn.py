"""MPNN edge-network message passing on B200: ``EdgeNetwork``, ``GatedRecurrentUnit``, ``MessagePassing`` and
``SetGather`` with the reference's constructor arguments, attribute names and ``forward(inputs)`` list contracts.

Reference: deepchem/models/torch_models/layers.py:4006-4088 (EdgeNetwork), :2884-2919 (GatedRecurrentUnit),
:2976-3138 (SetGather); deepchem/models/layers.py:3648-3710 (MessagePassing, Keras only).  The torch port of these
layers is forward-only (weights are plain tensors, the LSTM step detaches), and so are these.

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
from ._lib import ACT_NONE, GEMM_FP32, GEMM_TF32X3
from .ops import _ld, _p, _stream, check, group_gemm_fwd, padded_empty


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


class EdgeNetwork(nn.Module):
    """layers.py:4006-4088.  ``forward([pair_features [Np, P], atom_features [Na, h], atom_to_pair [Np, 2]])`` ->
    ``[n_dst, h]`` with ``out[i] = sum_{p: atom_to_pair[p,0] == i} reshape(pair_features[p] . W + b, [h, h]) .
    atom_features[atom_to_pair[p,1]]``; ``atom_to_pair[:, 0]`` must be sorted (the reference's segment_sum)."""

    def __init__(self, n_pair_features=8, n_hidden=100, init='xavier_uniform_', gemm_mode=GEMM_TF32X3, **kwargs):
        super(EdgeNetwork, self).__init__(**kwargs)
        self.n_pair_features, self.n_hidden, self.init = n_pair_features, n_hidden, init
        self.W = _init(init, (n_pair_features, n_hidden * n_hidden))
        self.b = torch.zeros((n_hidden * n_hidden,))
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
        return self._w.get((self.W, self.b), dev, build)

    def _pair_csr(self, atom_to_pair, dev):
        """Pairs grouped by destination (host, cached per index tensor): (pair_ptr, pair_id, pair_src, n_dst)."""
        key = (id(atom_to_pair), getattr(atom_to_pair, "_version", 0))
        if self._pairs[0] == key:
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
        csr = (torch.from_numpy(ptr).to(dev), torch.arange(dst.size, dtype=torch.int32, device=dev),
               torch.from_numpy(src.astype(np.int32)).to(dev), n_dst, int(src.max()) if src.size else -1)
        self._pairs = (key, csr)
        return csr

    def forward(self, inputs):
        pair_features, atom_features, atom_to_pair = inputs
        dev = _device()
        P, h = self.n_pair_features, self.n_hidden
        pf, x = _f32(pair_features, dev), _f32(atom_features, dev)
        if pf.dim() != 2 or pf.shape[1] != P or x.dim() != 2 or x.shape[1] != h:
            raise ValueError("EdgeNetwork: pair_features must be [n_pairs, %d] and atom_features [n_atoms, %d]" % (P, h))
        ptr, pid, src, n_dst, max_src = self._pair_csr(atom_to_pair, dev)
        if pid.shape[0] != pf.shape[0] or max_src >= x.shape[0]:
            raise IndexError("EdgeNetwork: atom_to_pair does not match pair_features / atom_features")
        z = padded_empty(n_dst, (P + 1) * h, dev)
        check(_lib.lib().dcgc_pair_contract_fwd(_p(x), _ld(x), _p(pf), _ld(pf), _p(ptr), _p(pid), _p(src), n_dst, P, h,
                                                _p(z), _ld(z), _stream()))
        w_ext = self._w_ext(dev)
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

    def __init__(self, n_hidden=100, init='xavier_uniform_', gemm_mode=GEMM_TF32X3, **kwargs):
        super(GatedRecurrentUnit, self).__init__(**kwargs)
        self.n_hidden, self.init = n_hidden, init
        for name in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh"):
            setattr(self, name, _init(init, (n_hidden, n_hidden)))
        for name in ("bz", "br", "bh"):
            setattr(self, name, torch.zeros((n_hidden,)))
        self.gemm_mode = gemm_mode
        self._w = _WeightCache()

    def _packed(self, dev):
        def build(ts):
            Wz, Wr, Wh, Uz, Ur, Uh, bz, br, bh = ts
            top = torch.cat([Wz, Wr, Wh], 1)                                  # x . [Wz Wr Wh]
            bot = torch.cat([Uz, Ur, torch.zeros_like(Uh)], 1)                # h . [Uz Ur 0]
            return torch.cat([top, bot], 0).contiguous(), Uh.contiguous(), bz.contiguous(), br.contiguous(), bh.contiguous()
        return self._w.get(tuple(getattr(self, n) for n in ("Wz", "Wr", "Wh", "Uz", "Ur", "Uh", "bz", "br", "bh")),
                           dev, build)

    def forward(self, inputs):
        h_tm1, x = inputs
        dev = _device()
        h = self.n_hidden
        hp, xm = _f32(h_tm1, dev), _f32(x, dev)
        if hp.shape != xm.shape or hp.dim() != 2 or hp.shape[1] != h:
            raise ValueError("GatedRecurrentUnit: both inputs must be [n, %d]" % h)
        n = hp.shape[0]
        w1, uh, bz, br, bh = self._packed(dev)
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

    def __init__(self, T, message_fn='enn', update_fn='gru', n_hidden=100, gemm_mode=GEMM_TF32X3, **kwargs):
        super(MessagePassing, self).__init__(**kwargs)
        self.T, self.message_fn, self.update_fn, self.n_hidden = T, message_fn, update_fn, n_hidden
        self.gemm_mode = gemm_mode
        self.message_function = None
        self.update_function = None
        self.built = False

    def build(self, n_pair_features):
        if self.message_fn == 'enn':
            self.message_function = EdgeNetwork(n_pair_features, self.n_hidden, gemm_mode=self.gemm_mode)
        if self.update_fn == 'gru':
            self.update_function = GatedRecurrentUnit(self.n_hidden, gemm_mode=self.gemm_mode)
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

    def __init__(self, M, batch_size, n_hidden=100, init='orthogonal', gemm_mode=GEMM_TF32X3, **kwargs):
        super(SetGather, self).__init__(**kwargs)
        self.M, self.batch_size, self.n_hidden, self.init = M, batch_size, n_hidden, init
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
        x = _f32(atom_features, dev)
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
