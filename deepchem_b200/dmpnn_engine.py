"""Flat-slab D-MPNN training engine binding (dcgc_dmpnn_model_* in include/dcgc.h): the D-MPNN counterpart of
``engine.FlatEngine``.  Every parameter of a ``DMPNN`` module (``encoder.W_i/W_h/W_o``, ``ffn.linears.N``) becomes a
view of one contiguous fp32 slab, so ``state_dict()`` keeps the reference's keys and shapes
(deepchem/models/torch_models/dmpnn.py:246-449) while ONE C call runs forward + L2 loss + backward and one fused
launch runs Adam.  The per-layer autograd path needed ~2 ms of Python per step to issue 1.6 ms of GPU work at
B = 4096 (scripts/dmpnn_profile.py)."""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._lib import check
from .engine import AdamSlabState

_AGG = {"mean": 0, "sum": 1, "norm": 2}
_workspaces = {}


def _ws(nbytes, device):
    key = (device.type, device.index)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def tables_struct(topo):
    """dcgc_dmpnn_tables for a DmpnnTopology (cached on the object)."""
    st = getattr(topo, "_c_tables", None)
    if st is not None:
        return st
    st = _lib.DmpnnTables()
    st.n_mols, st.n_atoms, st.n_rows = topo.n_mols, topo.n_atoms, topo.n_rows
    for name in ("mol_ptr", "a2b_ptr", "a2b_idx", "a2b_t_ptr", "a2b_t_idx", "map_ptr", "map_idx", "map_t_ptr",
                 "map_t_idx"):
        setattr(st, name, getattr(topo, name).data_ptr())
    topo._c_tables = st
    return st


class DmpnnEngine(AdamSlabState):
    """Owns the parameter / gradient / Adam slabs of one ``DMPNN`` module."""

    def __init__(self, model, device, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model, self.device = model, torch.device(device)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        enc, ffn = model.encoder, model.ffn
        cfg = _lib.DmpnnModelConfig()
        cfg.atom_fdim, cfg.bond_fdim = enc.atom_fdim, enc.concat_fdim - enc.atom_fdim
        cfg.hidden, cfg.depth = enc.W_i.out_features, enc.depth
        cfg.ffn_layers = len(ffn.linears)
        cfg.ffn_hidden = ffn.linears[0].out_features
        cfg.n_out = ffn.linears[-1].out_features
        cfg.aggregation, cfg.aggregation_norm = _AGG[enc.aggregation], float(enc.aggregation_norm)
        cfg.gemm_mode = enc.gemm_mode
        self.cfg = cfg
        offs = (ctypes.c_int64 * (4 + 2 * cfg.ffn_layers))()
        n_params = ctypes.c_int64()
        check(_lib.lib().dcgc_dmpnn_model_layout(ctypes.byref(cfg), offs, ctypes.byref(n_params)))
        self.offsets, self.n_params = list(offs), n_params.value
        dev = self.device
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        tensors = [enc.W_i.weight, enc.W_h.weight, enc.W_o.weight, enc.W_o.bias]
        for lin in ffn.linears:
            tensors += [lin.weight, lin.bias]
        self._slots = []
        for p, off in zip(tensors, self.offsets):
            n = p.numel()
            self._slots.append((p, self.params[off:off + n].view(p.shape), self.grads[off:off + n].view(p.shape)))
        self.adopt()

    @staticmethod
    def eligible(model, mode="regression", global_features_size=0):
        """The engine covers the reference's default fit path: regression, ReLU, dropout 0, no encoder bias, no
        global features, >= 2 feed-forward linears, widths that are multiples of 4."""
        enc, ffn = model.encoder, model.ffn
        if mode != "regression" or global_features_size:
            return False
        if enc.bias or not isinstance(getattr(enc, "activation", None), nn.ReLU) or enc.depth < 2:
            return False
        if enc.dropout.p != 0.0 or enc.aggregation not in _AGG:
            return False
        if not isinstance(getattr(ffn, "activation", None), nn.ReLU):
            return False
        n = len(ffn.linears)
        if n < 2 or n > _lib.DMPNN_MAX_FFN or any(d.p != 0.0 for d in ffn.dropout_p):
            return False
        h, fh = enc.W_i.out_features, ffn.linears[0].out_features
        if h % 4 or fh % 4 or enc.W_o.out_features != h or ffn.linears[0].in_features != h:
            return False
        return all(l.out_features == fh for l in ffn.linears[:-1]) and all(l.in_features == fh for l in ffn.linears[1:])

    def adopt(self):
        """Copy parameter values into the slab where they are not views of it yet; point p.data / p.grad at it."""
        with torch.no_grad():
            for p, view, gview in self._slots:
                if p.data_ptr() != view.data_ptr() or p.data.stride() != view.stride():
                    view.copy_(p.data.to(view.device))
                    p.data = view
                p.grad = gview

    def aliased(self):
        return all(p.data_ptr() == v.data_ptr() for p, v, _ in self._slots)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _workspace(self, topo):
        L = _lib.lib()
        nbytes = int(L.dcgc_dmpnn_model_workspace_bytes(ctypes.byref(self.cfg), topo.n_rows, topo.n_atoms, topo.n_mols))
        if nbytes < 0:
            raise ValueError(_lib.last_error())
        buf = _ws(nbytes + 256, self.device)
        pad = (-buf.data_ptr()) % 256
        return buf.data_ptr() + pad, buf.numel() - pad

    def train_step(self, topo, atom_features, f_ini, y, w, out=None):
        """forward + loss + backward; gradients land in self.grads.  Returns the device loss scalar."""
        if not self.aliased():
            self.adopt()
        ws, ws_bytes = self._workspace(topo)
        check(_lib.lib().dcgc_dmpnn_model_train_step(
            ctypes.byref(self.cfg), ctypes.byref(tables_struct(topo)), atom_features.data_ptr(),
            atom_features.stride(0), f_ini.data_ptr(), f_ini.stride(0), y.data_ptr(),
            w.data_ptr() if w is not None else None, self.params.data_ptr(), self.grads.data_ptr(), ws, ws_bytes,
            self.loss.data_ptr(), out.data_ptr() if out is not None else None, self._stream()))
        return self.loss

    def forward(self, topo, atom_features, f_ini, want_encoding=False):
        if not self.aliased():
            self.adopt()
        ws, ws_bytes = self._workspace(topo)
        out = torch.empty(topo.n_mols, self.cfg.n_out, dtype=torch.float32, device=self.device)
        enc = torch.empty(topo.n_mols, self.cfg.hidden, dtype=torch.float32, device=self.device) if want_encoding else None
        check(_lib.lib().dcgc_dmpnn_model_forward(
            ctypes.byref(self.cfg), ctypes.byref(tables_struct(topo)), atom_features.data_ptr(),
            atom_features.stride(0), f_ini.data_ptr(), f_ini.stride(0), self.params.data_ptr(), ws, ws_bytes,
            out.data_ptr(), enc.data_ptr() if enc is not None else None, self._stream()))
        return (out, enc) if want_encoding else out

    def adam_step(self, grad_scale=1.0):
        self.step_count += 1
        check(_lib.lib().dcgc_adam_step(self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                        self.exp_avg_sq.data_ptr(), self.n_params, self.lr, self.betas[0],
                                        self.betas[1], self.eps, self.step_count, grad_scale, self._stream()))

    def state_dict(self):
        return self.optimizer_state_dict(self.model.parameters())

    def load_state_dict(self, sd):
        self.load_optimizer_state_dict(sd, self.model.parameters())
