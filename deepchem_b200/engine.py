"""Flat-slab training engine binding (dcgc_gcmodel_* in include/dcgc.h).

All parameters of a ``_GraphConvTorchModel`` are re-homed into ONE contiguous fp32 buffer whose
layout is the one the C++ engine computes; every ``nn.Parameter`` (and BatchNorm running buffer)
becomes a view into it, so ``state_dict()`` keeps the reference's keys and shapes
(``graph_convs.{i}.W_list.{k}``, ``batch_norms.{i}.*``, ``dense.*``, ...) while one C call
runs forward + loss + backward and writes every gradient into a parallel flat gradient slab
(``p.grad`` are views of it).  Data parallel = one all-reduce of that slab; the optimizer is one
fused Adam launch.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import check
from .mol_graphs import GROUP_STRIDE

_workspaces = {}


def _ws(nbytes, device):
    key = (device.type, device.index)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def topology_struct(topo):
    """dcgc_topology for a DeviceTopology (cached on the object)."""
    st = getattr(topo, "_c_struct", None)
    if st is not None:
        return st
    st = _lib.Topology()
    st.n_atoms, st.n_edges, st.n_segments, st.n_tiles = topo.n_atoms, topo.n_edges, topo.n_segments, topo.n_tiles
    for d in range(11):
        st.deg_count[d] = topo.deg_count[d]
    ptr = getattr(topo, "ptr", None) or (lambda name: getattr(topo, name).data_ptr())
    for name in ("row_ptr", "col_idx", "t_row_ptr", "t_src", "t_slot", "mol_ptr", "mol_atoms", "membership", "tiles"):
        setattr(st, name, ptr(name))
    st.symmetric = 1 if getattr(topo, "symmetric", False) else 0
    st.groups = ptr("groups") + 4 * GROUP_STRIDE      # past the header row
    st.n_groups, st.group_max_rows = topo.n_groups, topo.group_max_rows
    st.group_max_entries = topo.group_max_entries
    rec = getattr(topo, "mg_records", None)
    st.mg_records = rec.data_ptr() if rec is not None else None
    topo._c_struct = st
    return st


class AdamSlabState(object):
    """Optimizer state of a flat-slab engine in ``torch.optim.Adam``'s own ``state_dict`` layout.

    The reference stores ``optimizer.state_dict()`` under 'optimizer_state_dict' (torch_model.py:1011-1017) and its
    restore calls ``Adam.load_state_dict`` on it (torch_model.py:1085-1090); the engines keep their moments in two
    slabs, so checkpoints are written per parameter — ``state[i] = {step, exp_avg, exp_avg_sq}`` in the order of
    ``model.parameters()`` plus ``param_groups`` — and read back from that layout.  A checkpoint written by an
    engine model therefore restores into the per-layer autograd model (plain ``torch.optim.Adam``), into the
    reference, and the other way round.  Needs: ``_slots`` [(parameter, slab view, grad view)], ``params``,
    ``exp_avg``, ``exp_avg_sq``, ``step_count``, ``lr``, ``betas``, ``eps``."""

    def _moment_views(self, p_view):
        off = p_view.storage_offset() - self.params.storage_offset()
        return (torch.as_strided(self.exp_avg, p_view.shape, p_view.stride(), off),
                torch.as_strided(self.exp_avg_sq, p_view.shape, p_view.stride(), off))

    def optimizer_state_dict(self, parameters):
        parameters = list(parameters)
        view_of = {id(p): v for p, v, _ in self._slots}
        state = {}
        if self.step_count > 0:                       # torch creates a parameter's state at its first step
            for i, p in enumerate(parameters):
                ea, eas = self._moment_views(view_of[id(p)])
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": ea.clone(),
                            "exp_avg_sq": eas.clone()}
        group = dict(torch.optim.Adam([torch.zeros(1)], lr=self.lr, betas=tuple(self.betas),
                                      eps=self.eps).state_dict()["param_groups"][0])
        group["params"] = list(range(len(parameters)))
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd, parameters):
        if "exp_avg" in sd and "state" not in sd:     # round-1 checkpoints: the raw slabs
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.step_count = int(sd["step"])
            return
        parameters = list(parameters)
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(parameters):
            raise ValueError("optimizer state has %d parameters, the model has %d" % (len(ids), len(parameters)))
        view_of = {id(p): v for p, v, _ in self._slots}
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        step = 0
        for key, p in zip(ids, parameters):
            st = sd["state"].get(key)
            if st is None:
                continue
            ea, eas = self._moment_views(view_of[id(p)])
            if tuple(st["exp_avg"].shape) != tuple(ea.shape):
                raise ValueError("optimizer state of parameter %d has shape %s, expected %s"
                                 % (key, tuple(st["exp_avg"].shape), tuple(ea.shape)))
            ea.copy_(st["exp_avg"])
            eas.copy_(st["exp_avg_sq"])
            step = max(step, int(float(st["step"])))
        self.step_count = step
        g0 = groups[0]
        self.lr, self.betas, self.eps = float(g0["lr"]), tuple(g0["betas"]), float(g0["eps"])


class FlatEngine(AdamSlabState):
    """Owns the parameter / gradient / optimizer slabs of one model."""

    def __init__(self, model, device, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model = model
        self.device = torch.device(device)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        self._mailboxes = None        # parallel.PeerMailboxes of a sync_batch_norm model, made at the first step
        L = len(model.graph_convs)
        cfg = _lib.GcModelConfig()
        cfg.n_layers = L
        cfg.n_feat = model.graph_convs[0].number_input_features
        for l, conv in enumerate(model.graph_convs):
            cfg.widths[l] = conv.out_channel
        cfg.dense = model.dense.out_features
        head = model.reshape_dense if model.mode == "classification" else model.regression_dense
        cfg.n_out = head.out_features
        cfg.n_classes = model.n_classes if model.mode == "classification" else 1
        cfg.mode = 1 if model.mode == "classification" else 0
        cfg.batch_norm = 1 if isinstance(model.batch_norms[0], torch.nn.BatchNorm1d) else 0
        cfg.gemm_mode = model.gemm_mode
        cfg.bn_eps, cfg.bn_momentum = 1e-3, 0.99
        self.cfg = cfg
        self.head = head
        n_off = 4 * L + 6
        offs = (ctypes.c_int64 * n_off)()
        boffs = (ctypes.c_int64 * (2 * (L + 1)))()
        n_params, n_bn = ctypes.c_int64(), ctypes.c_int64()
        check(_lib.lib().dcgc_gcmodel_layout(ctypes.byref(cfg), offs, boffs, ctypes.byref(n_params),
                                             ctypes.byref(n_bn)))
        self.offsets, self.bn_offsets = list(offs), list(boffs)
        self.n_params, self.n_bn = n_params.value, n_bn.value
        dev = self.device
        self.params = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.bn_running = torch.zeros(max(self.n_bn, 1), dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._slots = []      # (parameter, param view, grad view)
        self._bn_slots = []   # (module, name, view)
        self._build_slots()
        self.adopt()

    @staticmethod
    def eligible(model):
        """The engine covers the standard fit path: classification / regression without the
        uncertainty head, widths that are multiples of 4.  Anything else uses the autograd layers."""
        if getattr(model, "uncertainty", False):
            return False
        if getattr(model, "sync_batch_norm", False):
            # statistics exchanged between ranks inside the step: the engine does it through peer-memory mailboxes
            # (dcgc_gcmodel_train_step_sync) for the ranks of one node on CUDA; otherwise the per-layer autograd path
            # with parallel.SyncBatchNorm1d (any backend, e.g. gloo on the CPU)
            import torch.distributed as dist
            if not (torch.cuda.is_available() and dist.is_available() and dist.is_initialized()
                    and dist.get_backend() == "nccl" and dist.get_world_size() <= _lib.SYNC_MAX_RANKS
                    and os.environ.get("DCGC_ENGINE_SYNCBN", "1") != "0"):
                return False
        if len(model.graph_convs) > _lib.MODEL_MAX_LAYERS:
            return False
        widths = [c.out_channel for c in model.graph_convs] + [model.dense.out_features]
        if any(w % 4 for w in widths):
            return False
        ins = [c.number_input_features for c in model.graph_convs]
        if ins[1:] != widths[:len(ins) - 1] or model.dense.in_features != widths[len(ins) - 1]:
            return False
        bn = [isinstance(b, torch.nn.BatchNorm1d) for b in model.batch_norms]
        return all(bn) or not any(bn)

    # ------------------------------------------------------------------ slab <-> module views
    def _view(self, slab, off, shape):
        n = 1
        for s in shape:
            n *= s
        return slab[off:off + n].view(*shape)

    def _build_slots(self):
        m, cfg = self.model, self.cfg
        L = cfg.n_layers
        for l, conv in enumerate(m.graph_convs):
            f, c = conv.number_input_features, conv.out_channel
            fp = (f + 3) // 4 * 4
            w_off, b_off, g_off, be_off = self.offsets[4 * l:4 * l + 4]
            wp = self._view(self.params, w_off, (11, 2 * fp, c))
            gp = self._view(self.grads, w_off, (11, 2 * fp, c))
            for k in range(21):
                if k == 20:
                    d, r0 = 0, 0
                elif k % 2 == 0:
                    d, r0 = k // 2 + 1, fp      # neighbour weight of degree d
                else:
                    d, r0 = (k - 1) // 2 + 1, 0  # self weight of degree d
                self._slots.append((conv.W_list[k], wp[d, r0:r0 + f, :], gp[d, r0:r0 + f, :]))
            bp = self._view(self.params, b_off, (21, c))
            bg = self._view(self.grads, b_off, (21, c))
            for k in range(21):
                self._slots.append((conv.b_list[k], bp[k], bg[k]))
            if cfg.batch_norm:
                self._add_bn(m.batch_norms[l], g_off, be_off, self.bn_offsets[2 * l], self.bn_offsets[2 * l + 1], c)
        w_off, b_off, g_off, be_off = self.offsets[4 * L:4 * L + 4]
        d_out, d_in = m.dense.out_features, m.dense.in_features
        self._slots.append((m.dense.weight, self._view(self.params, w_off, (d_out, d_in)),
                            self._view(self.grads, w_off, (d_out, d_in))))
        self._slots.append((m.dense.bias, self._view(self.params, b_off, (d_out,)),
                            self._view(self.grads, b_off, (d_out,))))
        if cfg.batch_norm:
            self._add_bn(m.batch_norms[L], g_off, be_off, self.bn_offsets[2 * L], self.bn_offsets[2 * L + 1], d_out)
        hw, hb = self.offsets[4 * L + 4], self.offsets[4 * L + 5]
        n_out = self.head.out_features
        self._slots.append((self.head.weight, self._view(self.params, hw, (n_out, 2 * d_out)),
                            self._view(self.grads, hw, (n_out, 2 * d_out))))
        self._slots.append((self.head.bias, self._view(self.params, hb, (n_out,)),
                            self._view(self.grads, hb, (n_out,))))

    def _add_bn(self, bn, g_off, b_off, m_off, v_off, c):
        self._slots.append((bn.weight, self._view(self.params, g_off, (c,)), self._view(self.grads, g_off, (c,))))
        self._slots.append((bn.bias, self._view(self.params, b_off, (c,)), self._view(self.grads, b_off, (c,))))
        self._bn_slots.append((bn, "running_mean", self._view(self.bn_running, m_off, (c,))))
        self._bn_slots.append((bn, "running_var", self._view(self.bn_running, v_off, (c,))))

    def adopt(self):
        """Copy current parameter values into the slab (where they are not already views of it) and
        re-point ``p.data`` / ``p.grad`` / BN buffers at the slab."""
        with torch.no_grad():
            for p, view, gview in self._slots:
                if p.data_ptr() != view.data_ptr() or p.data.stride() != view.stride():
                    view.copy_(p.data.to(view.device))
                    p.data = view
                p.grad = gview
            for bn, name, view in self._bn_slots:
                buf = getattr(bn, name)
                if buf.data_ptr() != view.data_ptr():
                    view.copy_(buf.to(view.device))
                    setattr(bn, name, view)

    def aliased(self):
        return all(p.data_ptr() == v.data_ptr() for p, v, _ in self._slots) and \
            all(getattr(bn, name).data_ptr() == v.data_ptr() for bn, name, v in self._bn_slots)

    # ------------------------------------------------------------------ compute
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def grad_slices(self):
        """[(lo, hi)] of the gradient slab in the order the backward pass completes them (the order of the
        ``grad_events`` of dcgc_gcmodel_train_step_ev): [dense, its BatchNorm, head], then conv layer L-1 ... 0."""
        L = self.cfg.n_layers
        conv_w = [self.offsets[4 * l] for l in range(L)] + [self.offsets[4 * L]]
        return [(conv_w[L], self.n_params)] + [(conv_w[l], conv_w[l + 1]) for l in range(L - 1, -1, -1)]

    def train_step(self, topo, x, y, w, n_samples, out=None, forward_event=None, grad_events=None):
        """forward + loss + backward; gradients land in self.grads.  Returns the device loss scalar.
        ``forward_event`` / ``grad_events`` (torch.cuda.Event): recorded between forward and backward / when each
        slice of ``grad_slices()`` is final."""
        if not self.aliased():
            self.adopt()
        L = _lib.lib()
        self.cfg.input_exact = 1 if getattr(x, "_dcgc_input_exact", False) else 0
        nbytes = int(L.dcgc_gcmodel_workspace_bytes(ctypes.byref(self.cfg), topo.n_atoms, topo.n_segments))
        ws = _ws(nbytes, self.device)
        n_ev, evs = 0, None
        if grad_events:
            n_ev = len(grad_events)
            evs = (ctypes.c_void_p * n_ev)(*[e.cuda_event for e in grad_events])
        sync = None
        if getattr(self.model, "sync_batch_norm", False) and self.cfg.batch_norm:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                if self._mailboxes is None:
                    from .parallel import PeerMailboxes
                    cap = max([self.cfg.widths[l] for l in range(self.cfg.n_layers)] + [self.cfg.dense])
                    self._mailboxes = PeerMailboxes(cap)
                sync = ctypes.byref(self._mailboxes.struct(2 * (self.cfg.n_layers + 1)))
        check(L.dcgc_gcmodel_train_step_sync(
            ctypes.byref(self.cfg), ctypes.byref(topology_struct(topo)), x.data_ptr(), x.stride(0),
            y.data_ptr(), w.data_ptr() if w is not None else None, n_samples, self.params.data_ptr(),
            self.grads.data_ptr(), self.bn_running.data_ptr() if self.n_bn else None, ws.data_ptr(), ws.numel(),
            self.loss.data_ptr(), out.data_ptr() if out is not None else None,
            forward_event.cuda_event if forward_event is not None else None, evs, n_ev, sync, self._stream()))
        return self.loss

    def forward(self, topo, x, n_samples, training=False, want_probs=True):
        """-> (out [n_samples, n_out], probs or None, fingerprint [n_segments, 2D])"""
        if not self.aliased():
            self.adopt()
        L = _lib.lib()
        cfg = self.cfg
        cfg.input_exact = 1 if getattr(x, "_dcgc_input_exact", False) else 0
        nbytes = int(L.dcgc_gcmodel_workspace_bytes(ctypes.byref(cfg), topo.n_atoms, topo.n_segments))
        ws = _ws(nbytes, self.device)
        out = torch.empty(n_samples, cfg.n_out, dtype=torch.float32, device=self.device)
        probs = torch.empty_like(out) if (want_probs and cfg.mode == 1) else None
        fp = torch.empty(topo.n_segments, 2 * cfg.dense, dtype=torch.float32, device=self.device)
        check(L.dcgc_gcmodel_forward(
            ctypes.byref(cfg), ctypes.byref(topology_struct(topo)), x.data_ptr(), x.stride(0), n_samples,
            self.params.data_ptr(), self.bn_running.data_ptr() if self.n_bn else None, 1 if training else 0,
            ws.data_ptr(), ws.numel(), out.data_ptr(), probs.data_ptr() if probs is not None else None,
            fp.data_ptr(), self._stream()))
        return out, probs, fp

    def adam_step(self, grad_scale=1.0):
        self.step_count += 1
        check(_lib.lib().dcgc_adam_step(self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                        self.exp_avg_sq.data_ptr(), self.n_params, self.lr, self.betas[0],
                                        self.betas[1], self.eps, self.step_count, grad_scale, self._stream()))

    # kernel launches of one train step (bench.py gpu_launches bookkeeping)
    def launches_per_step(self):
        L = self.cfg.n_layers
        bn = 3 if self.cfg.batch_norm else 1
        fwd = L * (4 + (2 if self.cfg.batch_norm else 0)) + 1 + (2 if self.cfg.batch_norm else 0) + 2
        bwd = 5 + 1 + (bn + 3) + L * (1 + bn + 3) + (L - 1) * 2
        return fwd + bwd + 1

    def state_dict(self):
        return self.optimizer_state_dict(self.model.parameters())

    def load_state_dict(self, sd):
        self.load_optimizer_state_dict(sd, self.model.parameters())
