"""D-MPNN edge message passing on the B200 path.

Mirrors deepchem/models/torch_models/dmpnn.py (``_MapperDMPNN`` :38, ``DMPNN`` :246, ``DMPNNModel`` :452)
and ``DMPNNEncoderLayer`` / ``PositionwiseFeedForward`` (deepchem/models/torch_models/layers.py:1261-1649,
795-910): same constructor arguments, parameter names (``encoder.W_i/W_h/W_o``, ``ffn.linears.N``),
forward signatures and error behaviour.

What changes underneath:
  * the per-molecule Python mapper + per-molecule H2D copies + torch_geometric collation
    (dmpnn.py:592-755) become one C++ pass over a packed shard (``dcgc_dmpnn_plan/build``) and two H2D
    copies per batch (integer slab, features);
  * ``message[mapping].sum(1)`` and ``h_message[atom_to_incoming_bonds].sum(1)`` (layers.py:1629, 1539)
    run as CSR gather-sum kernels with transposed-CSR backward (no [rows, K, hidden] temporary, no
    atomics); W_i / W_h / W_o are the grouped-GEMM entry points (fp32 SIMT or tcgen05 TF32x3);
    ``cat(atom_features, messages)`` is never materialised; the readout is one segmented kernel instead
    of a Python loop over molecules (layers.py:1571-1583);
  * the dead W_h products of the reference loop (the gather always reads ``message``, layers.py:1628-1633,
    so only the last iteration's W_h output is used) are not computed — results are identical.
There is no CPU fallback.
"""
import ctypes
import logging
import os
import time

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import ACT_NONE
from .dmpnn_data import GraphData, PackedGraphs
from .layers import gemm_mode_code

logger = logging.getLogger(__name__)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


_DM_FIELDS = (
    # name, offset attribute, length
    ("row_of_mol", "off_row_of_mol", lambda i: i.n_mols + 1),
    ("mol_ptr", "off_mol_ptr", lambda i: i.n_mols + 1),
    ("bond_src", "off_bond_src", lambda i: i.n_rows),
    ("bond_edge", "off_bond_edge", lambda i: i.n_rows),
    ("a2b_ell", "off_a2b_ell", lambda i: i.n_atoms * i.k),
    ("map_ell", "off_map_ell", lambda i: i.n_rows * i.k),
    ("a2b_ptr", "off_a2b_ptr", lambda i: i.n_atoms + 1),
    ("a2b_idx", "off_a2b_idx", lambda i: i.n_a2b_entries),
    ("a2b_t_ptr", "off_a2b_t_ptr", lambda i: i.n_rows + 1),
    ("a2b_t_idx", "off_a2b_t_idx", lambda i: i.n_a2b_entries),
    ("map_ptr", "off_map_ptr", lambda i: i.n_rows + 1),
    ("map_idx", "off_map_idx", lambda i: i.n_map_entries),
    ("map_t_ptr", "off_map_t_ptr", lambda i: i.n_rows + 1),
    ("map_t_idx", "off_map_t_idx", lambda i: i.n_map_entries),
)


class DmpnnLayout(object):
    """Host slab of the batched D-MPNN index tables (C++ builder) + int32 numpy views."""

    def __init__(self, info, slab, slab_tensor=None):
        self.info, self.slab, self.slab_tensor = info, slab, slab_tensor
        for name, off_attr, length in _DM_FIELDS:
            off, n = getattr(info, off_attr), int(length(info))
            setattr(self, name, slab[off:off + 4 * n].view(np.int32))
        k = int(info.k)
        self.a2b_ell = self.a2b_ell.reshape(-1, k)
        self.map_ell = self.map_ell.reshape(-1, k)

    n_mols = property(lambda self: int(self.info.n_mols))
    n_atoms = property(lambda self: int(self.info.n_atoms))
    n_bonds = property(lambda self: int(self.info.n_bonds))
    n_rows = property(lambda self: int(self.info.n_rows))
    k = property(lambda self: int(self.info.k))

    @staticmethod
    def build(packed, keep_pads=False, pinned=False):
        L = _lib.lib()
        info = _lib.DmpnnInfo()
        args = (packed.n_mols, _ptr(packed.node_ptr), _ptr(packed.edge_ptr), _ptr(packed.edge_src),
                _ptr(packed.edge_dst))
        _lib.check(L.dcgc_dmpnn_plan(*args, 1 if keep_pads else 0, ctypes.byref(info)))
        nbytes = int(info.slab_bytes)
        if pinned and torch.cuda.is_available():
            t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
            slab = t.numpy()[:nbytes]
        else:
            t, slab = None, np.empty(nbytes, dtype=np.uint8)
        _lib.check(L.dcgc_dmpnn_build(*args, ctypes.byref(info), _ptr(slab)))
        return DmpnnLayout(info, slab, t)

    def to_device(self, device, non_blocking=True):
        return DmpnnTopology(self, device, non_blocking)


class DmpnnTopology(object):
    """Device-resident tables: one buffer, int32 views, plus the two gather patterns as CsrPair."""

    def __init__(self, layout, device, non_blocking=True):
        self.layout, self.device = layout, torch.device(device)
        info = layout.info
        nbytes = int(info.slab_bytes)
        src = layout.slab_tensor if layout.slab_tensor is not None else torch.from_numpy(layout.slab)
        self.buffer = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        if nbytes:
            self.buffer[:nbytes].copy_(src[:nbytes], non_blocking=non_blocking)
        for name, off_attr, length in _DM_FIELDS:
            off, n = getattr(info, off_attr), int(length(info))
            setattr(self, name, self.buffer[off:off + 4 * n].view(torch.int32))
        k = layout.k
        self.a2b_ell = self.a2b_ell.view(-1, k)
        self.map_ell = self.map_ell.view(-1, k)
        self.n_mols, self.n_atoms, self.n_rows, self.k = layout.n_mols, layout.n_atoms, layout.n_rows, k
        self.mapping_csr = ops.CsrPair(self.map_ptr, self.map_idx, self.map_t_ptr, self.map_t_idx, self.n_rows,
                                       self.n_rows)
        self.a2b_csr = ops.CsrPair(self.a2b_ptr, self.a2b_idx, self.a2b_t_ptr, self.a2b_t_idx, self.n_atoms,
                                   self.n_rows)
        self.molecules_unbatch_key = np.diff(layout.mol_ptr).tolist()


def _attach(tensor, topo):
    tensor._dcgc_dmpnn = topo
    return tensor


def _topology_from_tables(atom_to_incoming_bonds, mapping, molecules_unbatch_key, keep_pads):
    """Slow path for hand-built inputs (the reference's layer tests): derive the CSR patterns from the ELL
    tensors on the host.  Entries equal to a row's own pad value cannot be told apart from real ones here,
    so EVERY entry is kept, negative ones resolved as torch indexing would (exact for any bias)."""
    dev = mapping.device
    a2b = atom_to_incoming_bonds.detach().cpu().numpy().astype(np.int64)
    mp = mapping.detach().cpu().numpy().astype(np.int64)
    n_rows, n_atoms = mp.shape[0], a2b.shape[0]

    def csr(tab, n_in):
        tab = np.where(tab < 0, tab + n_in, tab)
        ptr = (np.arange(tab.shape[0] + 1) * tab.shape[1]).astype(np.int32)
        idx = tab.reshape(-1).astype(np.int32)
        order = np.argsort(idx, kind="stable")
        t_idx = (order // max(tab.shape[1], 1)).astype(np.int32)
        t_ptr = np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=n_in))]).astype(np.int32)
        return [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (ptr, idx, t_ptr, t_idx)]

    class _T(object):
        pass
    t = _T()
    t.n_rows, t.n_atoms = n_rows, n_atoms
    t.mapping_csr = ops.CsrPair(*csr(mp, n_rows), n_rows, n_rows)
    t.a2b_csr = ops.CsrPair(*csr(a2b, n_rows), n_atoms, n_rows)
    key = list(molecules_unbatch_key)
    t.n_mols = len(key)
    t.mol_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(key)]).astype(np.int32)).to(dev)
    t.molecules_unbatch_key = key
    return t


class _MapperDMPNN(object):
    """Reference-shaped per-molecule view (dmpnn.py:38-243) computed by the C++ builder on a batch of one.
    ``values`` = (atom_features, f_ini_atoms_bonds, atom_to_incoming_bonds, mapping, global_features)."""

    def __init__(self, graph):
        self.num_atoms, self.num_bonds = graph.num_nodes, graph.num_edges
        self.num_atom_features, self.num_bond_features = graph.num_node_features, graph.num_edge_features
        self.atom_features, self.bond_features = graph.node_features, graph.edge_features
        self.bond_index = graph.edge_index
        self.global_features = getattr(graph, "global_features", np.empty(0))
        packed = PackedGraphs.from_graphs([graph], self.num_atom_features, self.num_bond_features)
        lay = DmpnnLayout.build(packed)
        e = self.num_bonds
        # one molecule: row offset 0, so the batched tables are the reference's local ones; the builder's
        # table width is max(1, max in-degree), as in dmpnn.py:222-223
        self.atom_to_incoming_bonds = lay.a2b_ell.astype(int)
        self.mapping = lay.map_ell.astype(int)
        if e == 0:
            self.bond_to_ini_atom = np.empty(0)
            self.f_ini_atoms_bonds = np.zeros((1, self.num_atom_features + self.num_bond_features))
        else:
            self.bond_to_ini_atom = np.asarray(self.bond_index)[0]
            f = np.hstack((np.asarray(self.atom_features)[self.bond_to_ini_atom], np.asarray(self.bond_features)))
            self.f_ini_atoms_bonds = np.pad(f, ((0, 1), (0, 0)))

    @property
    def values(self):
        return (self.atom_features, self.f_ini_atoms_bonds, self.atom_to_incoming_bonds, self.mapping,
                self.global_features)


_ACTS = {'relu': nn.ReLU, 'leakyrelu': lambda: nn.LeakyReLU(0.1), 'prelu': nn.PReLU, 'tanh': nn.Tanh,
         'selu': nn.SELU, 'elu': nn.ELU}


def _linear(x, lin, act_code, mode, x2=None, split=None):
    """act(x . W^T + b) through the grouped-GEMM kernels; with x2, W acts on cat(x[:, :split], x2) where x may
    be a zero-padded view wider than ``split`` (zero weight rows are inserted for the pad columns)."""
    w = lin.weight
    if x2 is None:
        k = x.shape[1]
        wt = w.t()
        if k != w.shape[1]:                      # zero-padded input view
            wt = torch.cat([wt, wt.new_zeros(k - w.shape[1], w.shape[0])], 0)
        return ops.GroupLinear2Fn.apply(x, None, wt, lin.bias, act_code, mode, ops.forward_gemm_mode(mode))
    k1 = x.shape[1]
    parts = [w[:, :split].t()]
    if k1 != split:
        parts.append(w.new_zeros(k1 - split, w.shape[0]))
    parts.append(w[:, split:].t())
    return ops.GroupLinear2Fn.apply(x, x2, torch.cat(parts, 0), lin.bias, act_code, mode, ops.forward_gemm_mode(mode))


class DMPNNEncoderLayer(nn.Module):
    """Encoder of the Directed Message Passing Neural Network (layers.py:1261-1649)."""

    def __init__(self, use_default_fdim=True, atom_fdim=133, bond_fdim=14, d_hidden=300, depth=3, bias=False,
                 activation='relu', dropout_p=0.0, aggregation='mean', aggregation_norm=100, gemm_mode="tf32x3"):
        super(DMPNNEncoderLayer, self).__init__()
        if use_default_fdim:       # GraphConvConstants.ATOM_FDIM / BOND_FDIM (dmpnn_featurizer.py)
            atom_fdim, bond_fdim = 133, 14
        self.atom_fdim, self.concat_fdim = atom_fdim, atom_fdim + bond_fdim
        self.depth, self.aggregation, self.aggregation_norm = depth, aggregation, aggregation_norm
        if activation in _ACTS:
            self.activation = _ACTS[activation]()
        self._act_code = ops.act_code(getattr(self, "activation", None))
        self.dropout = nn.Dropout(dropout_p)
        self.W_i = nn.Linear(self.concat_fdim, d_hidden, bias=bias)
        self.W_h = nn.Linear(d_hidden, d_hidden, bias=bias)
        self.W_o = nn.Linear(self.atom_fdim + d_hidden, d_hidden)
        self.gemm_mode = gemm_mode_code(gemm_mode)
        self.bias = bias

    def _act_linear(self, x, lin, x2=None, split=None):
        """activation(lin(...)) with the activation fused into the GEMM epilogue when it is ReLU / tanh."""
        if self._act_code is not None:
            return _linear(x, lin, self._act_code, self.gemm_mode, x2, split)
        return self.activation(_linear(x, lin, ACT_NONE, self.gemm_mode, x2, split))

    def forward(self, atom_features, f_ini_atoms_bonds, atom_to_incoming_bonds, mapping, global_features,
                molecules_unbatch_key):
        topo = getattr(mapping, "_dcgc_dmpnn", None)
        if topo is None:
            topo = _topology_from_tables(atom_to_incoming_bonds, mapping, molecules_unbatch_key, self.bias)
        inp = _linear(f_ini_atoms_bonds, self.W_i, ACT_NONE, self.gemm_mode)     # layers.py:1622
        message = self.activation(inp)                                            # :1624
        if self.depth < 2:
            # the reference leaves h_message unbound for depth 1 (layers.py:1627-1637)
            raise NameError("name 'h_message' is not defined")
        for _ in range(1, self.depth):                                            # :1627-1629
            message = ops.CsrGatherSumFn.apply(message, topo.mapping_csr)
        # :1630-1633 — only the last iteration's W_h product is live (the gather never reads h_message)
        h_message = self.dropout(self.activation(inp + _linear(message, self.W_h, ACT_NONE, self.gemm_mode)))
        m2a = ops.CsrGatherSumFn.apply(h_message, topo.a2b_csr)                   # :1539
        atoms_hidden = self.dropout(self._act_linear(atom_features, self.W_o, m2a, self.atom_fdim))  # :1541-1547
        out = ops.SegmentReadoutFn.apply(atoms_hidden, topo.mol_ptr, topo.n_mols, self.aggregation,
                                         self.aggregation_norm)                   # :1550-1583
        if global_features.size()[0] != 0:                                        # :1644-1647
            if len(global_features.shape) == 1:
                global_features = global_features.view(len(out), -1)
            out = torch.cat([out, global_features], dim=1)
        return out


class PositionwiseFeedForward(nn.Module):
    """layers.py:795-910, linears through the GEMM kernels."""

    def __init__(self, d_input=1024, d_hidden=1024, d_output=1024, activation='leakyrelu', n_layers=1,
                 dropout_p=0.0, dropout_at_input_no_act=False, gemm_mode="tf32x3"):
        super(PositionwiseFeedForward, self).__init__()
        self.dropout_at_input_no_act = dropout_at_input_no_act
        if activation == "linear":
            self.activation = lambda x: x
        elif activation in _ACTS:
            self.activation = _ACTS[activation]()
        self._act_code = ops.act_code(self.activation) if activation != "linear" else ACT_NONE
        self.n_layers = n_layers
        d_output = d_output if d_output != 0 else d_input
        d_hidden = d_hidden if d_hidden != 0 else d_input
        if n_layers == 1:
            lin = [nn.Linear(d_input, d_output)]
        else:
            lin = [nn.Linear(d_input, d_hidden)] + [nn.Linear(d_hidden, d_hidden) for _ in range(n_layers - 2)] + \
                [nn.Linear(d_hidden, d_output)]
        self.linears = nn.ModuleList(lin)
        dropout_layer = nn.Dropout(dropout_p)
        self.dropout_p = nn.ModuleList([dropout_layer for _ in range(n_layers)])
        self.gemm_mode = gemm_mode_code(gemm_mode)

    def _lin(self, x, i, act):
        lin = self.linears[i]
        if act and self._act_code is not None:
            return _linear(x, lin, self._act_code, self.gemm_mode)
        y = _linear(x, lin, ACT_NONE, self.gemm_mode)
        return self.activation(y) if act else y

    def forward(self, x):
        if not self.n_layers:
            return x
        if self.n_layers == 1:
            if self.dropout_at_input_no_act:
                return self._lin(self.dropout_p[0](x), 0, False)
            return self.dropout_p[0](self._lin(x, 0, True))
        if self.dropout_at_input_no_act:
            x = self.dropout_p[0](x)
        for i in range(self.n_layers - 1):
            x = self.dropout_p[i](self._lin(x, i, True))
        return self._lin(x, self.n_layers - 1, False)


class DmpnnBatch(dict):
    """What ``DMPNN.forward`` reads from the PyG batch in the reference (dmpnn.py:425-441): the five tensors
    by key and the atoms-per-molecule list."""
    molecules_unbatch_key = None


class DMPNN(nn.Module):
    """Directed Message Passing Neural Network (dmpnn.py:246-449)."""

    def __init__(self, mode='regression', n_classes=3, n_tasks=1, global_features_size=0, use_default_fdim=True,
                 atom_fdim=133, bond_fdim=14, enc_hidden=300, depth=3, bias=False, enc_activation='relu',
                 enc_dropout_p=0.0, aggregation='mean', aggregation_norm=100, ffn_hidden=300,
                 ffn_activation='relu', ffn_layers=3, ffn_dropout_p=0.0, ffn_dropout_at_input_no_act=True,
                 gemm_mode="tf32x3"):
        super(DMPNN, self).__init__()
        self.mode, self.n_classes, self.n_tasks = mode, n_classes, n_tasks
        self.encoder = DMPNNEncoderLayer(use_default_fdim=use_default_fdim, atom_fdim=atom_fdim, bond_fdim=bond_fdim,
                                         d_hidden=enc_hidden, depth=depth, bias=bias, activation=enc_activation,
                                         dropout_p=enc_dropout_p, aggregation=aggregation,
                                         aggregation_norm=aggregation_norm, gemm_mode=gemm_mode)
        ffn_input = enc_hidden + global_features_size
        ffn_output = self.n_tasks if mode == 'regression' else self.n_tasks * self.n_classes
        self.ffn = PositionwiseFeedForward(d_input=ffn_input, d_hidden=ffn_hidden, d_output=ffn_output,
                                           activation=ffn_activation, n_layers=ffn_layers, dropout_p=ffn_dropout_p,
                                           dropout_at_input_no_act=ffn_dropout_at_input_no_act, gemm_mode=gemm_mode)

    def forward(self, batch):
        enc = self.encoder(batch['atom_features'], batch['f_ini_atoms_bonds'], batch['atom_to_incoming_bonds'],
                           batch['mapping'], batch['global_features'], batch.molecules_unbatch_key)
        output = self.ffn(enc)
        if self.mode == 'regression':
            return output
        if self.n_tasks == 1:
            output = output.view(-1, self.n_classes)
            return nn.functional.softmax(output, dim=1), output
        output = output.view(-1, self.n_tasks, self.n_classes)
        return nn.functional.softmax(output, dim=2), output


class _GraphDataset(object):
    """iterbatches() over a PackedGraphs shard (the contract of deepchem/data/datasets.py:843-898)."""

    def __init__(self, packed, y=None, w=None, ids=None, n_tasks=1):
        self.packed = packed
        n = packed.n_mols
        self.y = np.zeros((n, n_tasks), np.float32) if y is None else np.asarray(y).reshape(n, -1)
        self.w = np.ones_like(self.y, dtype=np.float32) if w is None else np.asarray(w).reshape(n, -1)
        self.ids = np.arange(n) if ids is None else np.asarray(ids)

    def __len__(self):
        return self.packed.n_mols

    def iterbatches(self, batch_size=None, epochs=1, deterministic=False, pad_batches=False):
        n = len(self)
        batch_size = batch_size or n
        for _ in range(epochs):
            perm = np.arange(n) if deterministic else np.random.permutation(n)
            for b in range(0, n, batch_size):
                idx = perm[b:b + batch_size]
                X = self.packed.slice(int(idx[0]), int(idx[-1]) + 1) if deterministic else self.packed.take(idx)
                yield X, self.y[idx], self.w[idx], self.ids[idx]


GraphDataset = _GraphDataset


class DMPNNModel(object):
    """Directed Message Passing Neural Network model (dmpnn.py:452-755): same constructor arguments;
    ``fit`` / ``predict`` take a dataset whose ``X`` is a sequence of GraphData-like objects or a
    PackedGraphs shard (``GraphDataset``)."""

    def __init__(self, mode='regression', n_classes=3, n_tasks=1, batch_size=1, global_features_size=0,
                 use_default_fdim=True, atom_fdim=133, bond_fdim=14, enc_hidden=300, depth=3, bias=False,
                 enc_activation='relu', enc_dropout_p=0.0, aggregation='mean', aggregation_norm=100, ffn_hidden=300,
                 ffn_activation='relu', ffn_layers=3, ffn_dropout_p=0.0, ffn_dropout_at_input_no_act=True,
                 learning_rate=0.001, model_dir=None, device=None, gemm_mode="tf32x3", log_frequency=100, **kwargs):
        if mode not in ('regression', 'classification'):
            raise ValueError("mode must be either 'regression' or 'classification'")
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("deepchem_b200.DMPNNModel needs a CUDA device (B200); there is no CPU path")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.mode, self.n_tasks, self.n_classes, self.batch_size = mode, n_tasks, n_classes, batch_size
        self.atom_fdim, self.bond_fdim = (133, 14) if use_default_fdim else (atom_fdim, bond_fdim)
        self.keep_pads = bool(bias)      # pad rows are non-zero only with bias (see dmpnn_layout.cpp)
        self.model = DMPNN(mode=mode, n_classes=n_classes, n_tasks=n_tasks, global_features_size=global_features_size,
                           use_default_fdim=use_default_fdim, atom_fdim=atom_fdim, bond_fdim=bond_fdim,
                           enc_hidden=enc_hidden, depth=depth, bias=bias, enc_activation=enc_activation,
                           enc_dropout_p=enc_dropout_p, aggregation=aggregation, aggregation_norm=aggregation_norm,
                           ffn_hidden=ffn_hidden, ffn_activation=ffn_activation, ffn_layers=ffn_layers,
                           ffn_dropout_p=ffn_dropout_p, ffn_dropout_at_input_no_act=ffn_dropout_at_input_no_act,
                           gemm_mode=gemm_mode).to(self.device)
        self.output_types = ['prediction'] if mode == 'regression' else ['prediction', 'loss']
        self._pytorch_optimizer = torch.optim.Adam(self.model.parameters(), lr=learning_rate)
        self._global_step, self.log_frequency, self.model_dir = 0, log_frequency, model_dir
        self._grad_slab = None
        # fused engine (one C call per step over flat slabs, dcgc_dmpnn_model_*) for the configurations it covers;
        # use_engine=False (or DCGC_DMPNN_ENGINE=0) keeps the per-layer autograd path over the same kernels
        self._engine, self._dp = None, False
        use_engine = kwargs.pop("use_engine", os.environ.get("DCGC_DMPNN_ENGINE", "1") != "0")
        if use_engine and self.device.type == "cuda":
            from .dmpnn_engine import DmpnnEngine
            if DmpnnEngine.eligible(self.model, mode, global_features_size):
                self._engine = DmpnnEngine(self.model, self.device, lr=learning_rate)
        self._prefetch_stream = None
        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 2
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
        self.host_workers = int(os.environ.get("DCGC_HOST_WORKERS", max(1, min(4, cores // ranks - 2))))

    # ------------------------------------------------------------------ batching
    def default_generator(self, dataset, epochs=1, mode='fit', deterministic=True, pad_batches=False, workers=None,
                          **kwargs):
        """(packed graphs + host layout, [y], [w]) per batch (dmpnn.py:677-755).  ``workers`` > 1 builds the index
        tables of upcoming batches concurrently on a thread pool (the C++ builder releases the GIL); batches are
        still yielded in dataset order."""
        workers = self.host_workers if workers is None else workers

        def build(X_b):
            packed = X_b if isinstance(X_b, PackedGraphs) else PackedGraphs.from_graphs(list(X_b), self.atom_fdim,
                                                                                         self.bond_fdim)
            return packed, DmpnnLayout.build(packed, keep_pads=self.keep_pads, pinned=True)

        batches = dataset.iterbatches(batch_size=self.batch_size, epochs=epochs, deterministic=deterministic,
                                      pad_batches=pad_batches)
        if workers <= 1:
            for (X_b, y_b, w_b, ids_b) in batches:
                yield (build(X_b), [y_b], [w_b])
            return
        import collections
        from concurrent.futures import ThreadPoolExecutor
        pending = collections.deque()
        with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="dcgc-dmpnn-layout") as pool:
            for (X_b, y_b, w_b, ids_b) in batches:
                pending.append((pool.submit(build, X_b), y_b, w_b))
                if len(pending) > workers:
                    fut, y0, w0 = pending.popleft()
                    yield (fut.result(), [y0], [w0])
            while pending:
                fut, y0, w0 = pending.popleft()
                yield (fut.result(), [y0], [w0])

    def _prepare_batch(self, batch):
        """Host -> device (dmpnn.py:645-675): integer slab + node / edge / global features, then f_ini assembled on
        the device.  The two index tensors keep the reference's values and carry the device topology."""
        (packed, layout), labels, weights = batch
        dev = self.device
        topo = layout.to_device(dev)
        # page-locked torch views of the shard's rows when it was pinned (PackedGraphs.pin_memory): true async DMA
        pn, pe = getattr(packed, "_pin_nf", None), getattr(packed, "_pin_ef", None)
        af = (pn if pn is not None else torch.from_numpy(packed.node_features)).to(dev, non_blocking=True)
        bf = (pe if pe is not None else torch.from_numpy(packed.edge_features)).to(dev, non_blocking=True)
        gf = torch.from_numpy(np.ascontiguousarray(packed.global_features).reshape(-1)).to(dev)
        fa = af.shape[1]
        af_pad = torch.zeros(af.shape[0], (fa + 3) // 4 * 4, device=dev)
        af_pad[:, :fa] = af
        atom_view = af_pad[:, :fa]
        bf_dev = bf if bf.shape[0] else bf.new_zeros(1, bf.shape[1])     # bond-free batch: keep a valid pointer
        f_ini = ops.dmpnn_concat_rows(af, bf_dev, topo.bond_src, topo.bond_edge, topo.n_rows)
        b = DmpnnBatch(atom_features=atom_view, f_ini_atoms_bonds=f_ini, atom_to_incoming_bonds=topo.a2b_ell,
                       mapping=_attach(topo.map_ell, topo), global_features=gf)
        b.molecules_unbatch_key = topo.molecules_unbatch_key
        b.topology = topo

        def conv(arrs):
            return [None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32), device=dev)
                    for a in (arrs or [])]
        b._device_tensors = [topo.buffer, af, bf, gf, af_pad, f_ini]      # for record_stream in the prefetcher
        return b, conv(labels), conv(weights)

    # ------------------------------------------------------------------ loss
    def _loss(self, outputs, labels, weights):
        """L2Loss / SparseSoftmaxCrossEntropy through _StandardLoss: (loss * w).mean()
        (models/losses.py:76-94, 262-297; torch_model.py:1275-1294)."""
        y, w = labels[0], weights[0]
        if self.mode == 'regression':
            per = (outputs - y.reshape(outputs.shape)) ** 2
        else:
            logits = outputs[1]
            if logits.dim() == 2:
                per = nn.functional.cross_entropy(logits, y.reshape(-1).long(), reduction='none')
            else:
                per = nn.functional.cross_entropy(logits.permute(0, 2, 1), y.reshape(logits.shape[:2]).long(),
                                                  reduction='none')
        if w.numel() == per.numel():
            w = w.reshape(per.shape)
        elif w.dim() < per.dim():
            w = w.reshape(tuple(w.shape) + (1,) * (per.dim() - w.dim()))
        return (per * w).mean()

    # ------------------------------------------------------------------ training / inference
    def fit(self, dataset, nb_epoch=10, max_checkpoints_to_keep=5, checkpoint_interval=1000, deterministic=False,
            restore=False, callbacks=[], all_losses=None):
        """TorchModel.fit (torch_model.py:288-343)."""
        return self.fit_generator(self.default_generator(dataset, epochs=nb_epoch, deterministic=deterministic),
                                  max_checkpoints_to_keep, checkpoint_interval, restore, callbacks=callbacks,
                                  all_losses=all_losses)

    def _prefetched(self, generator, depth):
        """Run ``_prepare_batch`` (uploads + f_ini assembly) for upcoming batches on a helper thread and a side
        stream while the GPU trains on the current one — the role the reference gives to DiskDataset's prefetch
        thread (data/datasets.py:1670-1693).  Tensors are handed to the training stream with an event and
        ``record_stream`` (they were allocated on the side stream)."""
        import queue
        import threading
        if self._prefetch_stream is None:
            self._prefetch_stream = torch.cuda.Stream(device=self.device)
        side, q, state = self._prefetch_stream, queue.Queue(maxsize=max(1, depth)), {"stop": False, "error": None}

        def run():
            try:
                torch.cuda.set_device(self.device)
                for batch in generator:
                    if state["stop"]:
                        break
                    with torch.cuda.stream(side):
                        prepared = self._prepare_batch(batch)
                        ev = torch.cuda.Event()
                        ev.record(side)
                    q.put((prepared, ev))
            except BaseException as e:      # surfaced in the consumer
                state["error"] = e
            finally:
                q.put(None)

        th = threading.Thread(target=run, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                (inputs, labels, weights), ev = item
                main = torch.cuda.current_stream(self.device)
                main.wait_event(ev)
                for t in list(getattr(inputs, "_device_tensors", [])) + [t for t in labels + weights if t is not None]:
                    t.record_stream(main)
                yield inputs, labels, weights
            if state["error"] is not None:
                raise state["error"]
        finally:
            state["stop"] = True
            while th.is_alive():
                try:
                    q.get_nowait()
                except Exception:
                    pass
                th.join(timeout=0.05)

    def fit_generator(self, generator, max_checkpoints_to_keep=5, checkpoint_interval=1000, restore=False,
                      callbacks=[], all_losses=None, prefetch=2):
        """TorchModel.fit_generator (torch_model.py:345-496): ``restore`` reloads the newest checkpoint before the first
        step, a checkpoint is written every ``checkpoint_interval`` steps and at the end (``model_dir`` set and
        interval > 0; rank 0 only in a data-parallel run), ``callbacks`` are called as f(model, step) after every step,
        ``all_losses`` collects the average loss of every ``log_frequency`` steps.  Returns the last average."""
        if not isinstance(callbacks, (list, tuple)):
            callbacks = [callbacks]
        self.model.train()
        t0 = time.time()
        if restore:
            self.restore()
        avg_sum, avg_n, last_avg = 0.0, 0, 0.0
        prepared_iter = self._prefetched(generator, prefetch) if (prefetch and self.device.type == "cuda") else \
            (self._prepare_batch(b) for b in generator)
        for inputs, labels, weights in prepared_iter:
            self._train_step(inputs, labels, weights)
            self._global_step += 1
            step = self._global_step
            want_loss = all_losses is not None or step % self.log_frequency == 0
            if want_loss:                       # a device->host read: only when somebody looks at it
                avg_sum += float(self._last_loss)
                avg_n += 1
            if step % self.log_frequency == 0 and avg_n:
                last_avg = avg_sum / avg_n
                logger.info('Ending global_step %d: Average loss %g' % (step, last_avg))
                if all_losses is not None:
                    all_losses.append(last_avg)
                avg_sum, avg_n = 0.0, 0
            if self.model_dir and checkpoint_interval > 0 and step % checkpoint_interval == checkpoint_interval - 1:
                self.save_checkpoint(max_checkpoints_to_keep)
            for c in callbacks:
                c(self, step)
        if avg_n:
            last_avg = avg_sum / avg_n
            if all_losses is not None:
                all_losses.append(last_avg)
        elif self._global_step and last_avg == 0.0 and getattr(self, "_last_loss", None) is not None:
            last_avg = float(self._last_loss)
        if self.model_dir and checkpoint_interval > 0:
            self.save_checkpoint(max_checkpoints_to_keep)
        logger.info("TIMING: model fitting took %0.3f s" % (time.time() - t0))
        self._check_f16_range()
        return last_avg

    def _check_f16_range(self):
        """The forward GEMMs run with fp16 operand halves (csrc/dmpnn_model.cu forward_mode): fail loudly if an operand
        has left the range they need (same check as GraphConvModel's)."""
        if self.device.type == "cuda":
            from .graphconvmodel import _check_f16_range
            _check_f16_range(self, any_path=True)

    def _train_step(self, inputs, labels, weights):
        """zero_grad, forward, loss, backward, (gradient all-reduce), Adam step (torch_model.py:435-443)."""
        eng = self._engine
        if eng is not None and labels and labels[0] is not None and getattr(inputs, "topology", None) is not None:
            topo = inputs.topology
            y = labels[0].reshape(topo.n_mols, -1).contiguous()
            w = weights[0] if (weights and weights[0] is not None) else None
            if w is not None:
                w = w.reshape(topo.n_mols, -1)
                w = (w.expand_as(y) if w.shape != y.shape else w).contiguous()
            loss = eng.train_step(topo, inputs['atom_features'], inputs['f_ini_atoms_bonds'], y, w)
            scale = 1.0
            if self._dp:
                from .parallel import world_size
                import torch.distributed as dist
                if world_size() > 1:     # the one exchange of a data-parallel step: the flat gradient slab
                    dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM)
                    scale = 1.0 / world_size()
            eng.adam_step(scale)
            self._last_loss = loss
            return loss
        slab = self._grad_slab
        if slab is None:
            self._pytorch_optimizer.zero_grad(set_to_none=True)
        else:
            slab.zero()
            slab.attach()
        loss = self._loss(self.model(inputs), labels, weights)
        loss.backward()
        if slab is not None:
            slab.collect()
            slab.all_reduce_mean()      # the one exchange of a data-parallel step (molecules never interact)
        self._pytorch_optimizer.step()
        self._last_loss = loss
        return loss

    def enable_data_parallel(self):
        """Average gradients over the default process group every step (one all-reduce of a flat slab; equal
        per-rank batches and a mean loss make the averaged gradient exact).  Parameters are broadcast from rank 0."""
        import torch.distributed as dist
        from .parallel import GradSlab, world_size
        if self._engine is not None:
            if world_size() > 1:
                dist.broadcast(self._engine.params, src=0)
            self._dp = True
            return self
        if world_size() > 1:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=0)
        self._grad_slab = GradSlab(self.model.parameters())
        return self

    def fit_on_batch(self, X, y, w):
        ds = _GraphDataset(X if isinstance(X, PackedGraphs) else PackedGraphs.from_graphs(list(X), self.atom_fdim,
                                                                                          self.bond_fdim), y, w)
        bs, self.batch_size = self.batch_size, max(self.batch_size, len(ds))
        try:
            return self.fit_generator(self.default_generator(ds, deterministic=True))
        finally:
            self.batch_size = bs

    def predict(self, dataset, transformers=[]):
        self.model.eval()
        outs = []
        with torch.no_grad():
            for batch in self.default_generator(dataset, mode='predict', deterministic=True):
                inputs, _, _ = self._prepare_batch(batch)
                if self._engine is not None:
                    o = self._engine.forward(inputs.topology, inputs['atom_features'], inputs['f_ini_atoms_bonds'])
                else:
                    o = self.model(inputs)
                outs.append((o if self.mode == 'regression' else o[0]).detach().cpu().numpy())
        self._check_f16_range()
        y = np.concatenate(outs, 0) if outs else np.zeros((0, self.n_tasks), np.float32)
        for t in reversed(list(transformers)):              # deepchem.trans.undo_transforms (torch_model.py:625-634)
            if getattr(t, "transform_y", False):
                y = t.untransform(y)
        return y

    def predict_on_batch(self, X):
        ds = _GraphDataset(X if isinstance(X, PackedGraphs) else PackedGraphs.from_graphs(list(X), self.atom_fdim,
                                                                                          self.bond_fdim))
        bs, self.batch_size = self.batch_size, max(self.batch_size, len(ds))
        try:
            return self.predict(ds)
        finally:
            self.batch_size = bs

    # ------------------------------------------------------------------ checkpoints (torch_model.py:996-1090)
    def save_checkpoint(self, max_checkpoints_to_keep=5, model_dir=None):
        model_dir = model_dir or self.model_dir
        if model_dir is None:
            raise ValueError("model_dir is not set")
        if self._dp:
            from .parallel import rank
            if rank() != 0:                      # replicas are identical: one writer
                return
        os.makedirs(model_dir, exist_ok=True)
        paths = [os.path.join(model_dir, 'checkpoint%d.pt' % (i + 1)) for i in range(max_checkpoints_to_keep)]
        tmp = os.path.join(model_dir, 'temp_checkpoint.pt')
        torch.save({'model_state_dict': self.model.state_dict(),
                    'optimizer_state_dict': (self._engine.state_dict() if self._engine is not None
                                             else self._pytorch_optimizer.state_dict()),
                    'global_step': self._global_step}, tmp)
        if os.path.exists(paths[-1]):
            os.remove(paths[-1])
        for i in reversed(range(max_checkpoints_to_keep - 1)):
            if os.path.exists(paths[i]):
                os.rename(paths[i], paths[i + 1])
        os.rename(tmp, paths[0])

    def restore(self, checkpoint=None, model_dir=None):
        model_dir = model_dir or self.model_dir
        if checkpoint is None:
            checkpoint = os.path.join(model_dir, 'checkpoint1.pt')
            if not os.path.exists(checkpoint):
                raise ValueError('No checkpoint found')
        data = torch.load(checkpoint, map_location=self.device)
        self.model.load_state_dict(data['model_state_dict'])
        if self._engine is not None:
            self._engine.load_state_dict(data['optimizer_state_dict'])
        else:
            self._pytorch_optimizer.load_state_dict(data['optimizer_state_dict'])
        self._global_step = data['global_step']

    def get_checkpoints(self, model_dir=None):
        """Checkpoint files, newest first (torch_model.py:1044-1059)."""
        model_dir = model_dir or self.model_dir
        if not model_dir or not os.path.isdir(model_dir):
            return []
        files = [f for f in os.listdir(model_dir) if f.startswith("checkpoint") and f.endswith(".pt")]
        files.sort(key=lambda f: int(f[len("checkpoint"):-3]))
        return [os.path.join(model_dir, f) for f in files]

    def get_global_step(self):
        return self._global_step

    def evaluate(self, dataset, metrics, transformers=[], per_task_metrics=False, use_sample_weights=False,
                 n_classes=2):
        """Model.evaluate (models.py:191-236) through the Evaluator's logic: labels and predictions both have the
        y-transformers undone (deepchem/utils/evaluate.py:303-307)."""
        from .graphconvmodel import evaluate_model
        return evaluate_model(self, dataset, metrics, transformers, per_task_metrics, use_sample_weights, n_classes)
