"""CPU oracle of the D-MPNN edge message-passing path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates, in numpy (integer tables) and torch-CPU (layer math):
  * ``_MapperDMPNN``            deepchem/models/torch_models/dmpnn.py:123-243
  * ``_ModData.__inc__`` + ``DMPNNModel.default_generator`` padding + PyG ``Batch.from_data_list``
                                 dmpnn.py:17-35, 677-755 (torch_geometric is an un-vendored, unpinned
                                 dependency: requirements/torch/env_torch.cpu.yml).  The collation is
                                 restated from ``__inc__`` (every entry of ``atom_to_incoming_bonds`` /
                                 ``mapping``, INCLUDING the -1 pads, is shifted by the cumulative number
                                 of f_ini rows of the preceding molecules); the reference's own tests pin
                                 only batch-size-1 outputs, so multi-molecule collation is "parity
                                 unpinned" beyond that derivation.
  * ``DMPNNEncoderLayer``        deepchem/models/torch_models/layers.py:1436-1649 (literal loop: the
                                 gather always reads ``message``, so only the last W_h product is live)
  * ``PositionwiseFeedForward``  layers.py:795-910,  ``DMPNN``  dmpnn.py:246-449

Pinned by tests/test_oracle_dmpnn.py against fixtures generated from the reference itself
(tests/golden/make_golden_dmpnn.py -> tests/golden/ref_dmpnn.npz) and the reference's known answers
(models/tests/test_mapper_dmpnn.py:17-111, models/tests/test_layers.py:798-827).
"""
import numpy as np
import torch
import torch.nn as nn


class OracleGraph(object):
    """The fields of deepchem.feat.GraphData the path reads (feat/graph_data.py)."""

    def __init__(self, node_features, edge_index, edge_features=None, global_features=None):
        self.node_features = np.asarray(node_features)
        self.edge_index = np.asarray(edge_index).reshape(2, -1).astype(np.int64)
        self.edge_features = None if edge_features is None else np.asarray(edge_features)
        self.global_features = np.empty(0) if global_features is None else np.asarray(global_features)
        self.num_nodes = self.node_features.shape[0]
        self.num_node_features = self.node_features.shape[1]
        self.num_edges = self.edge_index.shape[1]
        self.num_edge_features = 0 if self.edge_features is None else self.edge_features.shape[1]


def mapper_values(graph):
    """(atom_features, f_ini_atoms_bonds, atom_to_incoming_bonds, mapping, global_features) of one
    molecule (dmpnn.py:123-243)."""
    n_atoms, n_bonds = graph.num_nodes, graph.num_edges
    if n_bonds == 0:                                                    # dmpnn.py:154-161
        f_ini = np.zeros((1, graph.num_node_features + graph.num_edge_features))
        a2b = np.asarray([[-1]] * n_atoms, dtype=int)
        mapping = np.asarray([[-1]], dtype=int)
        return graph.node_features, f_ini, a2b, mapping, graph.global_features
    src, dst = graph.edge_index[0], graph.edge_index[1]
    f_ini = np.hstack((graph.node_features[src], graph.edge_features))  # :183-184
    f_ini = np.pad(f_ini, ((0, 1), (0, 0)))                             # :187-188
    lists = [list(np.where(dst == i)[0]) for i in range(n_atoms)]       # :217-219
    k = max(1, max(len(l) for l in lists))                              # :222-223
    a2b = np.asarray([l + [-1] * (k - len(l)) for l in lists], dtype=int)
    mapping = a2b[src].copy()                                           # :203
    for b in range(n_bonds):                                            # :234-243, reverse bond = b ^ 1
        rev = b + 1 if b % 2 == 0 else b - 1
        mapping[b][mapping[b] == rev] = -1
    mapping = np.pad(mapping, ((0, 1), (0, 0)), constant_values=-1)     # :207-208
    return graph.node_features, f_ini, a2b, mapping, graph.global_features


def collate(values_list):
    """Batch of mapper outputs -> the tensors DMPNN.forward reads from the PyG batch (dmpnn.py:425-441).

    Padding to the batch-wide maximum in-degree with -1 (dmpnn.py:741-753), then concatenation with
    the per-molecule increment ``len(f_ini)`` added to EVERY entry of the two index tables
    (dmpnn.py:27-35)."""
    k = max(1, max(v[2].shape[1] for v in values_list))
    atoms, f_inis, a2bs, maps, globs, key = [], [], [], [], [], []
    inc = 0
    for af, f_ini, a2b, mapping, gf in values_list:
        pad = k - a2b.shape[1]
        a2b = np.pad(a2b, ((0, 0), (0, pad)), constant_values=-1)
        mapping = np.pad(mapping, ((0, 0), (0, pad)), constant_values=-1)
        atoms.append(af)
        f_inis.append(f_ini)
        a2bs.append(a2b + inc)
        maps.append(mapping + inc)
        globs.append(np.asarray(gf).reshape(-1))
        key.append(af.shape[0])
        inc += f_ini.shape[0]
    return (np.concatenate(atoms, 0), np.concatenate(f_inis, 0), np.concatenate(a2bs, 0),
            np.concatenate(maps, 0), np.concatenate(globs, 0), key)


_ACT = {'relu': nn.ReLU, 'leakyrelu': lambda: nn.LeakyReLU(0.1), 'prelu': nn.PReLU, 'tanh': nn.Tanh,
        'selu': nn.SELU, 'elu': nn.ELU}


class OracleDMPNNEncoder(nn.Module):
    """layers.py:1436-1649, same parameter names (W_i, W_h, W_o) and forward signature."""

    def __init__(self, atom_fdim=133, bond_fdim=14, d_hidden=300, depth=3, bias=False, activation='relu',
                 aggregation='mean', aggregation_norm=100):
        super(OracleDMPNNEncoder, self).__init__()
        self.atom_fdim, self.concat_fdim = atom_fdim, atom_fdim + bond_fdim
        self.depth, self.aggregation, self.aggregation_norm = depth, aggregation, aggregation_norm
        self.activation = _ACT[activation]()
        self.W_i = nn.Linear(self.concat_fdim, d_hidden, bias=bias)
        self.W_h = nn.Linear(d_hidden, d_hidden, bias=bias)
        self.W_o = nn.Linear(self.atom_fdim + d_hidden, d_hidden)

    def forward(self, atom_features, f_ini_atoms_bonds, atom_to_incoming_bonds, mapping, global_features,
                molecules_unbatch_key):
        inp = self.W_i(f_ini_atoms_bonds)                                # :1622
        message = self.activation(inp)                                   # :1624
        for _ in range(1, self.depth):                                   # :1627-1633
            message = message[mapping].sum(1)
            h_message = self.activation(inp + self.W_h(message))
        m2a = h_message[atom_to_incoming_bonds].sum(1)                   # :1539
        atoms_hidden = self.activation(self.W_o(torch.cat((atom_features, m2a), 1)))   # :1541-1545
        vecs = []
        for mol in torch.split(atoms_hidden, list(molecules_unbatch_key)):   # :1571-1583
            if self.aggregation == 'mean':
                vecs.append(mol.sum(dim=0) / len(mol))
            elif self.aggregation == 'sum':
                vecs.append(mol.sum(dim=0))
            elif self.aggregation == 'norm':
                vecs.append(mol.sum(dim=0) / self.aggregation_norm)
            else:
                raise Exception("Invalid aggregation")
        out = torch.stack(vecs, dim=0)
        if global_features.size()[0] != 0:                               # :1644-1647
            if len(global_features.shape) == 1:
                global_features = global_features.view(len(out), -1)
            out = torch.cat([out, global_features], dim=1)
        return out


class OracleFFN(nn.Module):
    """PositionwiseFeedForward (layers.py:795-910) with dropout 0."""

    def __init__(self, d_input, d_hidden, d_output, activation='relu', n_layers=3):
        super(OracleFFN, self).__init__()
        self.activation = (lambda x: x) if activation == 'linear' else _ACT[activation]()
        self.n_layers = n_layers
        d_output = d_output if d_output != 0 else d_input
        d_hidden = d_hidden if d_hidden != 0 else d_input
        if n_layers == 1:
            lin = [nn.Linear(d_input, d_output)]
        else:
            lin = [nn.Linear(d_input, d_hidden)] + [nn.Linear(d_hidden, d_hidden) for _ in range(n_layers - 2)] + \
                [nn.Linear(d_hidden, d_output)]
        self.linears = nn.ModuleList(lin)

    def forward(self, x):
        if not self.n_layers:
            return x
        if self.n_layers == 1:
            return self.linears[0](x)          # dropout_at_input_no_act=True, p=0 (dmpnn.py:300)
        for i in range(self.n_layers - 1):
            x = self.activation(self.linears[i](x))
        return self.linears[-1](x)


class OracleDMPNN(nn.Module):
    """dmpnn.py:246-449 (encoder + ffn, regression or classification head)."""

    def __init__(self, mode='regression', n_classes=3, n_tasks=1, global_features_size=0, atom_fdim=133,
                 bond_fdim=14, enc_hidden=300, depth=3, bias=False, enc_activation='relu', aggregation='mean',
                 aggregation_norm=100, ffn_hidden=300, ffn_activation='relu', ffn_layers=3):
        super(OracleDMPNN, self).__init__()
        self.mode, self.n_classes, self.n_tasks = mode, n_classes, n_tasks
        self.encoder = OracleDMPNNEncoder(atom_fdim, bond_fdim, enc_hidden, depth, bias, enc_activation,
                                          aggregation, aggregation_norm)
        out = n_tasks if mode == 'regression' else n_tasks * n_classes
        self.ffn = OracleFFN(enc_hidden + global_features_size, ffn_hidden, out, ffn_activation, ffn_layers)

    def forward(self, batch):
        af, f_ini, a2b, mapping, gf, key = batch
        enc = self.encoder(af, f_ini, a2b, mapping, gf, key)
        out = self.ffn(enc)
        if self.mode == 'regression':
            return out
        if self.n_tasks == 1:
            out = out.view(-1, self.n_classes)
            return torch.softmax(out, dim=1), out
        out = out.view(-1, self.n_tasks, self.n_classes)
        return torch.softmax(out, dim=2), out


def to_torch_batch(collated, dtype=torch.float32):
    af, f_ini, a2b, mapping, gf, key = collated
    return (torch.from_numpy(np.asarray(af)).to(dtype), torch.from_numpy(np.asarray(f_ini)).to(dtype),
            torch.from_numpy(a2b).long(), torch.from_numpy(mapping).long(),
            torch.from_numpy(np.asarray(gf)).to(dtype), key)
