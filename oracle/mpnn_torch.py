"""CPU oracle of the MPNN edge-network message passing (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates in torch-CPU, literally (same operations in the same order), the forward passes of
  * ``EdgeNetwork``          deepchem/models/torch_models/layers.py:4006-4088  (Keras: models/layers.py:3712-3753)
                              + ``segment_sum`` deepchem/utils/pytorch_utils.py:77-123
  * ``GatedRecurrentUnit``   torch_models/layers.py:2884-2919                   (Keras: models/layers.py:3755-3800)
  * ``MessagePassing``       models/layers.py:3648-3710 (Keras only: pad to n_hidden, T x (EdgeNetwork, GRU))
  * ``SetGather``            torch_models/layers.py:2976-3138                   (Keras: models/layers.py:3802-3887)
The torch port of these layers is forward-only (plain-tensor weights, ``.detach()`` in the LSTM step), so is this.

``SetGather`` calls ``torch_geometric.utils.scatter`` (torch_geometric: un-vendored, unpinned,
requirements/torch/env_torch.cpu.yml); its documented default (``reduce='sum'`` along ``dim=0``, output rows =
``index.max() + 1``) is restated as ``scatter_sum`` below.

Pinned by tests/test_oracle_mpnn.py against
  * the reference's own known answer for SetGather (models/tests/test_layers.py:998-1016: assets
    atom_feat_SetGather.npy, weights_SetGather_tf.npy, result_SetGather_tf.npy, stored in tests/golden/ref_mpnn.npz);
  * outputs of the reference layers run in the build container on seeded inputs (tests/golden/make_golden_mpnn.py).
The EdgeNetwork asset (edgenetwork_result.npy, test_layers.py:1144-1206) needs WeaveFeaturizer output of 'CCC'
(RDKit) and is not reproducible here; EdgeNetwork is pinned by the reference-generated outputs only.
"""
import numpy as np
import torch


def scatter_sum(src, index, n_rows=None):
    """torch_geometric.utils.scatter(src, index, dim=0) with the default reduce='sum'."""
    index = torch.as_tensor(index).long()
    n = int(index.max()) + 1 if n_rows is None else n_rows
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


def segment_sum(data, segment_ids):
    """pytorch_utils.py:77-123: sorted ids; the number of segments is the number of DISTINCT ids."""
    ids = torch.as_tensor(segment_ids).long()
    assert bool((ids[1:] >= ids[:-1]).all()), "elements of segment_ids must be sorted"
    num_segments = len(torch.unique(ids))
    out = torch.zeros((num_segments,) + tuple(data.shape[1:]), dtype=data.dtype)
    return out.scatter_add(0, ids.view(-1, *([1] * (data.dim() - 1))).expand_as(data), data)


def edge_network(pair_features, atom_features, atom_to_pair, W, b, n_hidden):
    """layers.py:4068-4087."""
    A = torch.add(torch.matmul(pair_features, W), b)
    A = torch.reshape(A, (-1, n_hidden, n_hidden))
    out = torch.unsqueeze(atom_features[atom_to_pair[:, 1]], dim=2)
    out_squeeze = torch.squeeze(torch.matmul(A, out), dim=2)
    return segment_sum(out_squeeze, atom_to_pair[:, 0])


def gated_recurrent_unit(h_tm1, x, Wz, Wr, Wh, Uz, Ur, Uh, bz, br, bh):
    """layers.py:2907-2919 (note the last term: ``z * x``, the MESSAGE, exactly as written there)."""
    z = torch.sigmoid(torch.matmul(x, Wz) + torch.matmul(h_tm1, Uz) + bz)
    r = torch.sigmoid(torch.matmul(x, Wr) + torch.matmul(h_tm1, Ur) + br)
    return (1 - z) * torch.tanh(torch.matmul(x, Wh) + torch.matmul(h_tm1 * r, Uh) + bh) + z * x


def message_passing(atom_features, pair_features, atom_to_pair, T, n_hidden, enn, gru):
    """models/layers.py:3692-3710.  enn = (W, b), gru = (Wz, Wr, Wh, Uz, Ur, Uh, bz, br, bh)."""
    n_feat = atom_features.shape[-1]
    if n_feat < n_hidden:
        out = torch.nn.functional.pad(atom_features, (0, n_hidden - n_feat))
    elif n_feat > n_hidden:
        raise ValueError("Too large initial feature vector")
    else:
        out = atom_features
    for _ in range(T):
        message = edge_network(pair_features, out, atom_to_pair, enn[0], enn[1], n_hidden)
        out = gated_recurrent_unit(out, message, *gru)
    return out


def lstm_step(h, c, U, b, n_hidden):
    """layers.py:3081-3108."""
    z = torch.nn.functional.linear(h.float(), U.float().T, b)
    i = torch.sigmoid(z[:, :n_hidden])
    f = torch.sigmoid(z[:, n_hidden:2 * n_hidden])
    o = torch.sigmoid(z[:, 2 * n_hidden:3 * n_hidden])
    z3 = z[:, 3 * n_hidden:]
    c_out = f * c + i * torch.tanh(z3)
    return o * torch.tanh(c_out), c_out


def set_gather(atom_features, atom_split, U, b, M, batch_size, n_hidden):
    """layers.py:3041-3079.  atom_features: float array [N, n_hidden]; atom_split: int array [N]."""
    x = torch.as_tensor(np.asarray(atom_features))
    split = np.asarray(atom_split)
    c = torch.zeros((batch_size, n_hidden))
    h = torch.zeros((batch_size, n_hidden))
    q_star = None
    for _ in range(M):
        q_expanded = h[torch.as_tensor(split).long()]
        e = (x * q_expanded).sum(dim=-1)
        e_mols = [e[torch.as_tensor(split == i)] for i in range(batch_size)]
        e_mols = [torch.cat([e_mol, torch.tensor([-1000.], dtype=e.dtype)], dim=0) for e_mol in e_mols]
        a = torch.cat([torch.nn.functional.softmax(e_mol[:-1], dim=0) for e_mol in e_mols], dim=0)
        # the reference concatenates the per-molecule attention weights in MOLECULE order and multiplies them with
        # the atoms in ATOM order: identical only when atom_split is sorted, which its callers guarantee
        r = scatter_sum(torch.reshape(a, [-1, 1]) * x, split)
        q_star = torch.cat([h, r], dim=1)
        h, c = lstm_step(q_star, c, U, b, n_hidden)
    return q_star


# ---- differentiable, dtype-generic restatements (the Keras originals train: models/layers.py:3648-3887,
# graph_models.py:1045-1247).  Used in float64 as the gradient oracle of the CUDA backward kernels.
def lstm_step_d(h, c, U, b, n_hidden):
    """lstm_step without the float32 casts (layers.py:3081-3108)."""
    z = h @ U + b
    i = torch.sigmoid(z[:, :n_hidden])
    f = torch.sigmoid(z[:, n_hidden:2 * n_hidden])
    o = torch.sigmoid(z[:, 2 * n_hidden:3 * n_hidden])
    c_out = f * c + i * torch.tanh(z[:, 3 * n_hidden:])
    return o * torch.tanh(c_out), c_out


def set_gather_d(x, atom_split, U, b, M, batch_size, n_hidden):
    """set_gather with tensors of any floating dtype, differentiable end to end (Keras semantics: the gradient flows
    through the LSTM state; the torch port detaches it).  atom_split must be sorted, as its callers guarantee."""
    split = torch.as_tensor(np.asarray(atom_split)).long()
    c = x.new_zeros((batch_size, n_hidden))
    h = x.new_zeros((batch_size, n_hidden))
    q_star = None
    for _ in range(M):
        e = (x * h[split]).sum(dim=-1)
        m = torch.full((batch_size,), -float("inf"), dtype=x.dtype).scatter_reduce(0, split, e.detach(), "amax")
        ex = torch.exp(e - m[split])
        den = x.new_zeros((batch_size,)).index_add(0, split, ex)
        a = ex / den[split]
        r = x.new_zeros((batch_size, n_hidden)).index_add(0, split, a.unsqueeze(1) * x)
        q_star = torch.cat([h, r], dim=1)
        h, c = lstm_step_d(q_star, c, U, b, n_hidden)
    return q_star


def mpnn_model(params, inputs, T, M, n_hidden, batch_size, mode="regression", n_tasks=1, n_classes=2):
    """The MPNNModel network (graph_models.py:1121-1160): MessagePassing -> Dense -> SetGather -> Dense(relu) -> Dense.
    params: dict with enn (W, b), gru (9 tensors), atom_dense (kernel, bias), set_gather (U, b), dense1, out."""
    atom_features, pair_features, atom_split, atom_to_pair, n_samples = inputs
    a2p = torch.as_tensor(np.asarray(atom_to_pair)).long()
    h = message_passing(atom_features, pair_features, a2p, T, n_hidden, params["enn"], params["gru"])
    emb = h @ params["atom_dense"][0] + params["atom_dense"][1]
    mol = set_gather_d(emb, atom_split, params["set_gather"][0], params["set_gather"][1], M, batch_size, n_hidden)
    d1 = torch.relu(mol @ params["dense1"][0] + params["dense1"][1])
    out = d1 @ params["out"][0] + params["out"][1]
    if mode == "classification":
        logits = out.reshape(-1, n_tasks, n_classes)[:int(n_samples)]
        return [torch.softmax(logits, dim=2), logits]
    return [out[:int(n_samples)]]
