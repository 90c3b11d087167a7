"""Oracle (torch, CPU) for GraphConv / GraphPool / GraphGather and the GraphConv model.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the reference math with its gradient path intact:
  * GraphConv      deepchem/models/torch_models/layers.py:6167-6246
                   (the reference severs autograd with ``.detach().numpy()`` at :6216,
                   :6226, :6244; the differentiable semantics are those of the Keras
                   original, deepchem/models/layers.py:151-213)
  * GraphPool      deepchem/models/torch_models/layers.py:6319-6367
  * GraphGather    deepchem/models/torch_models/layers.py:6450-6479 with
                   unsorted_segment_sum (deepchem/utils/pytorch_utils.py:20-74) and
                   unsorted_segment_max (deepchem/utils/pytorch_utils.py:473-528)
  * model graph    deepchem/models/torch_models/graphconvmodel.py:142-249 (widths generic as
                   in the Keras model, deepchem/models/graph_models.py:835-902)
  * losses         deepchem/models/losses.py:76-94 (L2), :236-259 (softmax CE) reduced as
                   _StandardLoss does (torch_models/torch_model.py:1267-1294)

Works in fp32 (parity with the reference's libtorch kernels) and fp64 (gradcheck).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _counts(deg_slice):
    return [int(c) for c in deg_slice[:, 1].tolist()]


def graph_conv(x, deg_slice, deg_adj_lists, W_list, b_list, activation=None,
               min_deg=0, max_deg=10):
    """Y_d = S_d W[2(d-1)] + b[2(d-1)] + X_d W[2(d-1)+1] + b[2(d-1)+1]; Y_0 = X_0 W[20] + b[20].

    ``deg_adj_lists`` holds the lists for degree 1..max_deg (degree-0 list dropped, as in
    the model inputs).  Weight order follows layers.py:6189-6226.
    """
    blocks = torch.split(x, _counts(deg_slice))
    out = []
    k = 0
    for d in range(1, max_deg + 1):
        adj = deg_adj_lists[d - 1].long()
        nbr_sum = x[adj].sum(1)                                  # layers.py:6241-6243
        rel = nbr_sum @ W_list[k] + b_list[k]
        slf = blocks[d - min_deg] @ W_list[k + 1] + b_list[k + 1]
        out.append(rel + slf)
        k += 2
    if min_deg == 0:
        out.insert(0, blocks[0] @ W_list[k] + b_list[k])
    y = torch.cat(out, 0)
    if activation is not None:
        y = activation(y)
    return y


def graph_pool(x, deg_slice, deg_adj_lists, min_deg=0, max_deg=10):
    """max over [self, neighbours]; degree-0 rows pass through (layers.py:6342-6367).
    torch.max over dim 1 routes the gradient to the first slot that attains the max."""
    blocks = torch.split(x, _counts(deg_slice))
    out = []
    for d in range(1, max_deg + 1):
        adj = deg_adj_lists[d - 1].long()
        if adj.shape[0] == 0:
            out.append(x.new_zeros((0, x.shape[-1])))
            continue
        stacked = torch.cat([blocks[d - min_deg].unsqueeze(1), x[adj]], 1)
        out.append(torch.max(stacked, 1)[0])
    if min_deg == 0:
        out.insert(0, blocks[0])
    return torch.cat(out, 0)


def segment_sum(x, ids, num_segments):
    """pytorch_utils.py:20-74 (scatter_add into zeros)."""
    idx = ids.long().unsqueeze(-1).expand(-1, x.shape[1])
    return torch.zeros(num_segments, x.shape[1], dtype=x.dtype).scatter_add(0, idx, x)


def segment_max(x, ids, num_segments):
    """pytorch_utils.py:473-528: -inf for empty segments, ties to the lowest row index.

    The reference loops over segments with a full-tensor masked max (O(B*N*F)); this is
    the same function computed through a padded [B, Lmax, F] view so that it scales.
    """
    ids = ids.long()
    n = x.shape[0]
    if n == 0:
        return x.new_full((num_segments, x.shape[1]), -math.inf)
    order = torch.argsort(ids, stable=True)
    sorted_ids = ids[order]
    counts = torch.bincount(ids, minlength=num_segments)
    starts = torch.cumsum(counts, 0) - counts
    pos = torch.arange(n) - starts[sorted_ids]
    lmax = int(counts.max().item())
    padded = x.new_full((num_segments, max(lmax, 1), x.shape[1]), -math.inf)
    padded = padded.index_put((sorted_ids, pos), x[order])
    return torch.max(padded, 1)[0]


def graph_gather(x, membership, batch_size, activation=None):
    """[segment_sum | segment_max] then activation (layers.py:6464-6479)."""
    assert batch_size > 1, "graph_gather requires batches larger than 1"
    z = torch.cat([segment_sum(x, membership, batch_size),
                   segment_max(x, membership, batch_size)], 1)
    if activation is not None:
        z = activation(z)
    return z


class OracleGraphConvLayer(nn.Module):
    """Parameter container with the reference's names/shapes (layers.py:6139-6151)."""

    def __init__(self, out_channel, number_input_features, min_deg=0, max_deg=10,
                 activation_fn=None):
        super().__init__()
        self.min_degree, self.max_degree = min_deg, max_deg
        self.activation_fn = activation_fn
        n = 2 * max_deg + (1 - min_deg)
        self.W_list = nn.ParameterList([
            nn.Parameter(nn.init.xavier_uniform_(torch.empty(number_input_features, out_channel)))
            for _ in range(n)])
        self.b_list = nn.ParameterList([nn.Parameter(torch.zeros(out_channel)) for _ in range(n)])

    def forward(self, inputs):
        return graph_conv(inputs[0], inputs[1], inputs[3:], list(self.W_list), list(self.b_list),
                          self.activation_fn, self.min_degree, self.max_degree)


class OracleGraphConvModel(nn.Module):
    """graphconvmodel.py:77-249 with generic widths.  state_dict keys equal the reference's."""

    def __init__(self, n_tasks, graph_conv_layers=(64, 64), dense_layer_size=128, dropout=0.0,
                 mode="classification", number_atom_features=75, n_classes=2,
                 batch_normalize=True, uncertainty=False, batch_size=100):
        super().__init__()
        if mode not in ("classification", "regression"):
            raise ValueError("mode must be either 'classification' or 'regression'")
        graph_conv_layers = list(graph_conv_layers)
        self.n_tasks, self.n_classes, self.mode = n_tasks, n_classes, mode
        self.uncertainty = uncertainty
        if not isinstance(dropout, (list, tuple)):
            dropout = [dropout] * (len(graph_conv_layers) + 1)
        if len(dropout) != len(graph_conv_layers) + 1:
            raise ValueError("Wrong number of dropout probabilities provided")
        widths_in = [number_atom_features] + graph_conv_layers[:-1]
        self.graph_convs = nn.ModuleList([
            OracleGraphConvLayer(c, f, activation_fn=F.relu)
            for c, f in zip(graph_conv_layers, widths_in)])

        def bn(c):
            return nn.BatchNorm1d(c, eps=1e-3, momentum=0.99) if batch_normalize else nn.Identity()
        self.batch_norms = nn.ModuleList([bn(c) for c in graph_conv_layers] + [bn(dense_layer_size)])
        self.dropouts = nn.ModuleList([nn.Dropout(r) if r > 0.0 else nn.Identity() for r in dropout])
        self.dense = nn.Linear(graph_conv_layers[-1], dense_layer_size)
        self.batch_size = batch_size
        if mode == "classification":
            self.reshape_dense = nn.Linear(dense_layer_size * 2, n_tasks * n_classes)
        else:
            self.regression_dense = nn.Linear(dense_layer_size * 2, n_tasks)
            if uncertainty:
                self.uncertainty_dense = nn.Linear(dense_layer_size * 2, n_tasks)

    def forward(self, inputs, training=False):
        x, deg_slice, membership = inputs[0], inputs[1], inputs[2].long()
        n_samples = int(inputs[3])
        adjs = [a.long() for a in inputs[4:]]
        h = x
        for i, conv in enumerate(self.graph_convs):
            h = conv([h, deg_slice, membership] + adjs)
            h = self.batch_norms[i](h)
            if training:
                h = self.dropouts[i](h)
            h = graph_pool(h, deg_slice, adjs)
        h = self.batch_norms[-1](F.relu(self.dense(h)))
        if training:
            h = self.dropouts[-1](h)
        fp = graph_gather(h, membership, self.batch_size, torch.tanh)
        if self.mode == "classification":
            logits = self.reshape_dense(fp).reshape(-1, self.n_tasks, self.n_classes)[:n_samples]
            return [F.softmax(logits, dim=2), logits, fp]
        out = self.regression_dense(fp)[:n_samples]
        if self.uncertainty:
            log_var = self.uncertainty_dense(fp)[:n_samples]
            return [out, torch.exp(log_var), out, log_var, fp]
        return [out, fp]


def standard_loss(mode, outputs, y, w, uncertainty=False):
    """Scalar training loss as TorchModel computes it (torch_model.py:439-441, 1275-1294):
    elementwise criterion times broadcast weights, mean over all elements.
    classification: -sum_c y*log_softmax(logits) per (sample, task)  (losses.py:236-259)
    regression:     (out - y)^2                                       (losses.py:76-94)
    """
    if mode == "classification":
        per = -(y * F.log_softmax(outputs[1], dim=-1)).sum(-1)
    elif uncertainty:
        per = (outputs[0] - y) ** 2 / torch.exp(outputs[3]) + outputs[3]
    else:
        per = (outputs[0] - y.reshape(outputs[0].shape)) ** 2
    while w.dim() < per.dim():
        w = w.unsqueeze(-1)
    return (per * w).mean()
