"""Drop-in GraphConv / GraphPool / GraphGather layers.

Same constructor arguments, parameter names (``W_list`` / ``b_list``, 21 each) and
``forward(inputs: List[Tensor])`` contract as the reference torch layers
(deepchem/models/torch_models/layers.py:6061-6479), with the math executed by the sm_100a
kernels of libdcgc and a real gradient path (the reference detaches GraphConv's output,
layers.py:6216/6226/6244; the differentiable semantics are the Keras original's,
deepchem/models/layers.py:151-213).

``inputs = [atom_features, deg_slice, membership, deg_adj_1, ..., deg_adj_10]``.  The batch
topology (CSR, CSR^T, molecule CSR, GEMM tiles) is taken from the DeviceTopology attached to
``deg_slice`` by the batch generator, or derived once from the plain tensors.
"""
from typing import Callable, List, Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import ACT_NONE, GEMM_BF16, GEMM_FP32, GEMM_TF32X3
from .mol_graphs import topology_of

_GEMM_MODES = {"fp32": GEMM_FP32, "bf16": GEMM_BF16, "tf32x3": GEMM_TF32X3}


def gemm_mode_code(mode):
    if mode in _GEMM_MODES:
        return _GEMM_MODES[mode]
    if mode in _GEMM_MODES.values():
        return mode
    raise ValueError("gemm_mode must be one of %s" % sorted(_GEMM_MODES))


class GraphConv(nn.Module):
    """Graph convolution of Duvenaud et al. over a degree-bucketed ConvMol batch.

    Reference: torch_models/layers.py:6104-6246.  ``gemm_mode`` selects the arithmetic of the
    degree-grouped contraction: 'tf32x3' (default: tcgen05 tensor cores, three-term TF32 split, fp32-grade results inside the
    1e-5 parity bar), 'fp32' (SIMT FFMA) or 'bf16' (tensor cores, the 2e-2 mode).
    """

    def __init__(self, out_channel: int, number_input_features: int, min_deg: int = 0, max_deg: int = 10,
                 activation_fn: Optional[Callable] = None, gemm_mode: str = "tf32x3", **kwargs):
        super(GraphConv, self).__init__(**kwargs)
        if min_deg != 0 or max_deg != 10:
            raise ValueError("the B200 GraphConv kernels are built for degrees 0..10")
        self.out_channel = out_channel
        self.min_degree = min_deg
        self.max_degree = max_deg
        self.number_input_features = number_input_features
        self.activation_fn = activation_fn
        self.gemm_mode = gemm_mode_code(gemm_mode)
        num_deg = 2 * max_deg + (1 - min_deg)
        self.W_list = nn.ParameterList([
            nn.Parameter(nn.init.xavier_uniform_(torch.empty(number_input_features, out_channel)))
            for _ in range(num_deg)])
        self.b_list = nn.ParameterList([nn.Parameter(torch.zeros(out_channel)) for _ in range(num_deg)])
        self.built = True

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}(out_channel:{self.out_channel},min_deg:{self.min_degree},'
                f'max_deg:{self.max_degree},activation_fn:{self.activation_fn})')

    def forward(self, inputs: List[torch.Tensor]) -> torch.Tensor:
        x = inputs[0]
        topo = topology_of(inputs)
        fp = (self.number_input_features + 3) // 4 * 4
        w, b = ops.pack_graphconv_weights(list(self.W_list), list(self.b_list), fp)
        code = ops.act_code(self.activation_fn)
        y = ops.GraphConvFn.apply(x, w, b, topo, ACT_NONE if code is None else code, self.gemm_mode)
        if code is None:
            y = self.activation_fn(y)
        return y

    def sum_neigh(self, atoms: torch.Tensor, inputs: List[torch.Tensor]) -> torch.Tensor:
        """Neighbour sums for every row (degree-0 rows are zero), differentiable."""
        return ops.NeighborSum.apply(atoms, topology_of(inputs))


class GraphPool(nn.Module):
    """Max over each atom and its neighbours (torch_models/layers.py:6288-6367)."""

    def __init__(self, min_degree: int = 0, max_degree: int = 10, **kwargs):
        super(GraphPool, self).__init__(**kwargs)
        if min_degree != 0 or max_degree != 10:
            raise ValueError("the B200 GraphPool kernels are built for degrees 0..10")
        self.min_degree = min_degree
        self.max_degree = max_degree

    def get_config(self) -> str:
        return f'{self.__class__.__name__}(min_degree:{self.min_degree},max_degree:{self.max_degree})'

    def forward(self, inputs: List[torch.Tensor]) -> torch.Tensor:
        return ops.GraphPoolFn.apply(inputs[0], topology_of(inputs))


class GraphGather(nn.Module):
    """Per-molecule [sum | max] of atom features (torch_models/layers.py:6417-6479).

    Output always has ``batch_size`` rows; molecules absent from the batch give 0 in the sum half
    and -inf (tanh: -1) in the max half, as in the reference."""

    def __init__(self, batch_size: int, activation_fn: Optional[Callable] = None, **kwargs):
        super(GraphGather, self).__init__(**kwargs)
        self.batch_size = batch_size
        self.activation_fn = activation_fn

    def get_config(self) -> str:
        return f'{self.__class__.__name__}(batch_size:{self.batch_size},activation_fn:{self.activation_fn})'

    def forward(self, inputs: List[torch.Tensor]) -> torch.Tensor:
        assert self.batch_size > 1, "graph_gather requires batches larger than 1"
        membership = inputs[2]
        if membership.dim() != 1:
            raise AssertionError("segment_ids have be a 1-D tensor")
        if inputs[0].shape[0] != membership.shape[0]:
            raise AssertionError("segment_ids should be the same size as dimension 0 of input.")
        topo = topology_of(inputs, n_segments=self.batch_size)
        code = ops.act_code(self.activation_fn)
        z = ops.GraphGatherFn.apply(inputs[0], topo, self.batch_size, ACT_NONE if code is None else code)
        if code is None:
            z = self.activation_fn(z)
        return z
