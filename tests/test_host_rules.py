"""Small host-side rules of the models (no GPU)."""
import os

import torch


def test_batchnorm_batch_counters_are_flushed_when_somebody_can_see_them():
    """The fused engine counts BatchNorm1d.num_batches_tracked on the host (graphconvmodel.py::_engine_step) and adds
    the pending count in a state_dict pre-hook, at the end of fit_generator and before a restore replaces the buffers."""
    from deepchem_b200.graphconvmodel import GraphConvModel
    m = GraphConvModel(1, [64, 64], 128, mode="regression", batch_size=8, device=torch.device("cpu"))
    assert m._bn_steps_pending == 0
    m._bn_steps_pending = 3
    sd = m.model.state_dict()
    counters = [int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")]
    assert counters == [3, 3, 3] and m._bn_steps_pending == 0
    m._bn_steps_pending = 2
    m._flush_bn_counters()
    assert [int(bn.num_batches_tracked) for bn in m.model.batch_norms] == [5, 5, 5]
    m._flush_bn_counters()                                    # nothing pending: no change
    assert [int(bn.num_batches_tracked) for bn in m.model.batch_norms] == [5, 5, 5]


def test_layout_worker_count_follows_the_cores_of_the_rank(monkeypatch):
    """cores of this process / ranks of the node - 2 (launching + prefetch threads), between 1 and 4."""
    from deepchem_b200 import graphconvmodel as G
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(32)), raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert G._default_host_workers() == 2
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert G._default_host_workers() == 4
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: {0, 1}, raising=False)
    assert G._default_host_workers() == 1


def test_forward_products_of_the_tf32x3_mode_use_fp16_halves_unless_switched_off(monkeypatch):
    """ops.forward_gemm_mode: the rule of csrc/dmpnn_model.cu::forward_mode and csrc/model.cu::fwd_f16x3_on."""
    from deepchem_b200 import _lib, ops
    monkeypatch.delenv("DCGC_FWD_F16X3", raising=False)
    assert ops.forward_gemm_mode(_lib.GEMM_TF32X3) == _lib.GEMM_F16X3
    assert ops.forward_gemm_mode(_lib.GEMM_FP32) == _lib.GEMM_FP32
    assert ops.forward_gemm_mode(_lib.GEMM_BF16) == _lib.GEMM_BF16
    monkeypatch.setenv("DCGC_FWD_F16X3", "0")
    assert ops.forward_gemm_mode(_lib.GEMM_TF32X3) == _lib.GEMM_TF32X3
