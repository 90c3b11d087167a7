"""ctypes binding of libdcgc.so (include/dcgc.h).  No torch types cross this boundary: only
raw pointers, sizes and a cudaStream_t.  There is no CPU fallback: if the library cannot be
loaded (or built), importing the ops raises."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcgc.so")

DCGC_OK = 0
DCGC_ERR_INVALID = -1
DCGC_ERR_DEGREE = -2
DCGC_ERR_INDEX = -3
DCGC_ERR_CUDA = -4
DCGC_ERR_NOMEM = -5

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
GEMM_FP32, GEMM_BF16, GEMM_TF32X3, GEMM_F16X3 = 0, 1, 2, 3
N_DEG = 11
TILE_ROWS = 128


class LayoutInfo(Structure):
    """dcgc_layout_info"""
    _fields_ = [
        ("n_mols", c_int64), ("n_segments", c_int64), ("n_atoms", c_int64), ("n_edges", c_int64),
        ("n_tiles", c_int64), ("tile_rows", c_int32), ("reserved", c_int32),
        ("deg_count", c_int64 * N_DEG),
        ("off_deg_slice", c_int64), ("off_membership", c_int64), ("off_perm", c_int64),
        ("off_row_ptr", c_int64), ("off_col_idx", c_int64), ("off_t_row_ptr", c_int64),
        ("off_t_src", c_int64), ("off_t_slot", c_int64), ("off_mol_ptr", c_int64),
        ("off_mol_atoms", c_int64), ("off_tiles", c_int64), ("slab_bytes", c_int64),
        ("off_groups", c_int64), ("group_rows", c_int32), ("n_groups_alloc", c_int32),
    ]


MODEL_MAX_LAYERS = 8


class Topology(Structure):
    """dcgc_topology"""
    _fields_ = [
        ("n_atoms", c_int64), ("n_edges", c_int64), ("n_segments", c_int64), ("n_tiles", c_int64),
        ("deg_count", c_int64 * N_DEG),
        ("row_ptr", c_void_p), ("col_idx", c_void_p), ("t_row_ptr", c_void_p), ("t_src", c_void_p),
        ("t_slot", c_void_p), ("mol_ptr", c_void_p), ("mol_atoms", c_void_p), ("membership", c_void_p),
        ("tiles", c_void_p), ("symmetric", c_int32), ("reserved", c_int32),
        ("groups", c_void_p), ("n_groups", c_int32), ("group_max_rows", c_int32),
        ("group_max_entries", c_int32), ("reserved2", c_int32), ("mg_records", c_void_p),
    ]


class GcModelConfig(Structure):
    """dcgc_gcmodel_config"""
    _fields_ = [
        ("n_layers", c_int32), ("n_feat", c_int32), ("widths", c_int32 * MODEL_MAX_LAYERS),
        ("dense", c_int32), ("n_out", c_int32), ("n_classes", c_int32), ("mode", c_int32),
        ("batch_norm", c_int32), ("gemm_mode", c_int32), ("bn_eps", c_float), ("bn_momentum", c_float),
        ("input_exact", c_int32), ("reserved", c_int32),
    ]


SYNC_MAX_RANKS = 8


class BnSync(Structure):
    """dcgc_bn_sync"""
    _fields_ = [("world", c_int32), ("rank", c_int32), ("cap", c_int32), ("reserved", c_int32),
                ("seq0", ctypes.c_uint64), ("mailbox", ctypes.c_void_p * SYNC_MAX_RANKS)]


class DmpnnInfo(ctypes.Structure):
    _fields_ = [(n, c_int64) for n in ("n_mols", "n_atoms", "n_bonds", "n_rows", "k", "n_a2b_entries",
                                       "n_map_entries")] + \
        [("keep_pads", c_int32), ("reserved", c_int32)] + \
        [(n, c_int64) for n in ("off_row_of_mol", "off_mol_ptr", "off_bond_src", "off_bond_edge", "off_a2b_ell",
                                "off_map_ell", "off_a2b_ptr", "off_a2b_idx", "off_a2b_t_ptr", "off_a2b_t_idx",
                                "off_map_ptr", "off_map_idx", "off_map_t_ptr", "off_map_t_idx", "slab_bytes")]


DMPNN_MAX_FFN = 8


class DmpnnModelConfig(Structure):
    """dcgc_dmpnn_model_config"""
    _fields_ = [("atom_fdim", c_int32), ("bond_fdim", c_int32), ("hidden", c_int32), ("depth", c_int32),
                ("ffn_layers", c_int32), ("ffn_hidden", c_int32), ("n_out", c_int32), ("aggregation", c_int32),
                ("aggregation_norm", c_float), ("gemm_mode", c_int32)]


class DmpnnTables(Structure):
    """dcgc_dmpnn_tables"""
    _fields_ = [("n_mols", c_int64), ("n_atoms", c_int64), ("n_rows", c_int64)] + \
        [(n, c_void_p) for n in ("mol_ptr", "a2b_ptr", "a2b_idx", "a2b_t_ptr", "a2b_t_idx", "map_ptr", "map_idx",
                                 "map_t_ptr", "map_t_idx")]


_P = c_void_p
_SIGNATURES = {
    "dcgc_dmpnn_model_layout": (c_int32, [POINTER(DmpnnModelConfig), _P, POINTER(c_int64)]),
    "dcgc_dmpnn_model_workspace_bytes": (c_int64, [POINTER(DmpnnModelConfig), c_int64, c_int64, c_int64]),
    "dcgc_dmpnn_model_forward": (c_int32, [POINTER(DmpnnModelConfig), POINTER(DmpnnTables), _P, c_int64, _P, c_int64,
                                           _P, _P, c_int64, _P, _P, _P]),
    "dcgc_dmpnn_model_train_step": (c_int32, [POINTER(DmpnnModelConfig), POINTER(DmpnnTables), _P, c_int64, _P,
                                              c_int64, _P, _P, _P, _P, _P, c_int64, _P, _P, _P]),
    "dcgc_dmpnn_plan": (c_int32, [c_int64, _P, _P, _P, _P, c_int32, POINTER(DmpnnInfo)]),
    "dcgc_dmpnn_build": (c_int32, [c_int64, _P, _P, _P, _P, POINTER(DmpnnInfo), _P]),
    "dcgc_dmpnn_concat_rows": (c_int32, [_P, c_int64, c_int32, _P, c_int64, c_int32, _P, _P, c_int64, _P, c_int64, _P]),
    "dcgc_segment_readout_fwd": (c_int32, [_P, c_int64, _P, c_int64, c_int32, c_int32, c_float, _P, c_int64, _P]),
    "dcgc_segment_readout_bwd": (c_int32, [_P, c_int64, _P, c_int64, c_int64, c_int32, c_int32, c_float, _P, c_int64,
                                           _P]),
    "dcgc_last_error": (c_char_p, []),
    "dcgc_version": (c_int32, []),
    "dcgc_device_ok": (c_int32, []),
    "dcgc_launch_count": (c_int64, []),
    "dcgc_profile_begin": (c_int32, [c_char_p]),
    "dcgc_profile_end": (c_int32, [POINTER(ctypes.c_double), POINTER(c_int64)]),
    "dcgc_profile_report": (c_int32, [c_char_p, c_int64]),
    "dcgc_layout_plan": (c_int32, [c_int64, _P, _P, c_int64, c_int32, POINTER(LayoutInfo)]),
    "dcgc_layout_build": (c_int32, [c_int64, _P, _P, _P, POINTER(LayoutInfo), _P]),
    "dcgc_layout_permute_features_host": (c_int32, [_P, c_int64, _P, c_int64, c_int32, _P, c_int64, c_int32]),
    "dcgc_layout_plan_from_deg": (c_int32, [_P, c_int64, c_int32, POINTER(LayoutInfo)]),
    "dcgc_layout_build_from_deg": (c_int32, [_P, _P, _P, POINTER(LayoutInfo), _P]),
    "dcgc_packed_take_plan": (c_int32, [c_int64, _P, c_int64, _P, _P, POINTER(c_int64), POINTER(c_int64)]),
    "dcgc_packed_take": (c_int32, [c_int64, _P, c_int64, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, _P, _P,
                                   c_int32]),
    "dcgc_h2d_chunked": (c_int32, [_P, _P, c_int64, c_int64, _P]),
    "dcgc_permute_rows": (c_int32, [_P, c_int64, _P, c_int64, c_int32, _P, c_int64, _P]),
    "dcgc_permute_rows_i8": (c_int32, [_P, c_int64, _P, c_int64, c_int32, _P, c_int64, _P]),
    "dcgc_gather_sum": (c_int32, [_P, c_int64, _P, _P, c_int64, c_int32, _P, c_int64, _P, c_int64, _P]),
    "dcgc_gather_sum_bucketed": (c_int32, [_P, c_int64, _P, _P, c_int64, c_int32, _P, c_int64, _P, c_int64, _P]),
    "dcgc_mg_supported": (c_int32, [_P, c_int64, c_int64]),
    "dcgc_mg_record_bytes": (c_int64, [c_int64]),
    "dcgc_mg_prepare": (c_int32, [_P, _P, _P]),
    "dcgc_mg_gather_sum": (c_int32, [_P, c_int64, _P, c_int32, c_int32, _P, c_int64, _P, c_int64, _P]),
    "dcgc_mg_pool_fwd": (c_int32, [_P, c_int64, _P, _P, _P, c_int32, _P, c_int64, _P, c_int64, _P]),
    "dcgc_mg_pool_bwd": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, c_int32, _P, c_int64, _P]),
    "dcgc_mg_pool_bwd_stats": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int32, _P, c_int64, _P, c_int64, _P, _P,
                                         POINTER(c_int32), _P]),
    "dcgc_pool_fwd": (c_int32, [_P, c_int64, _P, _P, _P, _P, c_int64, c_int32, _P, c_int64, _P, c_int64, _P]),
    "dcgc_pool_bwd": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, _P, _P, c_int64, c_int32, _P, c_int64, _P]),
    "dcgc_gather_fwd": (c_int32, [_P, c_int64, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P, c_int64, _P, _P]),
    "dcgc_gather_bwd": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, c_int64, c_int32, c_int32, _P, c_int64, _P]),
    "dcgc_group_gemm_fwd": (c_int32, [c_int32, _P, c_int64, c_int32, _P, c_int64, c_int32, _P, _P, c_int32,
                                      _P, c_int64, c_int32, c_int64, c_int32, _P, c_int64, _P]),
    "dcgc_group_gemm_fwd_stats": (c_int32, [c_int32, _P, c_int64, c_int32, _P, c_int64, c_int32, _P, _P, c_int32,
                                            _P, c_int64, c_int32, c_int64, c_int32, _P, c_int64, _P,
                                            POINTER(c_int32), _P]),
    "dcgc_gemm_stats_max_chunks": (c_int32, []),
    "dcgc_linear_fwd_stats": (c_int32, [c_int32, _P, c_int64, c_int32, _P, _P, c_int32, c_int64, c_int32, _P, c_int64,
                                        _P, POINTER(c_int32), _P]),
    "dcgc_group_gemm_dgrad": (c_int32, [c_int32, _P, c_int64, c_int32, _P, c_int32, c_int32, _P, c_int64,
                                        c_int32, c_int64, _P, c_int64, _P, c_int64, _P]),
    "dcgc_group_gemm_wgrad_workspace": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "dcgc_group_gemm_wgrad": (c_int32, [c_int32, _P, c_int64, c_int32, _P, c_int64, c_int32, _P, c_int64,
                                        c_int32, _P, c_int32, _P, _P, _P, c_int64, _P]),
    "dcgc_linear_fwd": (c_int32, [c_int32, _P, c_int64, c_int32, _P, _P, c_int32, c_int64, c_int32, _P, c_int64, _P]),
    "dcgc_linear_dgrad": (c_int32, [c_int32, _P, c_int64, c_int32, _P, c_int32, c_int64, _P, c_int64, _P]),
    "dcgc_linear_wgrad_workspace": (c_int64, [c_int32, c_int32]),
    "dcgc_linear_wgrad": (c_int32, [c_int32, _P, c_int64, c_int32, _P, c_int64, c_int32, c_int64, _P, _P, _P,
                                    c_int64, _P]),
    "dcgc_pair_contract_fwd": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int32, c_int32, _P, c_int64, _P]),
    "dcgc_gru_gates_fwd": (c_int32, [_P, c_int64, _P, _P, _P, c_int64, c_int64, c_int32, _P, c_int64, _P, c_int64, _P]),
    "dcgc_gru_out_fwd": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, c_int64, _P, c_int64, c_int64, c_int32, _P,
                                   c_int64, _P]),
    "dcgc_setgather_attend_fwd": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, c_int64, c_int32, c_int32, _P, c_int64,
                                            _P]),
    "dcgc_lstm_step_fwd": (c_int32, [_P, c_int64, _P, c_int64, c_int32, _P, _P, _P]),
    "dcgc_pair_contract_bwd_x": (c_int32, [_P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int32, c_int32, _P, c_int64, _P]),
    "dcgc_gru_out_bwd": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, _P, _P, c_int64, _P, c_int64, c_int64, c_int32,
                                   _P, c_int64, _P, c_int64, _P, c_int64, _P]),
    "dcgc_gru_gates_bwd": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int64,
                                     c_int32, _P, c_int64, _P, c_int64, _P]),
    "dcgc_setgather_attend_bwd": (c_int32, [_P, c_int64, _P, c_int64, _P, c_int64, _P, _P, c_int64, c_int32, c_int32,
                                            _P, c_int64, _P, c_int64, _P]),
    "dcgc_lstm_step_bwd": (c_int32, [_P, c_int64, _P, _P, _P, c_int64, c_int32, _P, c_int64, _P, _P]),
    "dcgc_gcmodel_layout": (c_int32, [POINTER(GcModelConfig), _P, _P, POINTER(c_int64), POINTER(c_int64)]),
    "dcgc_gcmodel_workspace_bytes": (c_int64, [POINTER(GcModelConfig), c_int64, c_int64]),
    "dcgc_gcmodel_forward": (c_int32, [POINTER(GcModelConfig), POINTER(Topology), _P, c_int64, c_int64, _P, _P,
                                       c_int32, _P, c_int64, _P, _P, _P, _P]),
    "dcgc_gcmodel_train_step": (c_int32, [POINTER(GcModelConfig), POINTER(Topology), _P, c_int64, _P, _P, c_int64,
                                          _P, _P, _P, _P, c_int64, _P, _P, _P]),
    "dcgc_gcmodel_train_step_ev": (c_int32, [POINTER(GcModelConfig), POINTER(Topology), _P, c_int64, _P, _P, c_int64,
                                          _P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int32, _P]),
    "dcgc_gcmodel_train_step_sync": (c_int32, [POINTER(GcModelConfig), POINTER(Topology), _P, c_int64, _P, _P, c_int64,
                                               _P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int32, POINTER(BnSync), _P]),
    "dcgc_bn_sync_mailbox_bytes": (c_int64, [c_int32, c_int32]),
    "dcgc_p2p_alloc": (c_int32, [c_int64, POINTER(ctypes.c_void_p), _P]),
    "dcgc_p2p_open": (c_int32, [_P, POINTER(ctypes.c_void_p)]),
    "dcgc_p2p_close": (c_int32, [_P]),
    "dcgc_p2p_free": (c_int32, [_P]),
    "dcgc_tc_f16_overflow": (c_int32, []),
    "dcgc_adam_step": (c_int32, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_int64, c_float, _P]),
}

_lib = None


class DcgcError(RuntimeError):
    pass


def exported_symbols():
    """Names include/dcgc.h declares (kept in sync by tests/test_cabi.py)."""
    return sorted(_SIGNATURES)


def lib():
    """Load (building first if the .so is missing and nvcc is present) and return the library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # no nvcc, or compile error
            raise DcgcError(
                "libdcgc.so is not built (%s). Run `python -m deepchem_b200.build`; there is no "
                "CPU or pure-PyTorch fallback for this path." % (e,))
    try:
        handle = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise DcgcError("cannot load %s: %s (no fallback path exists)" % (LIB_PATH, e))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def last_error():
    return lib().dcgc_last_error().decode("utf-8", "replace")


def check(status):
    """Translate a C status into the exception the reference raises in the same situation."""
    if status == DCGC_OK:
        return
    msg = last_error()
    if status == DCGC_ERR_DEGREE:
        # featurizer raises ValueError for degree > max_deg (feat/graph_features.py:34-36)
        raise ValueError(msg)
    if status in (DCGC_ERR_INVALID, DCGC_ERR_INDEX):
        raise ValueError(msg)
    if status == DCGC_ERR_NOMEM:
        raise MemoryError(msg)
    raise DcgcError("libdcgc error %d: %s" % (status, msg))
