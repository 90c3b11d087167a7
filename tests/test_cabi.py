"""The C-ABI library loads on a CPU-only box and exports every symbol include/dcgc.h declares."""
import ctypes
import os
import re

import pytest

from deepchem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dcgc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dcgc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    handle = ctypes.CDLL(_lib.LIB_PATH) if os.path.exists(_lib.LIB_PATH) else _lib.lib()
    for name in _declared():
        assert hasattr(handle, name), name


def test_version_and_error_channel():
    L = _lib.lib()
    assert L.dcgc_version() >= 100
    info = _lib.LayoutInfo()
    st = L.dcgc_layout_plan(1, None, None, 1, 128, ctypes.byref(info))
    assert st == _lib.DCGC_ERR_INVALID and "null" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(st)


def test_device_probe_does_not_crash_without_gpu():
    assert _lib.lib().dcgc_device_ok() in (0, 1)


def test_struct_layout_matches_header():
    # 5 int64 + 2 int32 + 11 int64 + 12 int64
    assert ctypes.sizeof(_lib.LayoutInfo) == 5 * 8 + 8 + 11 * 8 + 12 * 8


def test_ops_refuse_cpu_tensors():
    import torch
    from deepchem_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops._check_dev(torch.zeros(2, 2), "x")
