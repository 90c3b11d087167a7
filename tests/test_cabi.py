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


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of every struct in dcgc.h, as gcc lays them out, against the ctypes mirrors."""
    import subprocess
    structs = {"dcgc_layout_info": _lib.LayoutInfo, "dcgc_topology": _lib.Topology,
               "dcgc_gcmodel_config": _lib.GcModelConfig, "dcgc_dmpnn_info": _lib.DmpnnInfo}
    lines = []
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dcgc.h"\nint main(void) {\n%s\nreturn 0; }\n'
                   % "\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)
    assert ctypes.sizeof(_lib.LayoutInfo) == 5 * 8 + 8 + 11 * 8 + 12 * 8 + 16


def test_ops_refuse_cpu_tensors():
    import torch
    from deepchem_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops._check_dev(torch.zeros(2, 2), "x")
