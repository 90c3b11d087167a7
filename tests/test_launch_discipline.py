"""Static checks of the launch discipline of csrc/ (no GPU): every kernel is launched through dcgc_launch (programmatic
stream serialization, csrc/common.h) and therefore MUST wait for the kernel in front of it before it touches global
memory — a kernel without the wait would run ahead of its input."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "deepchem_b200", "csrc")


def _kernels(src):
    """-> [(name, text from the kernel's opening brace to the next __global__ / end of file)]"""
    starts = [m.start() for m in re.finditer(r"__global__", src)]
    out = []
    for i, s in enumerate(starts):
        end = starts[i + 1] if i + 1 < len(starts) else len(src)
        chunk = src[s:end]
        body = chunk[chunk.index("{"):] if "{" in chunk else ""
        names = re.findall(r"(\w+)\s*\(", chunk[:chunk.index("{")] if "{" in chunk else chunk)
        names = [n for n in names if n not in ("__launch_bounds__", "__cluster_dims__")]
        out.append((names[0] if names else "?", body))
    return out


def test_every_kernel_waits_for_its_predecessor():
    missing = []
    for path in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        for name, body in _kernels(open(path).read()):
            if "dcgc_griddep_wait" not in body:
                missing.append("%s: %s" % (os.path.basename(path), name))
    assert not missing, missing


def test_no_plain_triple_chevron_launches_remain():
    left = []
    for path in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        for i, line in enumerate(open(path).read().splitlines(), 1):
            code = line.split("//")[0]
            if "<<<" in code:
                left.append("%s:%d" % (os.path.basename(path), i))
    assert not left, left


def test_nothing_global_is_read_before_the_wait_in_the_set_up_sections():
    """The persistent kernels wait AFTER their shared-memory / tensor-memory set-up: between the kernel's first line
    and the wait there may be no global load (only the debug timeline writes, which nobody reads on the device)."""
    bad = []
    for path in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        for name, body in _kernels(open(path).read()):
            head = body[:body.index("dcgc_griddep_wait")] if "dcgc_griddep_wait" in body else ""
            for token in ("__ldg", "ld.global", "cp.async", "tile_of(", "ldg4("):
                if token in head:
                    bad.append("%s: %s reads global memory (%s) before the wait" % (os.path.basename(path), name, token))
    assert not bad, bad
