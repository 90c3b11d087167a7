#!/usr/bin/env python
"""Which Blackwell-specific SASS instructions each kernel of libdcgc.so contains (tcgen05 MMA / TMEM / TMA / bulk copies /
mbarrier transactions), counted from `cuobjdump -sass`.  python scripts/sass_summary.py > profiles/<name>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deepchem_b200", "libdcgc.so")
WANT = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "LDGSTS", "REDUX",
        "ATOM", "RED.", "ACQBULK", "PREEXIT")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for w in WANT:
        if op.startswith(w):
            counts[fn][op if w in ("UTMALDG", "UBLKCP", "LDTM", "STTM", "SYNCS") else w.rstrip(".")] += 1
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print("# SASS of `deepchem_b200/libdcgc.so` (sm_100a): tensor-core / tensor-memory / TMA instructions per kernel\n")
print("`python scripts/sass_summary.py` (cuobjdump -sass, counted per kernel; kernels without any of them are omitted). "
      "UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (tensor-map TMA load), "
      "UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations; ATOM / RED = atomics (none on float data); ACQBULK / PREEXIT = griddepcontrol.wait / "
      "launch_dependents (programmatic dependent launch: every kernel has both).\n")
print("| kernel | instructions |")
print("|---|---|")
for n, d in zip(names, dem):
    c = counts[n]
    if not c:
        continue
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*$", "", d)
    print("| `%s` | %s |" % (d.replace("void ", ""), ", ".join("%s x%d" % kv for kv in sorted(c.items()))))
