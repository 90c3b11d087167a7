#!/usr/bin/env python
"""profiles/gather_sum_traffic.json from an `ncu --set full` capture of the staged gather-sum launches.

    ncu -i gpurun_out/rXX_mg_kernels.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_traffic.py /tmp/raw.csv "<the command that was profiled>" [kernel-name regex]

Writes the mean of dram__bytes_read.sum + dram__bytes_write.sum over the matching launches together with the commit
and the sha256 of the kernel source that was profiled; bench.py quotes it as roofline.traffic only while that source is
unchanged."""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path, command = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else r"mg_kernel<.*GatherSumOp")
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    launches = []
    for r in rows[hdr + 2:]:
        if len(r) < len(names) or not pat.search(r[col["Kernel Name"]]):
            continue
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[key]].replace(",", "")) * UNITS[units[col[key]]]
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", "")) if "gpu__time_duration.sum" in col else None
        launches.append({"kernel": r[col["Kernel Name"]][:80], "dram_bytes": tot, "duration": dur,
                         "duration_unit": units[col["gpu__time_duration.sum"]] if dur is not None else None})
    if not launches:
        raise SystemExit("no launch matches %s" % pat.pattern)
    src = os.path.join(ROOT, "deepchem_b200", "csrc", "molgroup_kernels.cu")
    out = {"traffic_bytes_per_launch": sum(l["dram_bytes"] for l in launches) / len(launches), "launches": launches,
           "metric": "dram__bytes_read.sum + dram__bytes_write.sum (ncu --set full --clock-control none)",
           "command": command,
           "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True,
                                    text=True).stdout.strip(),
           "kernel_source_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest()}
    json.dump(out, open(os.path.join(ROOT, "profiles", "gather_sum_traffic.json"), "w"), indent=1)
    print("%d launches, mean %.1f MB" % (len(launches), out["traffic_bytes_per_launch"] / 1e6))


if __name__ == "__main__":
    main()
