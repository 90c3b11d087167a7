"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for
the last full training step (steps are delimited by the single GraphGather forward launch).

    python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md
"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    items = []
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3}.get(r[ui], 1e-3)
        items.append((r[ki], v * scale))
    return items


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else 'gather_fwd_kernel'
    items = load(path)
    idx = [i for i, (k, _) in enumerate(items) if marker in k]
    s0, s1 = idx[-2], idx[-1]
    step = items[s0:s1]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in step:
        k2 = re.sub(r'\(.*', '', k)
        k2 = re.sub(r'^void ', '', k2)[:100]
        agg[k2][0] += 1
        agg[k2][1] += v
    tot = sum(v for _, v in step)
    ours = sum(v for k, v in step if '<unnamed>::' in k and 'native' not in k)
    print("# launch list summary: %s" % path)
    print()
    print("One training step (between two `%s` launches): %d launches, %.1f us of kernel time "
          "(ncu per-launch durations are cold-cache and serialised: compare shares, not absolutes)."
          % (marker, len(step), tot))
    print("libdcgc kernels: %.1f us (%.1f%% of the step)." % (ours, 100 * ours / tot))
    print()
    print("| us | share | launches | kernel |")
    print("|---:|---:|---:|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if v / tot < 0.002:
            continue
        print("| %.1f | %.1f%% | %d | `%s` |" % (v, 100 * v / tot, n, k))


if __name__ == "__main__":
    main()
