#!/usr/bin/env python
"""Per-tensor distances of the fused engine from the float64 oracle (no assertions): the numbers behind
tests/test_gpu_engine_fp64.py.  python scripts/parity_probe.py [case ...] > gpurun_out/parity_probe.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import fp64_anchored_errors, rel_err  # noqa: E402
import test_gpu_engine_fp64 as T  # noqa: E402


def main():
    cases = sys.argv[1:] or ["bench", "tox21"]
    rep = {}
    for case in cases:
        for mode in (os.environ.get("PROBE_MODES", "tf32x3,fp32,bf16").split(",")):
            r = T.run_case(case, mode)
            o32, l32, g32 = r["res"][torch.float32]
            o64, l64, g64 = r["res"][torch.float64]
            k = r["out_idx"]
            ref = o64[k].numpy()
            row = {"out": rel_err(r["out"].numpy().reshape(ref.shape), ref), "out_fp32_oracle": rel_err(o32[k].numpy(), ref),
                   "loss": abs(r["loss"] - l64) / abs(l64), "loss_fp32_oracle": abs(l32 - l64) / abs(l64), "grads": {}}
            for name, g in r["grads"].items():
                row["grads"][name] = list(fp64_anchored_errors(g, g32[name], g64[name]))
            worst = max(row["grads"].items(), key=lambda kv: kv[1][2] / max(kv[1][3], 1e-5))
            row["worst_ratio"] = [worst[0]] + ["%.2e" % v for v in worst[1]]
            import numpy as np
            g = row["grads"]
            for pref in ("graph_convs.0.W", "graph_convs.1.W", "graph_convs.2.W", "graph_convs.0.b", "graph_convs.1.b", "batch_norms", "dense", "regression_dense", "reshape_dense"):
                sel = [v for k2, v in g.items() if k2.startswith(pref) and v[3] > 0]
                if sel:
                    sys.stderr.write("   %-18s n=%2d  rms ours median %.2e max %.2e | oracle median %.2e max %.2e\n" % (
                        pref, len(sel), np.median([v[2] for v in sel]), max(v[2] for v in sel),
                        np.median([v[3] for v in sel]), max(v[3] for v in sel)))
            rep["%s/%s" % (case, mode)] = row
            sys.stderr.write("%s/%s out %.2e/%.2e loss %.2e/%.2e worst %s\n" % (
                case, mode, row["out"], row["out_fp32_oracle"], row["loss"], row["loss_fp32_oracle"], row["worst_ratio"]))
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
