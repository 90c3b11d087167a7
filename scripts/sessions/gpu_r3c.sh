#!/bin/bash
# GPU session r3c: Delaney GPU tests, bench with the per-kernel table, GEMM / wgrad per-role timelines
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_delaney.py tests/test_gpu_mpnn.py -m gpu -q > gpurun_out/r3c_pytest.log 2>&1; echo "pytest exit $?"
tail -n 4 gpurun_out/r3c_pytest.log
timeout 600 python bench.py --no-cpu-baseline --breakdown gpurun_out/r3c_breakdown.md > gpurun_out/r3c_bench_n1.json 2> gpurun_out/r3c_bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r3c_bench_n1.json") if l.startswith("{")][-1])
print("value %.0f ms %.4f e2e %s roof %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]))
for r in d["kernels"]["rows"]:
    print("%-26s %8.1f us %5.1f calls %s" % (r["scope"], r["us_per_step"], r["calls_per_step"],
          ("%.0f MB %.0f GB/s %.2f" % (r["algorithmic_mb_per_step"], r["achieved_gbs"], r["frac_of_hbm_peak"])) if r.get("achieved_gbs") else ""))
P
timeout 200 python scripts/gemm_timeline.py > gpurun_out/r3c_gemm_timeline.log 2>&1; echo "timeline exit $?"; head -n 30 gpurun_out/r3c_gemm_timeline.log
timeout 200 python scripts/wgrad_timeline.py > gpurun_out/r3c_wgrad_timeline.log 2>&1; echo "wgrad timeline exit $?"; head -n 30 gpurun_out/r3c_wgrad_timeline.log
