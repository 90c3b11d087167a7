#!/bin/bash
# GPU session r3i: early weight images on a side stream: full GPU suite + A/B bench; configs re-run; (2-GPU DP check separately)
mkdir -p gpurun_out
DCGC_EARLY_IMAGES=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3i_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r3i_pytest.log | cut -c1-200
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value %.0f ms %.4f launches %d" % (d["value"], d["ms_per_step"], d["gpu_launches"]))
        for r in d["kernels"]["rows"][:4]: print("   %-24s %7.1f us" % (r["scope"], r["us_per_step"]))'
for ei in 0 1 0 1; do
  echo "== DCGC_EARLY_IMAGES=$ei"
  DCGC_EARLY_IMAGES=$ei timeout 300 python bench.py --no-cpu-baseline --no-e2e 2> gpurun_out/r3i_ei$ei.err | tee gpurun_out/r3i_bench_ei$ei.json | python -c "$show"
done
DCGC_EARLY_IMAGES=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3i_bench_n1.json 2> gpurun_out/r3i_bench_n1.err; echo "bench exit $?"; python -c "$show" < gpurun_out/r3i_bench_n1.json
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r3i_bench_n1.json') if l.startswith('{')][-1]); print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'])"
timeout 600 python scripts/bench_configs.py cfg2 > gpurun_out/r3i_configs.json 2> gpurun_out/r3i_configs.err; echo "configs exit $?"; cut -c1-260 gpurun_out/r3i_configs.json
