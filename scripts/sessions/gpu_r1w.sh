#!/bin/bash
# GPU session r1w: molecule-group staged kernels — parity tests, A/B bench against the CSR kernels, group-size sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_staged.py -x -q > gpurun_out/r1w_staged.log 2>&1; echo "staged exit $?"
tail -15 gpurun_out/r1w_staged.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1w_pytest.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r1w_pytest.log
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value %.0f ms %.3f roof %.3f avg_us %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))'
for R in 64 32 48 96 128 192; do
  echo "== DCGC_GROUP_ROWS=$R"
  DCGC_GROUP_ROWS=$R timeout 300 python bench.py --no-cpu-baseline --no-e2e --breakdown gpurun_out/r1w_breakdown_R$R.md 2> gpurun_out/r1w_R$R.err | tee gpurun_out/r1w_bench_R$R.json | python -c "$show"
  grep -E "gather_sum|pool_" gpurun_out/r1w_breakdown_R$R.md
done
echo "== DCGC_NO_STAGED=1"
DCGC_NO_STAGED=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e --breakdown gpurun_out/r1w_breakdown_nostaged.md 2> gpurun_out/r1w_nostaged.err | tee gpurun_out/r1w_bench_nostaged.json | python -c "$show"
grep -E "gather_sum|pool_" gpurun_out/r1w_breakdown_nostaged.md
