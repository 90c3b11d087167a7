#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_staged.py -x -q > gpurun_out/r1z_staged.log 2>&1; echo "staged exit $?"
tail -3 gpurun_out/r1z_staged.log
for w in gather gather_add pool_fwd pool_bwd; do timeout 120 python scripts/mg_timeline.py $w > gpurun_out/r1z_timeline_$w.txt 2>&1; head -1 gpurun_out/r1z_timeline_$w.txt; done
cat gpurun_out/r1z_timeline_gather.txt | head -20
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value %.0f ms %.3f roof %.3f avg_us %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))'
for R in 64 96 128; do
  echo "== DCGC_GROUP_ROWS=$R"
  DCGC_GROUP_ROWS=$R timeout 300 python bench.py --no-cpu-baseline --no-e2e --breakdown gpurun_out/r1z_breakdown_R$R.md 2> gpurun_out/r1z_R$R.err | tee gpurun_out/r1z_bench_R$R.json | python -c "$show"
  grep -E "gather_sum|pool_" gpurun_out/r1z_breakdown_R$R.md
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1z_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r1z_pytest.log
