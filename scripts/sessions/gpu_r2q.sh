#!/bin/bash
# GPU session r2q: fused BatchNorm-backward sums in the staged pool backward: tests + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r2q_pytest.log
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("n=%d value %.0f ms %.3f e2e %s roof %.3f avg_us %.2f" % (d["n_gpus"], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step"), d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))'
for nf in 1 0; do
  echo "== DCGC_NO_FUSED_BN_BWD=$nf"
  DCGC_NO_FUSED_BN_BWD=$nf timeout 300 python bench.py --no-cpu-baseline --no-e2e --breakdown gpurun_out/r2q_breakdown_nofuse$nf.md 2> gpurun_out/r2q_nf$nf.err | tee gpurun_out/r2q_bench_nofuse$nf.json | python -c "$show"
  grep -E "bn_stats_bwd|pool_bwd|total" gpurun_out/r2q_breakdown_nofuse$nf.md
done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; echo "bench exit $?"
cat gpurun_out/r2q_bench_n1.json | python -c "$show"
