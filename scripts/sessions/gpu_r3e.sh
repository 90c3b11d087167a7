#!/bin/bash
# GPU session r3e: balanced GEMM row tiles (DCGC_TILE_BALANCE): full GPU suite + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3e_pytest.log 2>&1; echo "pytest exit $?"
tail -n 4 gpurun_out/r3e_pytest.log | cut -c1-300
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value %.0f ms %.4f" % (d["value"], d["ms_per_step"]))
        for r in d["kernels"]["rows"][:4]: print("   %-24s %7.1f us" % (r["scope"], r["us_per_step"]))'
for tb in 0 148 0 148; do
  echo "== DCGC_TILE_BALANCE=$tb"
  DCGC_TILE_BALANCE=$tb timeout 300 python bench.py --no-cpu-baseline --no-e2e 2> gpurun_out/r3e_tb$tb.err | tee gpurun_out/r3e_bench_tb$tb.json | python -c "$show"
done
