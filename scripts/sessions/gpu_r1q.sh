#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1q_pytest.log 2>&1; echo "pytest exit $?"
tail -25 gpurun_out/r1q_pytest.log | cut -c1-220
for mode in tf32x3 bf16; do
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --gemm-mode $mode --breakdown gpurun_out/r1q_breakdown_$mode.md > gpurun_out/r1q_bench_$mode.json 2> gpurun_out/r1q_bench_$mode.err; echo "bench $mode exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r1q_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode value %.0f ms %.3f e2e %.0f roof %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
head -12 gpurun_out/r1q_breakdown_$mode.md
done
