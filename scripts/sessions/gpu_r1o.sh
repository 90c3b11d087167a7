#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1o_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r1o_pytest.log
timeout 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --breakdown gpurun_out/r1o_breakdown.md > gpurun_out/r1o_bench.json 2> gpurun_out/r1o_bench.err; echo "bench exit $?"
cut -c1-2600 gpurun_out/r1o_bench.json; tail -3 gpurun_out/r1o_bench.err; cat gpurun_out/r1o_breakdown.md
