#!/bin/bash
# GPU session r1c: parity tests, bench in both GEMM modes, ncu launch lists, one full capture of gather_sum
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1c_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r1c_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r1c_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r1c_bench_fp32.json 2> gpurun_out/r1c_bench_fp32.err; echo "bench fp32 exit $?"
timeout 400 python bench.py --steps 20 --warmup 5 --gemm-mode tf32x3 --no-cpu-baseline > gpurun_out/r1c_bench_tc.json 2> gpurun_out/r1c_bench_tc.err; echo "bench tc exit $?"
cat gpurun_out/r1c_bench_fp32.json gpurun_out/r1c_bench_tc.json
tail -3 gpurun_out/r1c_bench_tc.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r1c_plain_fp32.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches_fp32.csv $CMD > gpurun_out/r1c_ncu_fp32.log 2>&1
echo "ncu list fp32 exit $?"
timeout 300 $CMD --gemm-mode tf32x3 > gpurun_out/r1c_plain_tc.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches_tc.csv $CMD --gemm-mode tf32x3 > gpurun_out/r1c_ncu_tc.log 2>&1
echo "ncu list tc exit $?"
timeout 300 $CMD > gpurun_out/r1c_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_sum -s 20 -c 5 -o gpurun_out/r1c_gather_sum $CMD > gpurun_out/r1c_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
