#!/bin/bash
# GPU session r1p: ncu launch list of the default bench command + full capture of the roofline kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r1p_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1p_launches.csv $CMD > gpurun_out/r1p_ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 300 $CMD > gpurun_out/r1p_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_sum -s 15 -c 5 -o gpurun_out/r1p_gather_sum $CMD > gpurun_out/r1p_ncu_full.log 2>&1
echo "ncu full exit $?"
