#!/bin/bash
# GPU session r1v: ncu --set full of the staged kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r1v_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mg_" -s 33 -c 11 -o gpurun_out/r1v_mg $CMD > gpurun_out/r1v_ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/r1v_ncu_full.log
