#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mg_kernel" -s 5 -c 1 -o gpurun_out/r2b_mg_gather python scripts/mg_timeline.py gather > gpurun_out/r2b_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/r2b_ncu.log
