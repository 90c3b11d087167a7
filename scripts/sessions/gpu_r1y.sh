#!/bin/bash
mkdir -p gpurun_out
for w in gather gather_add pool_fwd pool_bwd; do timeout 120 python scripts/mg_timeline.py $w > gpurun_out/r1y_timeline_$w.txt 2>&1; head -1 gpurun_out/r1y_timeline_$w.txt; done
cat gpurun_out/r1y_timeline_gather.txt
