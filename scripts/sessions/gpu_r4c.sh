#!/bin/bash
mkdir -p gpurun_out
PROBE_MODES=tf32x3,fp32 timeout 600 python scripts/parity_probe.py small small_zinc tox21 > gpurun_out/r4c_parity_probe.json 2> gpurun_out/r4c_parity_probe.err; echo "probe exit $?"; cat gpurun_out/r4c_parity_probe.err | cut -c1-300 | tail -n 80
