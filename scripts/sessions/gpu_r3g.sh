#!/bin/bash
# GPU session r3g: shuffled-epoch host path (C gather in the layout workers): parity + rate; configs 1/2 again
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3g_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r3g_pytest.log | cut -c1-200
timeout 600 python scripts/shuffled_e2e.py > gpurun_out/r3g_shuffled.json 2> gpurun_out/r3g_shuffled.err; echo "shuffled exit $?"; cat gpurun_out/r3g_shuffled.json; tail -n 3 gpurun_out/r3g_shuffled.err
timeout 900 python scripts/bench_configs.py > gpurun_out/r3g_configs.json 2> gpurun_out/r3g_configs.err; echo "configs exit $?"; cat gpurun_out/r3g_configs.json; tail -n 3 gpurun_out/r3g_configs.err
