#!/bin/bash
# GPU session r3f: configs 1 and 2 through the public API, tox21-shaped engine parity, D-MPNN e2e stability
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q > gpurun_out/r3f_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r3f_pytest.log | cut -c1-200
timeout 900 python scripts/bench_configs.py > gpurun_out/r3f_configs.json 2> gpurun_out/r3f_configs.err; echo "configs exit $?"; cat gpurun_out/r3f_configs.json; tail -n 3 gpurun_out/r3f_configs.err
timeout 300 python scripts/dmpnn_e2e_stability.py > gpurun_out/r3f_dmpnn_stability.json 2> gpurun_out/r3f_dmpnn_stability.err; echo "stability exit $?"; cat gpurun_out/r3f_dmpnn_stability.json; tail -n 3 gpurun_out/r3f_dmpnn_stability.err
