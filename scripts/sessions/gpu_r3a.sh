#!/bin/bash
# GPU session r3a: full GPU suite + D-MPNN end-to-end with the prefetching fit pipeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest exit $?"
tail -n 4 gpurun_out/r3a_pytest.log
timeout 600 python scripts/bench_extra.py dmpnn > gpurun_out/r3a_dmpnn.json 2> gpurun_out/r3a_dmpnn.err; echo "dmpnn exit $?"
cat gpurun_out/r3a_dmpnn.json; tail -n 5 gpurun_out/r3a_dmpnn.err
