#!/bin/bash
# GPU session r3j: D-MPNN gathers with the element-wise steps folded in; early weight images on by default: full suite + rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_pytest.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r3j_pytest.log | cut -c1-200
timeout 300 python scripts/bench_extra.py dmpnn > gpurun_out/r3j_dmpnn.json 2> gpurun_out/r3j_dmpnn.err; echo "dmpnn exit $?"; cut -c1-330 gpurun_out/r3j_dmpnn.json
timeout 300 python scripts/bench_extra.py predict > gpurun_out/r3j_predict.json 2> gpurun_out/r3j_predict.err; echo "predict exit $?"; cut -c1-330 gpurun_out/r3j_predict.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
