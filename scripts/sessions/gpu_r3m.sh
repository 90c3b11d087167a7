#!/bin/bash
# GPU session r3m (1 GPU): full GPU test suite (new: sharding exactness, real Tox21), config 2 on the real Tox21 file, default bench
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r3m_pytest.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/r3m_pytest.log | cut -c1-300
timeout 120 python scripts/bench_configs.py cfg2_real > gpurun_out/r3m_cfg2_real.json 2> gpurun_out/r3m_cfg2_real.err; echo "cfg2_real exit $?"; cut -c1-400 gpurun_out/r3m_cfg2_real.json; tail -n 3 gpurun_out/r3m_cfg2_real.err | cut -c1-300
timeout 200 python bench.py > gpurun_out/r3m_bench_n1.json 2> gpurun_out/r3m_bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r3m_bench_n1.json") if l.startswith("{")][-1])
    print("n=%d value %.0f ms %.4f e2e %.0f (%.4f ms) roof %.3f cpu %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["cpu_baseline"]))
except Exception as e:
    print("no bench line", e)
P
