#!/bin/bash
# Round 2, call 2: after the rounded lo residual, the cycle-free topology attachment and the rms/max anchored bounds.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r4b_pytest.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/r4b_pytest.log | tail -n 3; grep -E "^FAILED|AssertionError" gpurun_out/r4b_pytest.log | cut -c1-250 | head -n 40
timeout 600 python scripts/parity_probe.py > gpurun_out/r4b_parity_probe.json 2> gpurun_out/r4b_parity_probe.err; echo "probe exit $?"; cat gpurun_out/r4b_parity_probe.err | cut -c1-300 | tail -n 8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4b_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r4b_smoke.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r4b_bench_n1.json 2> gpurun_out/r4b_bench_n1.err; echo "bench exit $?"
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r4b_bench_n1.json") if l.startswith("{")][-1])
    print("value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
except Exception as e:
    print("no line", e)
P
