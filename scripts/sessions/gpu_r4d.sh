#!/bin/bash
# Round 2, call 5: suite with the rms/max anchored bounds + flip allowance, smoke on the Tox21-shaped case, the rewritten
# bench.py (frac / frac_strict, NVML clocks, 200-step e2e, CPU lines, dmpnn + predict sub-records), ncu capture for traffic.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4d_pytest.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/r4d_pytest.log | tail -n 3; grep -E "^FAILED|AssertionError|^E  " gpurun_out/r4d_pytest.log | cut -c1-250 | head -n 40
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4d_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r4d_smoke.log | cut -c1-300
timeout 600 python bench.py --breakdown gpurun_out/r4d_breakdown.md > gpurun_out/r4d_bench_n1.json 2> gpurun_out/r4d_bench_n1.err; echo "bench exit $?"; tail -n 5 gpurun_out/r4d_bench_n1.err | cut -c1-300
python - <<'P'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r4d_bench_n1.json") if l.startswith("{")][-1])
    print("value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"], d["e2e"]["prepared_at_t0"], "roofline", d["roofline"]["frac"], d["roofline"]["frac_strict"])
    print("clocks", d["clocks"], "e2e clocks", d["e2e"]["clocks"])
    print("dmpnn", d["dmpnn"]["value"], d["dmpnn"]["e2e"]["value"], d["dmpnn"].get("roofline", {}).get("frac"), d["dmpnn"].get("cpu_baseline"))
    print("predict", d["predict"]["value"], d["predict"]["seconds"], d["predict"].get("roofline", {}).get("frac"), d["predict"].get("cpu_baseline"))
    print("cpu", json.dumps(d["cpu_baseline"])[:600])
except Exception as e:
    print("no line", repr(e))
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mg_kernel -c 10 -o gpurun_out/r4d_mg_kernels python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sub "" > gpurun_out/r4d_ncu_full.log 2>&1; echo "ncu full exit $?"
