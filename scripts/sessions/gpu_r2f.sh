#!/bin/bash
# GPU session r2f (1 GPU): evidence for the molecule-group staged kernels — all GPU tests, default bench + in-situ
# breakdown, reference arm, ncu launch list, ncu full capture of the staged kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2f_pytest.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference_arm.json 2>/dev/null; echo "ref exit $?"
timeout 600 python bench.py --breakdown gpurun_out/r2f_breakdown.md > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench n1 exit $?"
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("n=%d value %.0f ms %.3f e2e %.0f e2e_ms %.3f roof %.3f avg_us %.2f cpu %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"], (d.get("cpu_baseline") or {}).get("value")))'
cat gpurun_out/r2f_bench_n1.json | python -c "$show"
cat gpurun_out/r2f_breakdown.md
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r2f_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mg_kernel" -s 33 -c 11 -o gpurun_out/r2f_mg_kernels $CMD > gpurun_out/r2f_ncu_full.log 2>&1
echo "ncu full exit $?"
