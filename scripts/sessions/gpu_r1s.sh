#!/bin/bash
# GPU session r1s (2 GPUs): final round-1 evidence — tests, default bench (N=1, N=2), reference arm, ncu launch list,
# ncu full capture of the roofline kernel, in-situ breakdown
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1s_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r1s_pytest.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1s_bench_reference_arm.json 2>/dev/null; echo "ref exit $?"
timeout 600 python bench.py --breakdown gpurun_out/r1s_breakdown.md > gpurun_out/r1s_bench_n1.json 2> gpurun_out/r1s_bench_n1.err; echo "bench n1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r1s_bench_n2.json 2> gpurun_out/r1s_bench_n2.err; echo "bench n2 exit $?"
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("n=%d value %.0f ms %.3f e2e %.0f e2e_ms %.3f roof %.3f cpu %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value")))'
cat gpurun_out/r1s_bench_n1.json gpurun_out/r1s_bench_n2.json | python -c "$show"
cat gpurun_out/r1s_breakdown.md
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > gpurun_out/r1s_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1s_launches.csv $CMD > gpurun_out/r1s_ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 300 $CMD > gpurun_out/r1s_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_sum|tc_gemm_kernel_v3|tc_wgrad" -s 16 -c 9 -o gpurun_out/r1s_top_kernels $CMD > gpurun_out/r1s_ncu_full.log 2>&1
echo "ncu full exit $?"
