#!/bin/bash
# GPU session r2p: int8 feature upload + chunked H2D: tests, e2e trace sweep, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_staged.py -x -q 2>&1 | tail -3
for c in 0 1 0.25; do for i8 in 0 1; do echo "chunk $c i8 $i8"; DCGC_FEATURES_I8=$i8 DCGC_H2D_CHUNK_MB=$c timeout 200 python scripts/e2e_trace.py 4 2>&1 | grep -E "workers=|main-stream|Error|error"; done; done
timeout 600 python bench.py --breakdown gpurun_out/r2p_breakdown.md > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; echo "bench n1 exit $?"
show='import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("n=%d value %.0f ms %.3f e2e %.0f e2e_ms %.3f h2d %d roof %.3f avg_us %.2f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))'
cat gpurun_out/r2p_bench_n1.json | python -c "$show"
tail -3 gpurun_out/r2p_bench_n1.err
