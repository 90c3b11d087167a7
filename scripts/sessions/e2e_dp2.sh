for o in 0 1; do DCGC_NO_OVERLAP=$o timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/e2e_trace_dp.py 2>&1 | grep "^rank 0"; done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29517 scripts/e2e_trace_dp.py 2>&1 | grep "^rank 0"
