#!/bin/bash
# 8-GPU: parity at 8 ranks with the cell exchange, the default bench at N = 8 / 4, one timeline run.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -k "8 and 1-" > $O/r2i_pytest_dist_8.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2i_pytest_dist_8.log
tail -3 $O/r2i_pytest_dist_8.log
run() {  # N port extra-args...
  local n=$1 port=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --warmup 3 --no-cpu "$@"
}
run 8 29551 --steps 10 > $O/r2i_bench_default_n8.json 2> $O/r2i_bench_default_n8.err
run 4 29552 --steps 10 --no-secondary > $O/r2i_bench_c3_n4.json 2> $O/r2i_bench_c3_n4.err
CNIIC_TLOG=1 run 8 29553 --steps 4 --no-secondary > $O/r2i_tlog_n8.json 2> $O/r2i_tlog_n8.err
grep "tlog rank 0" $O/r2i_tlog_n8.err | tail -6
