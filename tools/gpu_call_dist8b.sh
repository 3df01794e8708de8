#!/bin/bash
# 8-GPU diagnosis: device timelines (CNIIC_TLOG=1) of the row-sharded C3 loop, default vs no-PDL vs NCCL exchange.
set -u
O=gpurun_out
mkdir -p $O
nproc > $O/r2g_nproc.txt
run() {  # N port extra-args...
  local n=$1 port=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 6 --warmup 3 --no-cpu --no-secondary "$@"
}
run 8 29531 > $O/r2g_bench_c3_n8.json 2> $O/r2g_bench_c3_n8.err
CNIIC_TLOG=1 run 8 29532 > $O/r2g_tlog_n8.json 2> $O/r2g_tlog_n8.err
CNIIC_TLOG=1 CNIIC_NO_PDL=1 run 8 29533 > $O/r2g_tlog_n8_nopdl.json 2> $O/r2g_tlog_n8_nopdl.err
CNIIC_TLOG=1 CNIIC_P2P=0 run 8 29534 > $O/r2g_tlog_n8_nccl.json 2> $O/r2g_tlog_n8_nccl.err
ls -la $O | grep r2g
