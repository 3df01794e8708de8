#!/bin/bash
# 8-GPU call: 4- and 8-rank parity (both exchanges), then the default bench (C3 strong + secondary C2) at N = 8 and 4, and the NCCL
# exchange at N = 8 for comparison.    gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_call_dist8.sh'
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r2f_gpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "8 or 4" > $O/r2f_pytest_dist_4_8.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2f_pytest_dist_4_8.log
tail -4 $O/r2f_pytest_dist_4_8.log
run() {  # N port extra-args...
  local n=$1 port=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 --no-cpu "$@"
}
CNIIC_BENCH_DEBUG=1 run 8 29521 > $O/r2f_bench_default_n8.json 2> $O/r2f_bench_default_n8.err
run 4 29522 --no-secondary > $O/r2f_bench_c3_n4.json 2> $O/r2f_bench_c3_n4.err
CNIIC_P2P=0 run 8 29523 --no-secondary > $O/r2f_bench_c3_n8_nccl.json 2> $O/r2f_bench_c3_n8_nccl.err
run 8 29524 --workload c4 --steps 5 > $O/r2f_bench_c4_n8.json 2> $O/r2f_bench_c4_n8.err
run 8 29525 --workload fill > $O/r2f_bench_fill_n8.json 2> $O/r2f_bench_fill_n8.err
run 8 29526 --workload c5 > $O/r2f_bench_c5_n8.json 2> $O/r2f_bench_c5_n8.err
ls -la $O | grep r2f
