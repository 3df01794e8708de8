#!/bin/bash
# 2-GPU check of the self-validating-cell exchange: parity (both exchanges), timeline, bench.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -x -q -k "dist or xyrgb or kmeans_rgb_per_pixel" > $O/r2h_pytest_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2h_pytest_2gpu.log
tail -4 $O/r2h_pytest_2gpu.log
CNIIC_TLOG=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu --no-secondary > $O/r2h_tlog_n2.json 2> $O/r2h_tlog_n2.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/r2h_bench_default_n2.json 2> $O/r2h_bench_default_n2.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary > $O/r2h_bench_c3_n1.json 2> $O/r2h_bench_c3_n1.err
grep "tlog rank 0" $O/r2h_tlog_n2.err | tail -5
