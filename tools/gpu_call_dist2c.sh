#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_dist.py -x -q -k "2" > $O/r2w_pytest_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2w_pytest_2gpu.log
tail -3 $O/r2w_pytest_2gpu.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-secondary > $O/r2w_bench_n2.json 2> $O/r2w_bench_n2.err
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2w_bench_n1.json 2> $O/r2w_bench_n1.err
