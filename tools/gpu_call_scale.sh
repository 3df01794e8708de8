#!/bin/bash
# The driver's scaling run, rehearsed: 8-rank parity, then the default bench at N = 1, 2, 4, 8 back to back on one 8-GPU box.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -k "8 and 1-" > $O/r2t_pytest_dist_8.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2t_pytest_dist_8.log
tail -2 $O/r2t_pytest_dist_8.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-secondary > $O/r2t_scale_n1.json 2> $O/r2t_scale_n1.err
for n in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-secondary > $O/r2t_scale_n$n.json 2> $O/r2t_scale_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29568 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > $O/r2t_scale_n8.json 2> $O/r2t_scale_n8.err
CNIIC_TLOG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29569 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu --no-secondary > $O/r2t_tlog_n8.json 2> $O/r2t_tlog_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --workload fill --steps 10 --warmup 3 --no-cpu > $O/r2t_fill_n8.json 2> $O/r2t_fill_n8.err
ls $O | grep r2t
