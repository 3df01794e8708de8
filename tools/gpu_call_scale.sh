#!/bin/bash
# The driver's scaling run, rehearsed: the default bench at N = 1, 2, 4, 8 back to back on one 8-GPU box.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-secondary > $O/r2s_scale_n1.json 2> $O/r2s_scale_n1.err
for n in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-secondary > $O/r2s_scale_n$n.json 2> $O/r2s_scale_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29568 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > $O/r2s_scale_n8.json 2> $O/r2s_scale_n8.err
CNIIC_TLOG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29569 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu --no-secondary > $O/r2s_tlog_n8.json 2> $O/r2s_tlog_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29570 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 --no-secondary > $O/r2s_ref_n8.json 2> $O/r2s_ref_n8.err
ls $O | grep r2s
