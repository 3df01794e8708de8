#!/bin/bash
# final-tree sanity of the driver's N > 1 bench launch on two GPUs (the Lloyd path itself is unchanged since tools/gpu_call_scale.sh)
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu \
    > $O/r2n_scale_n2.json 2> $O/r2n_scale_n2.err; echo "n2 rc=$?"
tail -c 600 $O/r2n_scale_n2.json
