#!/bin/bash
# 2-GPU call: parity of the multi-CTA push-exchange update kernel + first strong-scaling numbers of the default bench.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_call_dist2.sh'
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2b_pytest_gpu_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2b_pytest_gpu_2gpu.log
tail -5 $O/r2b_pytest_gpu_2gpu.log
python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu > $O/r2b_bench_c3_n1.json 2> $O/r2b_bench_c3_n1.err
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > $O/r2b_bench_c2_n1.json 2> $O/r2b_bench_c2_n1.err
CNIIC_BENCH_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/r2b_bench_default_n2.json 2> $O/r2b_bench_default_n2.err
CNIIC_P2P=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-secondary > $O/r2b_bench_c3_n2_nccl.json 2> $O/r2b_bench_c3_n2_nccl.err
tail -3 $O/*.err
