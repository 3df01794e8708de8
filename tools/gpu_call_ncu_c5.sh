#!/bin/bash
# ncu --set full of the two tile-stage kernels of bench c5 (the same command exited 0 without ncu in the previous call)
set -u
O=gpurun_out
mkdir -p $O
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"hilbert_tile_tma2_kernel" -s 6 -c 2 -o $O/r2w_prof_c5 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu > $O/r2w_ncu_c5.log 2>&1
echo "ncu rc=$?"
ls -la $O/r2w_prof_c5.ncu-rep
