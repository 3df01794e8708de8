#!/bin/bash
# 1-GPU call: TMA-fed tile stages (A/B against plain loads), the smem-resident three-level fill, captures of both.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2e_pytest_gpu.log
tail -4 $O/r2e_pytest_gpu.log
for wl in c5 fill; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 > $O/r2e_bench_$wl.json 2> $O/r2e_bench_$wl.err
done
CNIIC_STAGES_NO_TMA=1 timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/r2e_bench_c5_notma.json 2> $O/r2e_bench_c5_notma.err
timeout 300 python tools/bench_stages.py > $O/r2e_stages.jsonl 2> $O/r2e_stages.err
timeout 300 python tools/bench_codecs.py > $O/r2e_codecs.jsonl 2> $O/r2e_codecs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hilbert_tile_tma_kernel|fill_kernel" -s 6 -c 4 -o $O/r2e_prof_stages \
    python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu > $O/r2e_prof_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fill_kernel" -s 4 -c 2 -o $O/r2e_prof_fill \
    python bench.py --workload fill --steps 1 --warmup 1 --no-cpu > $O/r2e_prof_fill.log 2>&1
ls -la $O | grep r2e
