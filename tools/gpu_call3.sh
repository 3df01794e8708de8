#!/bin/bash
# 1-GPU call: parity after PDL + unique-colour path, default bench (C3 + secondary C2), PDL A/B, launch lists and captures of the new kernels.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2c_pytest_gpu.log
tail -4 $O/r2c_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/r2c_bench_default.json 2> $O/r2c_bench_default.err
CNIIC_NO_PDL=1 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2c_bench_default_nopdl.json 2> $O/r2c_bench_default_nopdl.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c_bench_ref_default.json 2> $O/r2c_bench_ref_default.err
python tools/bench_codecs.py > $O/r2c_codecs.jsonl 2> $O/r2c_codecs.err
for wl in c2 c3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2c_launches_$wl.csv \
      python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > $O/r2c_ncu_$wl.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"km_assign_rgb_cull2|dedup_hist_kernel|dedup_compact_kernel" -s 6 -c 4 -o $O/r2c_prof_c2_unique \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu > $O/r2c_prof_c2.log 2>&1
ls -la $O | grep r2c
