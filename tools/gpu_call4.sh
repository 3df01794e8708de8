#!/bin/bash
# 1-GPU call: parity + timings after folding the level-1 cull into the D=5 assign kernel, the Morton-ordered unique-colour path,
# the 32-bit weighted accumulation and the three-level fill.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2d_pytest_gpu.log
tail -4 $O/r2d_pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2d_bench_default.json 2> $O/r2d_bench_default.err
python bench.py --workload fill --steps 10 --warmup 3 --no-cpu > $O/r2d_bench_fill.json 2> $O/r2d_bench_fill.err
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > $O/r2d_bench_c4.json 2> $O/r2d_bench_c4.err
for wl in c2 c3 fill; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2d_launches_$wl.csv \
      python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > $O/r2d_ncu_$wl.log 2>&1
done
ls -la $O | grep r2d
