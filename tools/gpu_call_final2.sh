#!/bin/bash
# Final 1-GPU evidence call of round 2: the whole GPU suite, the bench lines, smoke(), the launch list of the default bench and an
# ncu --set full capture of the tile-stage kernels (each ncu pass repeats a command that has just exited 0 without ncu).
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2z_pytest.log
tail -2 $O/r2z_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2z_bench_default.json 2> $O/r2z_bench_default.err; echo "default rc=$?"
timeout 200 python bench.py --workload fill --steps 10 --warmup 3 --no-cpu > $O/r2z_bench_fill.json 2> $O/r2z_bench_fill.err; echo "fill rc=$?"
timeout 200 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/r2z_bench_c5.json 2> $O/r2z_bench_c5.err; echo "c5 rc=$?"
timeout 200 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > $O/r2z_bench_c4.json 2> $O/r2z_bench_c4.err; echo "c4 rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2z_smoke.log
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $O/r2z_bench_short.json 2> $O/r2z_bench_short.err; echo "short rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2z_launches_default.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > $O/r2z_ncu_default.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"hilbert_tile_tma2_kernel" -s 6 -c 2 -o $O/r2z_prof_c5 \
    python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu > $O/r2z_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
