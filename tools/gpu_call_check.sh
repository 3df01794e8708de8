#!/bin/bash
# quick 1-GPU confirmation: the whole GPU suite, the default bench line, fill and c5
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2x_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2x_pytest.log
tail -2 $O/r2x_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2x_bench_default.json 2> $O/r2x_bench_default.err
timeout 300 python bench.py --workload fill --steps 10 --warmup 3 --no-cpu > $O/r2x_bench_fill.json 2> $O/r2x_bench_fill.err
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/r2x_bench_c5.json 2> $O/r2x_bench_c5.err
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > $O/r2x_bench_c4.json 2> $O/r2x_bench_c4.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2x_smoke.log 2>&1; tail -1 $O/r2x_smoke.log
