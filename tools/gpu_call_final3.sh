#!/bin/bash
# last 1-GPU confirmation of the round's final tree: the whole GPU suite, smoke(), bench c5
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2f_pytest.log
tail -2 $O/r2f_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r2f_smoke.log
timeout 150 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/r2f_bench_c5.json 2> $O/r2f_bench_c5.err; echo "c5 rc=$?"
