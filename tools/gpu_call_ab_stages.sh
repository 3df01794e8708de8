#!/bin/bash
# the Hilbert tile stages on one GPU: the stage tests, then bench c5
set -u
O=gpurun_out
mkdir -p $O
timeout 150 python -m pytest tests -m gpu -q -x -k "hist or delta or codec or bins or hilbert or golden or c5 or stages" > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2g_pytest.log
tail -3 $O/r2g_pytest.log
timeout 200 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/r2g_c5.json 2> $O/r2g_c5.err; echo "c5 rc=$?"
