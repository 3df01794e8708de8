#!/bin/bash
# A/B of the Hilbert tile stages on one GPU: the stage tests, then bench c5 (default kernels, first TMA version)
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "hist or delta or codec or bins or hilbert or golden or c5 or stages" > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2y_pytest.log
tail -3 $O/r2y_pytest.log
run() { # name, env assignments...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > $O/r2y_c5_$name.json 2> $O/r2y_c5_$name.err
  echo "$name rc=$?"
}
run v2 CNIIC_X=0
run v1 CNIIC_TILE_V1=1
