#!/bin/bash
# A/B of the fused delta-histogram stage on one GPU: the stage tests, then bench c5 per variant
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "hist or delta or codec or bins or hilbert or golden or c5" > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2w_pytest.log
tail -3 $O/r2w_pytest.log
run() { # name, env assignments...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > $O/r2w_c5_$name.json 2> $O/r2w_c5_$name.err
  echo "$name rc=$?"
}
run v2 CNIIC_X=0
run cube15 CNIIC_HIST_CUBE_R=15
