#!/bin/bash
# A/B of the Hilbert tile stages on one GPU: the whole GPU suite on the new default, then bench c5 per variant
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2u_pytest.log
tail -3 $O/r2u_pytest.log
run() { # name, env assignments...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu > $O/r2u_c5_$name.json 2> $O/r2u_c5_$name.err
  echo "$name rc=$?"
}
run v2_nb5 CNIIC_X=0
run v2_nb6 CNIIC_TILE_NB6=1
run v1 CNIIC_TILE_V1=1
run cube14 CNIIC_HIST_CUBE_R=14
run cube11 CNIIC_HIST_CUBE_R=11
