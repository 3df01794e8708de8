#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "xyrgb or hilbert or delta or codec or golden or kmeans_rgb_per_pixel" > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2y_pytest.log
tail -2 $O/r2y_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2y_bench_default.json 2> $O/r2y_bench_default.err
timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/r2y_bench_c5.json 2> $O/r2y_bench_c5.err
timeout 300 python tools/bench_stages.py > $O/r2y_stages.jsonl 2> $O/r2y_stages.err
