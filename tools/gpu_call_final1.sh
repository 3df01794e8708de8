#!/bin/bash
# Round-2 evidence call on one B200: full GPU test suite, every bench line with clocks, launch lists, ncu --set full of every default
# kernel of the path (source import on), codecs, reference arm.  Results are copied from gpurun_out/ into profiles/r02_*.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2z_pytest_gpu.log
tail -3 $O/r2z_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2z_bench_default.json 2> $O/r2z_bench_default.err
for wl in c1 c2 c4 c5 fill; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 > $O/r2z_bench_$wl.json 2> $O/r2z_bench_$wl.err
done
CNIIC_STAGES_NO_TMA=1 timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu > $O/r2z_bench_c5_plain_loads.json 2> /dev/null
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2z_bench_reference_arm.json 2> $O/r2z_bench_reference_arm.err
timeout 300 python tools/bench_stages.py > $O/r2z_stages.jsonl 2> $O/r2z_stages.err
timeout 300 python tools/bench_codecs.py > $O/r2z_codecs.jsonl 2> $O/r2z_codecs.err
for wl in c3 c2 c5 fill; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2z_launches_$wl.csv \
      python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > $O/r2z_ncu_$wl.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"km_assign_xyrgb_cull2|km_finalize" -s 12 -c 4 -o $O/r2z_prof_c3 \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu > $O/r2z_prof_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"km_assign_rgb_cull2|dedup_hist_kernel|dedup_compact_kernel|dedup_count_kernel" -s 3 -c 8 -o $O/r2z_prof_c2 \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu > $O/r2z_prof_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hilbert_tile_tma_kernel" -s 6 -c 2 -o $O/r2z_prof_c5 \
    python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu > $O/r2z_prof_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fill_kernel" -s 4 -c 1 -o $O/r2z_prof_fill \
    python bench.py --workload fill --steps 1 --warmup 1 --no-cpu > $O/r2z_prof_fill.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); g.smoke_extras()" > $O/r2z_smoke.log 2>&1
cuobjdump -sass cniic_b200/libcniic_b200.so | grep -E "UTMALDG|UBLKCP|SYNCS|ACQBULK|IDP|RED\.E" | awk '{print $2}' | sort | uniq -c > $O/r2z_sass_mnemonics.txt
ls -la $O | grep r2z | wc -l
