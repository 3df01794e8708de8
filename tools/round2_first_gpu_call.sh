#!/bin/bash
# First GPU call of round 2: the evidence gap VERDICT r01 item 1 lists, in one go.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_gpu_call.sh'
# Outputs land in gpurun_out/ (what should be judged is copied into profiles/ afterwards).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest_gpu.log
# bench lines (with clocks) for every workload; c2 carries the cpu_baseline leg, the others skip it to save box time
python bench.py --workload c2 --steps 10 --warmup 3 > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err
for wl in c3 c1 c4 c5 fill; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu > $O/r2_bench_$wl.json 2> $O/r2_bench_$wl.err
done
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref_c2.json 2>&1
# A/B of the kernel versions on the two Lloyd workloads
CNIIC_RGB_CULL_V1=1 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > $O/r2_bench_c2_v1.json 2>&1
CNIIC_XY_CULL_V1=1 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu > $O/r2_bench_c3_v1.json 2>&1
python tools/bench_codecs.py > $O/r2_codecs.jsonl 2> $O/r2_codecs.err
# launch lists (cold-cache, serialised: shares, not absolutes)
for wl in c2 c3 c4 c5 fill; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r2_launches_$wl.csv \
      python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > $O/r2_ncu_$wl.log 2>&1
done
# full captures of the kernels that are actually the default
ncu --set full --clock-control none --import-source on -k regex:km_assign_rgb_cull2 -s 12 -c 2 -o $O/r2_prof_c2_cull2 \
    python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu > $O/r2_prof_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:km_assign_xyrgb_cull2 -s 12 -c 2 -o $O/r2_prof_c3_cull2 \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu > $O/r2_prof_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fill_kernel|fill_supercull_kernel" -s 4 -c 2 -o $O/r2_prof_fill \
    python bench.py --workload fill --steps 1 --warmup 1 --no-cpu > $O/r2_prof_fill.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hilbert_tile_kernel -s 4 -c 2 -o $O/r2_prof_c5 \
    python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu > $O/r2_prof_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"km_assign_rgb.*batch" -s 4 -c 1 -o $O/r2_prof_c4_batch \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > $O/r2_prof_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hd_sync_kernel|hd_write_kernel|rle_emit_kernel|pack_write_kernel" -c 6 -o $O/r2_prof_codecs \
    python tools/bench_codecs.py > $O/r2_prof_codecs.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/r2_gpu.txt; nproc >> $O/r2_gpu.txt
ls -la $O | tail -40
