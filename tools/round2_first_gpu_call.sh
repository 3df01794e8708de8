#!/bin/bash
# First GPU call of the next round: everything that was written without a GPU, measured in one go.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_gpu_call.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r2_sanitize.log 2>&1; echo "sanitize rc=$?" >> gpurun_out/r2_sanitize.log
for wl in c2 c4 c3 c1; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_c2.json 2>&1
python tools/bench_codecs.py > gpurun_out/r2_codecs.jsonl 2> gpurun_out/r2_codecs.err
# launch list of the default bench, then full captures of the kernels that are new (batch assign, Huffman decode sweeps, RLE emit)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c4.csv \
    python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:km_assign_rgb_batch -c 1 -o gpurun_out/r2_prof_c4_batch \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_prof_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hd_sync_kernel|hd_write_kernel|rle_emit_kernel" -c 6 -o gpurun_out/r2_prof_codecs \
    python tools/bench_codecs.py > gpurun_out/r2_prof_codecs.log 2>&1
for wl in c5 fill; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err
done
# A/B of the kernel versions and of the sort width on the headline workload
CNIIC_RGB_CULL_V1=1 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_c2_v1.json 2>&1
CNIIC_SORT_BITS=16 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_c2_sort16.json 2>&1
CNIIC_XY_CULL_V1=1 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_c3_v1.json 2>&1
ls -la gpurun_out | tail -30
# later, on 2 GPUs (gpurun --gpus 2): pytest tests/test_gpu_dist.py; torchrun ... bench.py --gpus 2 for c2, c3, c4, c5, fill
