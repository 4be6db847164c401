#!/bin/bash
# ncu evidence for profiles/: run on the GPU box AFTER the plain commands have exited 0.
#   bash tools/capture_profiles.sh r01b
set -x
tag=${1:-r01b}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_bench_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_$tag.log 2>&1
python tools/prof_run.py 16384 0 l2_chains=0 > gpurun_out/plain_prof_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_compress_window -s 1 -c 1 -f \
    -o gpurun_out/prof_window_smem_$tag python tools/prof_run.py 16384 0 l2_chains=0 > gpurun_out/ncu_a_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_compress_window -s 1 -c 1 -f \
    -o gpurun_out/prof_window_l2_$tag python tools/prof_run.py 16384 0 smem_chains=0 > gpurun_out/ncu_b_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_fragments -s 1 -c 1 -f \
    -o gpurun_out/prof_decode_$tag python tools/prof_run.py 16384 0 > gpurun_out/ncu_c_$tag.log 2>&1
for f in a b c; do tail -n 2 gpurun_out/ncu_${f}_$tag.log; done
