#!/bin/bash
# round 2, GPU call 25 (1 GPU): two micro-variants of the evaluation (no gather for untrusted lanes; one find-first-set)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for lib in "" _v1 _v2 _v12; do
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py ""
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py --input source --nfrag 8192 ""
done > $O/r2y_sweep_micro.txt 2>&1; cat $O/r2y_sweep_micro.txt
