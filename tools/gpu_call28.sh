#!/bin/bash
# round 2, GPU call 28 (1 GPU): L2 policy made once per fragment instead of at every table access; trimmed far gather
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for lib in _old "" _trim; do
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py ""
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py --input source --nfrag 8192 ""
done > $O/r2zb_sweep_policy.txt 2>&1; cat $O/r2zb_sweep_policy.txt
