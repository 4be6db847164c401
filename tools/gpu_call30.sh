#!/bin/bash
# round 2, GPU call 30 (1 GPU): hop loop with one test per copy (the descriptor knows where its copy lands)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for lib in _old ""; do
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py ""
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py --input source --nfrag 8192 ""
done > $O/r2zd_sweep_hop.txt 2>&1; cat $O/r2zd_sweep_hop.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
