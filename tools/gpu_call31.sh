#!/bin/bash
# round 2, GPU call 31 (1 GPU): emission (flush) as a real function call instead of inlined at every record site
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for lib in "" _ni; do
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py ""
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/ab_sweep.py --input source --nfrag 8192 ""
done > $O/r2ze_sweep_noinline.txt 2>&1; cat $O/r2ze_sweep_noinline.txt
SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200_ni.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bit_exact or adversarial or sweep or string" 2>&1 | tail -2
