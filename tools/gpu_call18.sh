#!/bin/bash
# round 2, GPU call 18 (1 GPU): decoder -- stream prefetch distance / level, 64 resident warps per SM (decode_occupancy 16)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for lib in "" _dpf0 _dpf256 _dpf1024 _dpfL1; do
  echo "== libsnappy_b200$lib.so"
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python tools/time_decode.py 12,16,10
done > $O/r2r_decode_variants.txt 2>&1; cat $O/r2r_decode_variants.txt
