#!/bin/bash
# round 2, GPU call 35 (1 GPU): the slot's base address pinned in registers for the emission (no rebuild in front of every store group)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for lib in "" _pin; do
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 200 python tools/ab_sweep.py --reps 4 ""
done > $O/r2zi_sweep_pin.txt 2>&1
SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200_pin.so timeout 200 python tools/ab_sweep.py --reps 3 --input source --nfrag 8192 "" >> $O/r2zi_sweep_pin.txt 2>&1
cat $O/r2zi_sweep_pin.txt
