#!/bin/bash
# round 2, GPU call 33 (1 GPU): warps per SM once more, after the instruction trims (the balance may have moved)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/ab_sweep.py "" "l2_chains=13" "l2_chains=15" "l2_chains=16" "l2_chains=12" > $O/r2zg_sweep_warps.txt 2>&1; cat $O/r2zg_sweep_warps.txt
