#!/bin/bash
# round 2, GPU call 9 (1 GPU): pipelined round with fewer warps in flight (does the working set in L2 change the picture?)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/ab_sweep.py "pipe=0" "pipe=0,l2_chains=0" "pipe=1,l2_chains=0" "pipe=0,l2_chains=4" "pipe=1,l2_chains=4" "pipe=3,l2_chains=4" "pipe=0,l2_chains=8" "pipe=1,l2_chains=8" "pipe=3,l2_chains=8" "pipe=3,l2_chains=10" "pipe=0,smem_chains=0,l2_chains=8" "pipe=2,smem_chains=0,l2_chains=8" > $O/r2i_sweep_pipe_warps.txt 2>&1; cat $O/r2i_sweep_pipe_warps.txt
timeout 300 python tools/ab_sweep.py --input source --nfrag 8192 "pipe=0" "pipe=0,l2_chains=0" "pipe=1,l2_chains=0" "pipe=1,l2_chains=4" "pipe=3,l2_chains=4" "pipe=0,l2_chains=8" "pipe=3,l2_chains=8"> $O/r2i_sweep_pipe_warps_source.txt 2>&1; cat $O/r2i_sweep_pipe_warps_source.txt
