#!/bin/bash
# round 2, GPU call 8 (1 GPU): the pipelined round (compress_pipe.cuh), option pipe = 0..3
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_sweep.py "pipe=0" "pipe=1" "pipe=2" "pipe=3" "pipe=3,l2_chains=12" "pipe=3,l2_chains=16" "pipe=1,l2_chains=12" > $O/r2h_sweep_pipe.txt 2>&1; cat $O/r2h_sweep_pipe.txt
timeout 300 python tools/ab_sweep.py --input source --nfrag 16384 "pipe=0" "pipe=1" "pipe=2" "pipe=3" > $O/r2h_sweep_pipe_source.txt 2>&1; cat $O/r2h_sweep_pipe_source.txt
timeout 200 python tools/trace_frags.py "pipe=3" > $O/r2h_trace_pipe.txt 2>&1; cat $O/r2h_trace_pipe.txt
