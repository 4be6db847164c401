#!/bin/bash
# round 2, GPU call 17 (1 GPU): source-level capture of the indexed decoder (current build)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/prof_run.py 16384 0 > $O/r2q_plain.log 2>&1; cat $O/r2q_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decode_fragments -s 1 -c 1 -f -o $O/r2q_prof_decode python tools/prof_run.py 16384 0 > $O/r2q_ncu.log 2>&1; tail -2 $O/r2q_ncu.log
