#!/bin/bash
# round 2, GPU call 11 (1 GPU): stall reasons of the pipelined round in the default warp configuration
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
R=smsp__average_warps_issue_stalled
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warp_latency_per_inst_issued.ratio
for s in long_scoreboard short_scoreboard wait not_selected branch_resolving no_instruction lg_throttle mio_throttle math_pipe_throttle barrier membar dispatch_stall drain imc_miss sleeping tex_throttle misc selected; do M=$M,${R}_${s}_per_issue_active.ratio; done
for opt in "pipe=0" "pipe=1" "pipe=2" "pipe=3"; do
  echo "== $opt"
  timeout 300 ncu --metrics $M --clock-control none -k regex:k_compress_window -s 1 -c 1 --csv python tools/prof_run.py 16384 0 $opt 2>&1 | grep -E "k_compress_window" | awk -F'","' '{print $(NF-2), $NF}' | sed 's/smsp__average_warps_issue_stalled_//; s/_per_issue_active.ratio//; s/"//'
done > $O/r2k_stalls.txt 2>&1; cat $O/r2k_stalls.txt
