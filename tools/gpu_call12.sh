#!/bin/bash
# round 2, GPU call 12 (1 GPU): one copy of the round's code for both table placements (option unified)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/ab_sweep.py "unified=0" "unified=1" "unified=1,pipe=3" "unified=1,l2_chains=12" "unified=1,l2_chains=16" "unified=1,pipe=3,l2_chains=10" "unified=1,pipe=3,l2_chains=8" > $O/r2l_sweep_unified.txt 2>&1; cat $O/r2l_sweep_unified.txt
timeout 300 python tools/ab_sweep.py --input source --nfrag 8192 "unified=0" "unified=1" "unified=1,pipe=3" "unified=1,l2_chains=10" > $O/r2l_sweep_unified_source.txt 2>&1; cat $O/r2l_sweep_unified_source.txt
R=smsp__average_warps_issue_stalled
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__average_warp_latency_per_inst_issued.ratio
for s in long_scoreboard short_scoreboard wait not_selected branch_resolving no_instruction math_pipe_throttle; do M=$M,${R}_${s}_per_issue_active.ratio; done
for opt in "unified=1" "unified=1 pipe=3"; do
  echo "== $opt"
  timeout 300 ncu --metrics $M --clock-control none -k regex:k_compress_window -s 1 -c 1 --csv python tools/prof_run.py 16384 0 $opt 2>&1 | grep -E "k_compress_window" | awk -F'","' '{print $(NF-2), $NF}' | sed 's/smsp__average_warps_issue_stalled_//; s/_per_issue_active.ratio//; s/"//'
done > $O/r2l_stalls.txt 2>&1; cat $O/r2l_stalls.txt
