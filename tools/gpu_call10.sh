#!/bin/bash
# round 2, GPU call 10 (1 GPU): why is the pipelined round slower?  instruction counts + a source-level capture
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_sweep.py "pipe=0,l2_chains=0" "pipe=1,l2_chains=0" "pipe=0,smem_chains=0,l2_chains=8" "pipe=2,smem_chains=0,l2_chains=8" "pipe=0,l2_fetch=32" "pipe=0,l2_fetch=128" "pipe=0,l2_fetch=64" > $O/r2j_sweep.txt 2>&1; cat $O/r2j_sweep.txt
M=smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
for opt in "pipe=0" "pipe=1" "pipe=3" "pipe=0 l2_chains=0" "pipe=1 l2_chains=0"; do
  echo "== $opt"
  timeout 300 ncu --metrics $M --clock-control none -k regex:k_compress_window -s 1 -c 1 --csv python tools/prof_run.py 16384 0 $opt 2>&1 | grep -E "k_compress_window" | awk -F'","' '{print $(NF-2), $NF}'
done > $O/r2j_counts.txt 2>&1; cat $O/r2j_counts.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_compress_window -s 1 -c 1 -f -o $O/r2j_prof_pipe_smem python tools/prof_run.py 16384 0 pipe=1 l2_chains=0 > $O/r2j_ncu.log 2>&1; tail -2 $O/r2j_ncu.log
