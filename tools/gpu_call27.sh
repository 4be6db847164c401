#!/bin/bash
# round 2, GPU call 27 (1 GPU): capture of the compress kernel in its final form (traffic.json), launch list of the bench command
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/prof_run.py 16384 0 > $O/r2za_plain.log 2>&1; cat $O/r2za_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compress_window_mixed -s 1 -c 1 -f -o $O/r2za_prof_compress python tools/prof_run.py 16384 0 > $O/r2za_ncu.log 2>&1; tail -2 $O/r2za_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2za_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > $O/r2za_ncu_bench.log 2>&1; echo "launch list rc=$?"
