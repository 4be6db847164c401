#!/bin/bash
# round 2, GPU call 16 (1 GPU): foreign-stream decode throughput, loopback cost of the multi-GPU path, captures of the multi kernels
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/time_foreign.py > $O/r2p_foreign.txt 2>&1; cat $O/r2p_foreign.txt
timeout 300 python tools/time_loopback.py 2 8 > $O/r2p_loopback.txt 2>&1; cat $O/r2p_loopback.txt
timeout 200 python tools/prof_multi.py > $O/r2p_plain_multi.log 2>&1; cat $O/r2p_plain_multi.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_scan_shards|k_assemble|k_check_header|k_pull|k_decode_serial_tiles" -c 12 -f -o $O/r2p_prof_multi python tools/prof_multi.py > $O/r2p_ncu_multi.log 2>&1; tail -2 $O/r2p_ncu_multi.log
