#!/bin/bash
# round 2, GPU call 7 (1 GPU): two-window round, pageable path tuning, full tests.  Every step under `timeout`.
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2g_pytest.log
tail -8 $O/r2g_pytest.log
timeout 300 python tools/ab_sweep.py "two=0" "two=1" "two=2" "two=3" "two=3,l2_chains=12" "two=1,l2_chains=12" > $O/r2g_sweep_two.txt 2>&1; cat $O/r2g_sweep_two.txt
timeout 300 python tools/ab_sweep.py --input source --nfrag 16384 "two=0" "two=1" "two=2" "two=3" > $O/r2g_sweep_two_source.txt 2>&1; cat $O/r2g_sweep_two_source.txt
timeout 200 python tools/trace_frags.py "two=3" > $O/r2g_trace_two.txt 2>&1; cat $O/r2g_trace_two.txt
timeout 120 python tools/host_register_probe.py > $O/r2g_hostreg.txt 2>&1; cat $O/r2g_hostreg.txt
for t in 4 8 12 16; do SNAPPY_B200_COPY_THREADS=$t timeout 200 python bench.py --steps 3 --warmup 2 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('copy threads $t: pageable e2e %.2f GB/s (%.1f ms), pinned %.2f (%.1f ms)' % (d['e2e']['pageable']['value'], d['e2e']['pageable']['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"; done 2>&1 | tee $O/r2g_copy_threads.txt
ls -la $O
