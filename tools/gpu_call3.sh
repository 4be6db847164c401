#!/bin/bash
# round 2, GPU call 3 (1 GPU): merged compress kernel (warp-slot priority), window-round page kernel, full test suite
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2c_pytest.log
tail -8 $O/r2c_pytest.log
python tools/ab_sweep.py "mixed=0" "mixed=1" "mixed=2" "mixed=0,l2_first=1" "mixed=1,l2_chains=12" "mixed=1,l2_chains=16" "mixed=1,smem_chains=5,l2_chains=16,ring_l2=2048" "mixed=1,l2_reserve=2" > $O/r2c_sweep_mixed.txt 2>&1; cat $O/r2c_sweep_mixed.txt
python tools/trace_frags.py "mixed=1" "mixed=2" "mixed=0" "mixed=0,l2_first=1" > $O/r2c_trace.txt 2>&1; cat $O/r2c_trace.txt
python tools/ab_sweep.py --input source --nfrag 16384 "mixed=0" "mixed=1" "mixed=1,l2_chains=10" > $O/r2c_sweep_source.txt 2>&1; cat $O/r2c_sweep_source.txt
python bench.py --config 4 --steps 3 --warmup 2 > $O/r2c_bench_c4.json 2> $O/r2c_bench_c4.err; echo "bench c4 rc=$?"; tail -3 $O/r2c_bench_c4.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench_c4.json").read().strip().splitlines()[-1])
    print("C4 value %.2f compress %.2f uncompress %.2f kernel %s %.2f ms decode %.2f ms" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["roofline"]["kernel"], d["roofline"]["kernel_ms"], d["other_kernel"]["k_decode_pages_ms"]))
except Exception as e:
    print("c4 parse failed", e)
PY
SNAPPY_B200_OPTIONS=pages_window=0 python bench.py --config 4 --steps 3 --warmup 2 --pages 262144 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 serial kernel (262144 pages): compress %.2f GB/s' % d['compress_gbps'])"
python bench.py --config 4 --steps 3 --warmup 2 --pages 262144 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 window kernel (262144 pages): compress %.2f GB/s' % d['compress_gbps'])"
# the merged kernel as ONE ncu capture (what the two concurrent kernels never allowed)
python tools/prof_run.py 16384 0 > $O/r2c_plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_compress_window_mixed -s 1 -c 1 -f \
    -o $O/r2c_prof_mixed python tools/prof_run.py 16384 0 > $O/r2c_ncu_mixed.log 2>&1
tail -2 $O/r2c_ncu_mixed.log
python tools/prof_misc.py 512 16384 > $O/r2c_plain_misc.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_parse|k_build_index|k_compress_pages|k_decode_pages|k_compact|k_scan|k_estimate|k_order' \
    -c 34 -f -o $O/r2c_prof_misc python tools/prof_misc.py 512 16384 > $O/r2c_ncu_misc.log 2>&1
tail -2 $O/r2c_ncu_misc.log
du -sh $O; ls -la $O
if [ $(du -sm $O | cut -f1) -gt 60 ]; then rm -f $O/r2c_prof_misc.ncu-rep; echo "dropped r2c_prof_misc.ncu-rep (too large)"; fi
