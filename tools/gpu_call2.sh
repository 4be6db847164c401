#!/bin/bash
# round 2, GPU call 2 (1 GPU): new tests (loopback worlds, configs at size), fragment traces, ring / warp sweeps,
# the restructured bench, small captures
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2b_pytest.log
tail -15 $O/r2b_pytest.log
python tools/trace_frags.py "" "lpt=0" "l2_chains=0" "smem_chains=0" > $O/r2b_trace.txt 2>&1; cat $O/r2b_trace.txt
python tools/ab_sweep.py "" "smem_chains=5,l2_chains=12,ring_l2=4096" "smem_chains=5,l2_chains=14,ring_l2=2048" \
  "smem_chains=5,l2_chains=16,ring_l2=2048" "smem_chains=5,l2_chains=20,ring_l2=1024" "smem_chains=5,ring_smem=8192" \
  "smem_chains=4,l2_chains=20,ring_l2=2048" "smem_chains=5,l2_chains=18,ring_l2=1024" > $O/r2b_sweep_rings.txt 2>&1; cat $O/r2b_sweep_rings.txt
python tools/trace_frags.py --input source --nfrag 4096 "" > $O/r2b_trace_source.txt 2>&1; cat $O/r2b_trace_source.txt
( time python bench.py --steps 5 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -5 $O/r2b_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
    print("C2 value %.2f compress %.2f uncompress %.1f e2e %.2f pageable %s gen %.1fs" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["e2e"]["value"], d["e2e"].get("pageable"), d["input_generation_s"]))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], "compress", v.get("compress_gbps"), "uncompress", v.get("uncompress_gbps"), "ms/step %.2f" % v["ms_per_step"], v["roofline"]["kernel"], "%.2f ms" % v["roofline"]["kernel_ms"])
except Exception as e:
    print("bench parse failed", e)
PY
# captures (small): the concurrent compress pair as one range; the remaining kernels on a small input
python tools/prof_run.py 16384 0 profile_range=1 > $O/r2b_plain_range.log 2>&1 && \
ncu --replay-mode range --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    -f -o $O/r2b_prof_range python tools/prof_run.py 16384 0 profile_range=1 > $O/r2b_ncu_range.log 2>&1
tail -3 $O/r2b_ncu_range.log
python tools/prof_misc.py 512 16384 > $O/r2b_plain_misc.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_parse|k_build_index|k_compress_pages|k_decode_pages|k_compact|k_scan|k_estimate|k_order' \
    -c 60 -f -o $O/r2b_prof_misc python tools/prof_misc.py 512 16384 > $O/r2b_ncu_misc.log 2>&1
tail -2 $O/r2b_ncu_misc.log
du -sh $O; ls -la $O
# keep the transfer below the 64 MiB limit
if [ $(du -sm $O | cut -f1) -gt 60 ]; then rm -f $O/r2b_prof_misc.ncu-rep; echo "dropped r2b_prof_misc.ncu-rep (too large)"; fi
