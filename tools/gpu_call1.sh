#!/bin/bash
# round 2, GPU call 1: tests after the context refactor, option / build sweeps, captures of the remaining kernels
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2a_pytest.log
tail -3 $O/r2a_pytest.log
python tools/ab_sweep.py "lpt=0" "lpt=1" "lpt=1,l2_reserve=0" "lpt=1,l2_reserve=2" "lpt=1,l2_chains=12" "lpt=1,l2_chains=16" \
   "lpt=1,ring_smem=4096" "lpt=1,ring_l2=2048" "lpt=0,l2_chains=0" "lpt=1,l2_chains=0" "lpt=1,smem_chains=0" "lpt=0,smem_chains=0" > $O/r2a_sweep_main.txt 2>&1
cat $O/r2a_sweep_main.txt
for v in cs exp; do
  SNAPPY_B200_LIB=/root/repo/snappy.jl_b200/libsnappy_b200_$v.so python tools/ab_sweep.py "lpt=0" "lpt=1" > $O/r2a_sweep_$v.txt 2>&1
  cat $O/r2a_sweep_$v.txt
done
python tools/ab_sweep.py --input source --nfrag 4096 "lpt=0" "lpt=1" "lpt=1,l2_chains=10" > $O/r2a_sweep_source.txt 2>&1
cat $O/r2a_sweep_source.txt
python bench.py --steps 5 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"
python tools/show_bench.py $O/r2a_bench.json 2>/dev/null || head -c 600 $O/r2a_bench.json
# captures
python tools/prof_misc.py > $O/r2a_plain_misc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_parse|k_build_index|k_compress_pages|k_decode_pages|k_compact|k_scan|k_estimate|k_order' \
    -f -o $O/r2a_prof_misc python tools/prof_misc.py > $O/r2a_ncu_misc.log 2>&1
tail -2 $O/r2a_ncu_misc.log
python tools/prof_run.py 16384 0 profile_range=1 > $O/r2a_plain_range.log 2>&1 && \
ncu --replay-mode range --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
    -f -o $O/r2a_prof_range python tools/prof_run.py 16384 0 profile_range=1 > $O/r2a_ncu_range.log 2>&1
tail -5 $O/r2a_ncu_range.log
ls -la $O | tail -20
