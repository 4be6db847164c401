#!/bin/bash
# round 2, GPU call 4 (1 GPU): foreign / corrupt stream paths, cache-policy and prefetch build variants
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_foreign.py tests/test_gpu_parity.py -m gpu -q -x > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2d_pytest.log
tail -12 $O/r2d_pytest.log
python tools/ab_sweep.py "" > $O/r2d_sweep_base.txt 2>&1; cat $O/r2d_sweep_base.txt
for v in pf1 pf2 el1 el2 ef; do
  SNAPPY_B200_LIB=/root/repo/snappy.jl_b200/libsnappy_b200_$v.so python tools/ab_sweep.py "" > $O/r2d_sweep_$v.txt 2>&1; cat $O/r2d_sweep_$v.txt
done
python tools/ab_sweep.py --input source --nfrag 4096 "" > $O/r2d_src_base.txt 2>&1; cat $O/r2d_src_base.txt
for v in pf1 pf2 el1; do
  SNAPPY_B200_LIB=/root/repo/snappy.jl_b200/libsnappy_b200_$v.so python tools/ab_sweep.py --input source --nfrag 4096 "" > $O/r2d_src_$v.txt 2>&1; cat $O/r2d_src_$v.txt
done
ls -la $O
