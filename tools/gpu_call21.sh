#!/bin/bash
# round 2, GPU call 21 (1 GPU): full GPU suite after the scan / parse changes; parse look-back distance (config 3)
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2u_pytest.log
tail -3 $O/r2u_pytest.log
for lib in "" _lb512 _lb256; do
  SNAPPY_B200_LIB=$PWD/snappy.jl_b200/libsnappy_b200$lib.so timeout 300 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('libsnappy_b200$lib.so: c3 value %.2f GB/s, %.2f ms per step' % (d['value'], d['ms_per_step']))"
done 2>&1 | tee $O/r2u_lookback.txt
