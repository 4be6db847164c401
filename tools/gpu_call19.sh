#!/bin/bash
# round 2, GPU call 19 (1 GPU): clean-cut re-tiling of foreign streams
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_foreign.py tests/test_gpu_configs.py -m gpu -q -x > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2s_pytest.log
tail -15 $O/r2s_pytest.log
timeout 600 python tools/time_foreign.py > $O/r2s_foreign.txt 2>&1; cat $O/r2s_foreign.txt
