#!/bin/bash
# round 2, GPU call 23 (8 GPUs): k_assemble with a per-rank rotated stream order (no NVLink incast), wider k_pull
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2w_pytest.log
SNAPPY_B200_TRACE_MULTI=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 3 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > $O/r2w_bench_n8.json 2> $O/r2w_bench_n8.err; echo "bench n8 rc=$?"
grep "snappy_b200 comm" $O/r2w_bench_n8.err | grep "total" | tail -16 > $O/r2w_trace_n8.txt; cat $O/r2w_trace_n8.txt | cut -c1-230
python -c "
import json; d=json.loads(open('gpurun_out/r2w_bench_n8.json').read().strip().splitlines()[-1]); print('N8 value %.2f compress %.2f uncompress %.1f ms/step %.2f' % (d['value'], d['compress_gbps'], d['uncompress_gbps'], d['ms_per_step']))"
