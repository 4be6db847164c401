#!/bin/bash
# round 2, GPU call 22 (8 GPUs): where the multi-GPU step goes (SNAPPY_B200_TRACE_MULTI phase events), config 2 only
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
SNAPPY_B200_TRACE_MULTI=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 3 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > $O/r2v_bench_n8.json 2> $O/r2v_bench_n8.err; echo "bench n8 rc=$?"
grep "snappy_b200 comm" $O/r2v_bench_n8.err | tail -32 > $O/r2v_trace_n8.txt; cat $O/r2v_trace_n8.txt | cut -c1-260
python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench_n8.json').read().strip().splitlines()[-1]); print('N8 value %.2f compress %.2f uncompress %.1f ms/step %.2f' % (d['value'], d['compress_gbps'], d['uncompress_gbps'], d['ms_per_step']))"
