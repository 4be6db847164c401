#!/bin/bash
# round 2, GPU call 5 (2 GPUs): the library communicator over NCCL + cudaIpc, bench at N = 2
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_foreign.py -m gpu -q -x > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2e_pytest.log
tail -15 $O/r2e_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2e_bench_n2.json 2> $O/r2e_bench_n2.err; echo "bench n2 rc=$?"
tail -5 $O/r2e_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2e_bench_n2.json").read().strip().splitlines()[-1])
    print("N2 C2 value %.2f compress %.2f uncompress %.1f ms/step %.2f kernel %.2f e2e %s note %s" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["ms_per_step"], d["roofline"]["kernel_ms"], (d.get("e2e") or {}).get("value"), d.get("comm_note")))
    print(d["config"]["sharding"])
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], "compress", v.get("compress_gbps"), "uncompress", v.get("uncompress_gbps"), "ms/step %.2f" % v["ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
SNAPPY_B200_TRACE_MULTI=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 2 --no-extra --no-e2e 2>&1 | tail -3 | cut -c1-600
ls -la $O
