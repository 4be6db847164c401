#!/bin/bash
# round 2, GPU call 13 (8 GPUs): the driver's N = 8 command (config 2 + nested config 5), library communicator
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 3 --warmup 3 > $O/r2m_bench_n8.json 2> $O/r2m_bench_n8.err; echo "bench n8 rc=$?"
tail -5 $O/r2m_bench_n8.err | cut -c1-400
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2m_bench_n8.json").read().strip().splitlines()[-1])
    print("N8 C2 value %.2f compress %.2f uncompress %.1f ms/step %.2f kernel %.2f e2e %s" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["ms_per_step"], d["roofline"]["kernel_ms"], (d.get("e2e") or {}).get("value")))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], "compress", v.get("compress_gbps"), "uncompress", v.get("uncompress_gbps"), "ms/step %.2f" % v["ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
