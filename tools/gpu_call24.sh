#!/bin/bash
# round 2, GPU call 24 (8 GPUs): the driver's N = 8 command as it will be run (config 2 + nested config 5) after the rotation
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2x_bench_n8.json 2> $O/r2x_bench_n8.err; echo "bench n8 rc=$?"
tail -3 $O/r2x_bench_n8.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2x_bench_n8.json").read().strip().splitlines()[-1])
    print("N8 C2 value %.2f compress %.2f uncompress %.1f ms/step %.2f kernel %.2f e2e %s" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["ms_per_step"], d["roofline"]["kernel_ms"], (d.get("e2e") or {}).get("value")))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], "compress", v.get("compress_gbps"), "uncompress", v.get("uncompress_gbps"), "ms/step %.2f" % v["ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2x_bench_n2.json 2> $O/r2x_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2x_bench_n2.json").read().strip().splitlines()[-1])
    print("N2 C2 value %.2f compress %.2f uncompress %.1f ms/step %.2f" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["ms_per_step"]))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], "compress", v.get("compress_gbps"), "uncompress", v.get("uncompress_gbps"), "ms/step %.2f" % v["ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
