#!/bin/bash
# round 2, GPU call 32 (8 GPUs): the driver's N = 8 and N = 4 commands on the final state
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for N in 8 4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2967$N bench.py --gpus $N --steps 5 --warmup 3 > $O/r2zf_bench_n$N.json 2> $O/r2zf_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2zf_bench_n$N.json").read().strip().splitlines()[-1])
    print("N$N C2 value %.2f compress %.2f uncompress %.1f ms/step %.2f kernel %.2f e2e %s" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["ms_per_step"], d["roofline"]["kernel_ms"], (d.get("e2e") or {}).get("value")))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], "compress", v.get("compress_gbps"), "uncompress", v.get("uncompress_gbps"), "ms/step %.2f" % v["ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
done
