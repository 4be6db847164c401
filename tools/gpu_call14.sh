#!/bin/bash
# round 2, GPU call 14 (1 GPU): full GPU suite, the driver's N = 1 bench command (all configs), pageable copy variants
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2n_pytest.log
tail -5 $O/r2n_pytest.log
/usr/bin/time -v timeout 900 python bench.py > $O/r2n_bench.json 2> $O/r2n_bench.err; echo "bench rc=$?"; grep -E "Elapsed|Maximum resident" $O/r2n_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2n_bench.json").read().strip().splitlines()[-1])
    print("C2 value %.2f compress %.2f uncompress %.1f e2e %.2f (%.1f ms) pageable %.2f (%.1f ms)" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["value"], d["e2e"]["pageable"]["ms_per_step"]))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], {a: round(v[a], 1) for a in v if a.endswith("gbps")}, "ms/step %.2f" % v["ms_per_step"], "e2e", (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
for env in "SNAPPY_B200_NO_STREAM_COPY=1 SNAPPY_B200_COPY_THREADS=4" "SNAPPY_B200_COPY_THREADS=4" "SNAPPY_B200_COPY_THREADS=6" "SNAPPY_B200_COPY_THREADS=8" "SNAPPY_B200_NO_STREAM_COPY=1 SNAPPY_B200_COPY_THREADS=8"; do
  env $env timeout 200 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$env: pageable e2e %.2f GB/s (%.1f ms), pinned %.2f (%.1f ms)' % (d['e2e']['pageable']['value'], d['e2e']['pageable']['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done 2>&1 | tee $O/r2n_pageable.txt
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread"
