#!/bin/bash
# round 2, GPU call 15 (1 GPU): the driver's N = 1 bench command (all configs), copy threads, launch list
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 900 python bench.py > $O/r2o_bench.json 2> $O/r2o_bench.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2o_bench.json").read().strip().splitlines()[-1])
    print("C2 value %.2f compress %.2f uncompress %.1f e2e %.2f (%.1f ms) pageable %.2f (%.1f ms)" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["value"], d["e2e"]["pageable"]["ms_per_step"]))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], {a: round(v[a], 1) for a in v if a.endswith("gbps")}, "ms/step %.2f" % v["ms_per_step"], "e2e", (v.get("e2e") or {}).get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
T0=$(date +%s)
timeout 600 python bench.py --impl reference > $O/r2o_bench_ref.json 2> $O/r2o_bench_ref.err; echo "reference arm rc=$? wall $(( $(date +%s) - T0 )) s"; cut -c1-400 $O/r2o_bench_ref.json
for env in "SNAPPY_B200_COPY_THREADS=10" "SNAPPY_B200_COPY_THREADS=12" "SNAPPY_B200_COPY_THREADS=14"; do
  env $env timeout 200 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$env: pageable e2e %.2f GB/s (%.1f ms), pinned %.2f (%.1f ms)' % (d['e2e']['pageable']['value'], d['e2e']['pageable']['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done 2>&1 | tee $O/r2o_pageable.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2o_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > $O/r2o_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"; wc -l $O/r2o_launches_bench.csv
