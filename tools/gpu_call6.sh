#!/bin/bash
# round 2, GPU call 6 (1 GPU): parse-computed status (fast rejection), pageable host buffers through the bounce slots
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2f_pytest.log
tail -12 $O/r2f_pytest.log
python bench.py --steps 5 --warmup 3 --no-extra > $O/r2f_bench.json 2> $O/r2f_bench.err; echo "bench rc=$?"; tail -3 $O/r2f_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
    print("C2 value %.2f compress %.2f uncompress %.1f e2e %.2f (%.1f ms) pageable %s" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pageable")))
except Exception as e:
    print("bench parse failed", e)
PY
for t in 2 4 8 12; do SNAPPY_B200_COPY_THREADS=$t python bench.py --steps 3 --warmup 2 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('copy threads $t: pageable e2e %.2f GB/s (%.1f ms), pinned %.2f' % (d['e2e']['pageable']['value'], d['e2e']['pageable']['ms_per_step'], d['e2e']['value']))"; done
ls -la $O
