#!/bin/bash
# round 2, GPU call 26 (1 GPU): validation of the final state: smoke, GPU suite, the driver's bench command, both arms
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2z_pytest.log
tail -3 $O/r2z_pytest.log
T0=$(date +%s)
timeout 900 python bench.py > $O/r2z_bench.json 2> $O/r2z_bench.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
    print("C2 value %.2f compress %.2f uncompress %.1f kernel %.2f e2e %.2f (%.1f ms) pageable %.2f (%.1f ms) frac %.4f issue %.3f" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["value"], d["e2e"]["pageable"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["issue"]["frac"]))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], {a: round(v[a], 1) for a in v if a.endswith("gbps")}, "ms/step %.2f" % v["ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference > $O/r2z_bench_ref.json 2> $O/r2z_bench_ref.err; echo "reference arm rc=$?"; cut -c1-330 $O/r2z_bench_ref.json
