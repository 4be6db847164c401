#!/bin/bash
# round 2, GPU call 29 (1 GPU): the state after the policy hoist: GPU suite, bench (both arms), compress capture, launch list
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2zc_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2zc_pytest.log
tail -2 $O/r2zc_pytest.log
timeout 900 python bench.py > $O/r2zc_bench.json 2> $O/r2zc_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2zc_bench.json").read().strip().splitlines()[-1])
    print("C2 value %.2f compress %.2f uncompress %.1f kernel %.2f e2e %.2f (%.1f ms) pageable %.2f (%.1f ms) frac %.4f" % (d["value"], d["compress_gbps"], d["uncompress_gbps"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["pageable"]["value"], d["e2e"]["pageable"]["ms_per_step"], d["roofline"]["frac"]))
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.2f" % v["value"], {a: round(v[a], 1) for a in v if a.endswith("gbps")}, "ms/step %.2f" % v["ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference > $O/r2zc_bench_ref.json 2> $O/r2zc_bench_ref.err; echo "reference arm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_compress_window_mixed -s 1 -c 1 -f -o $O/r2zc_prof_compress python tools/prof_run.py 16384 0 > $O/r2zc_ncu.log 2>&1; tail -1 $O/r2zc_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2zc_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-e2e > $O/r2zc_ncu_bench.log 2>&1; echo "launch list rc=$?"
