#!/bin/bash
# round 2, GPU call 20 (1 GPU): index-free parse without the second full walk (config 3), launch list of config 3
cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_foreign.py tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q -x > $O/r2t_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2t_pytest.log
tail -4 $O/r2t_pytest.log
timeout 600 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2t_bench_c3.json 2> $O/r2t_bench_c3.err; echo "bench c3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2t_bench_c3.json').read().strip().splitlines()[-1]); print('c3 value %.2f GB/s, %.2f ms per step, decode kernel %.2f ms' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2t_launches_c3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline > $O/r2t_ncu_c3.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2t_launches_c3.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
# the last step's kernels: everything after the last k_parse_guess
names=[(r[ki].split("(")[0], float(r[vi].replace(",",""))) for r in rows[1:]]
last=max(i for i,(k,v) in enumerate(names) if "k_parse_guess" in k)
agg={}
for k,v in names[last:]:
    a=agg.setdefault(k,[0,0]); a[0]+=1; a[1]+=v
for k,v in agg.items(): print("%-50s x%-3d %8.3f ms"%(k[:50],v[0],v[1]/1e6))
PY
