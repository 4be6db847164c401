#!/bin/bash
# Runs bench.py N times, each under a watchdog that dumps the Python stack if a run exceeds 70 s
# (a hang inside a library call shows up as the ctypes frame it sits in).  Usage: tools/bench_loop.sh [N]
N=${1:-5}
mkdir -p gpurun_out
for i in $(seq 1 $N); do
  s=$(date +%s)
  timeout 100 python -X faulthandler -c "
import faulthandler, runpy, sys
faulthandler.dump_traceback_later(70, exit=True)
sys.argv = ['bench.py']
runpy.run_path('bench.py', run_name='__main__')
" > gpurun_out/loop_$i.json 2> gpurun_out/loop_$i.err
  rc=$?
  e=$(date +%s)
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/loop_$i.json').read().strip().splitlines()[-1])
    print('run $i rc=$rc', $e - $s, 's value', round(d['value'], 2), 'e2e', round(d['e2e']['value'], 2), 'kernel_ms', round(d['roofline']['kernel_ms'], 2))
except Exception as ex:
    print('run $i rc=$rc', $e - $s, 's NO JSON', ex)
    print(open('gpurun_out/loop_$i.err').read()[-1500:])
PY
done
