#!/bin/bash
# A/B of library options through bench.py (device-resident legs only):  bash tools/ab_bench.sh "overlap_compact=1" "overlap_compact=0"
for o in "$@"; do
  SNAPPY_B200_OPTIONS=$o python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null > /tmp/ab.json
  python - "$o" <<'PY'
import json, sys
d = json.load(open("/tmp/ab.json"))
print("%-28s value %.2f GB/s  compress %.2f GB/s (kernel %.2f ms)  uncompress %.1f GB/s" % (
    sys.argv[1], d["value"], d["compress_gbps"], d["roofline"]["kernel_ms"], d["uncompress_gbps"]))
PY
done
