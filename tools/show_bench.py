"""One-line summary of a bench.py JSON line:  python tools/show_bench.py gpurun_out/bench.json"""
import json, sys
d = json.load(open(sys.argv[1]))
r, e, c = d["roofline"], d.get("e2e") or {}, d.get("cpu_baseline") or {}
print("value %.2f GB/s | compress %.2f | uncompress %.1f | kernel %.2f ms (%s, frac %.4f, traffic %s) | e2e %s GB/s (%s ms) | cpu %s GB/s x%s | launches %s | clocks %s" % (
    d["value"], d["compress_gbps"], d["uncompress_gbps"], r["kernel_ms"], r["kernel"], r["frac"], r.get("traffic"),
    round(e.get("value", 0), 2), round(e.get("ms_per_step", 0), 1), round(c.get("value", 0), 2), c.get("cores"),
    d.get("gpu_launches"), d.get("clocks")))
