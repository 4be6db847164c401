"""Sweep library options (and, through SNAPPY_B200_LIB, builds) on one input in ONE process per build:
compress kernel time, whole-call time, uncompress time, and byte identity of every setting with the first one.

    python tools/ab_sweep.py [--nfrag 16384] [--input mix|source] "lpt=0" "lpt=1" "lpt=1,l2_reserve=2" ...

The input is cached in /tmp (generating 1 GiB takes ~30 s), so a shell loop over several builds pays for it once.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from snappy_jl_b200 import device, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nfrag", type=int, default=16384)
ap.add_argument("--input", default="mix")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("settings", nargs="*", default=[""])
a = ap.parse_args()

cache = "/tmp/ab_%s_%d.npy" % (a.input, a.nfrag)
if os.path.exists(cache):
    raw = np.load(cache)
else:
    raw = synth.mix(a.nfrag, seed=2026) if a.input == "mix" else synth.source_like(a.nfrag * 65536, seed=2026)
    np.save(cache, raw)
d = torch.from_numpy(raw).cuda()
n = d.numel()
out = torch.empty(32 + n + n // 6, dtype=torch.uint8, device="cuda")
back = torch.empty(n, dtype=torch.uint8, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
first = None
lib = os.environ.get("SNAPPY_B200_LIB", "default build")
print("# %s, %s x %d fragments" % (os.path.basename(lib), a.input, a.nfrag), flush=True)
touched = {}
for setting in a.settings:
    opts = dict(kv.split("=") for kv in setting.split(",") if kv)
    for k, v in opts.items():
        if k not in touched:
            touched[k] = device._abi.lib().snappy_b200_get_option(k.encode())
        device.set_option(k, int(v))
    kc, tc, tu, ku = [], [], [], []
    for it in range(a.reps + 2):
        ev[0].record()
        s, idx = device.compress_device(d, out=out, want_index=True)
        ev[1].record()
        device.uncompress_device(s, out=back, index=idx, claimed=n)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            kc.append(device.last_kernel_ms(0))
            ku.append(device.last_kernel_ms(1))
            tc.append(ev[0].elapsed_time(ev[1]))
            tu.append(ev[1].elapsed_time(ev[2]))
    same = "first"
    if first is None:
        first = s.clone()
        assert torch.equal(back, d)
    else:
        same = "identical" if (s.numel() == first.numel() and torch.equal(s, first)) else "DIFFERENT"
    print("%-40s compress kernel %6.2f ms (min %6.2f) call %6.2f ms = %6.1f GB/s | uncompress kernel %5.2f call %5.2f ms | %s"
          % (setting or "(defaults)", float(np.median(kc)), min(kc), float(np.median(tc)), n / float(np.median(tc)) / 1e6,
             float(np.median(ku)), float(np.median(tu)), same), flush=True)
    for k, v in touched.items():
        if v >= 0:
            device.set_option(k, v)
