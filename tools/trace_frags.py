"""Where the time of one compress launch goes, fragment by fragment (library option `trace`):
fragments per table placement, their mean duration per fragment class, the number of busy warps over time, and the
length of the tail (time between the moment the queue runs dry = the last fragment STARTS and the end of the launch).

    python tools/trace_frags.py [--nfrag 16384] [--input mix|source] ["lpt=1,l2_chains=14" ...]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from snappy_jl_b200 import _abi, device, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nfrag", type=int, default=16384)
ap.add_argument("--input", default="mix")
ap.add_argument("settings", nargs="*", default=[""])
a = ap.parse_args()
cache = "/tmp/ab_%s_%d.npy" % (a.input, a.nfrag)
if os.path.exists(cache):
    raw = np.load(cache)
else:
    raw = synth.mix(a.nfrag, seed=2026) if a.input == "mix" else synth.source_like(a.nfrag * 65536, seed=2026)
    np.save(cache, raw)
cls = synth.fragment_classes(a.nfrag, 2026) if a.input == "mix" else np.zeros(a.nfrag, dtype=np.int64)
names = ["random", "dictionary", "text", "records"] if a.input == "mix" else ["source"]
d = torch.from_numpy(raw).cuda()
for setting in a.settings:
    opts = dict(kv.split("=") for kv in setting.split(",") if kv)
    saved = {k: _abi.lib().snappy_b200_get_option(k.encode()) for k in opts}
    for k, v in opts.items():
        device.set_option(k, int(v))
    for _ in range(2):
        device.compress_device(d)
    device.set_option("trace", 1)
    device.compress_device(d)
    device.set_option("trace", 0)
    buf = np.zeros(2 * a.nfrag, dtype=np.uint64)
    n = _abi.lib().snappy_b200_debug_trace(buf.ctypes.data_as(ctypes.c_void_p), a.nfrag)
    for k, v in saved.items():
        if v >= 0:
            device.set_option(k, v)
    t = buf.reshape(-1, 2)[:n]
    begin, end = (t[:, 0] & ~np.uint64(1)).astype(np.float64), (t[:, 1] & ~np.uint64(0xff)).astype(np.float64)
    smem = (t[:, 0] & np.uint64(1)).astype(bool)
    t0 = begin.min()
    begin, end = (begin - t0) / 1e6, (end - t0) / 1e6  # ms
    dur = end - begin
    total = end.max()
    dry = begin.max()
    print("== %s: launch %.2f ms, queue dry at %.2f ms (tail %.2f ms), kernel ms %.2f" % (
        setting or "(defaults)", total, dry, total - dry, device.last_kernel_ms(0)))
    for name, m in (("shared-table warps", smem), ("global-table warps", ~smem)):
        if not m.any():
            continue
        line = "   %-19s %6d fragments, busy-time sum %8.1f ms, last end %.2f ms |" % (name, int(m.sum()), dur[m].sum(),
                                                                                    end[m].max())
        for c, cn in enumerate(names):
            mm = m & (cls[:n] == c)
            if mm.any():
                line += " %s %.3f ms x%d" % (cn, dur[mm].mean(), int(mm.sum()))
        print(line)
    # busy warps over time (10 slices) and over the tail
    edges = np.linspace(0, total, 11)
    busy = [float(np.clip(np.minimum(end, edges[i + 1]) - np.maximum(begin, edges[i]), 0, None).sum() /
                  (edges[i + 1] - edges[i])) for i in range(10)]
    print("   busy warps per tenth of the launch: " + " ".join("%.0f" % b for b in busy))
    tail_busy = float(np.clip(end - np.maximum(begin, dry), 0, None).sum() / max(total - dry, 1e-9))
    print("   busy warps during the tail: %.0f" % tail_busy, flush=True)
