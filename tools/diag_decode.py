"""GPU diagnostic: timing + launch counts of the uncompress paths (indexed / no index / exact)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import snappy_jl_b200 as S
from snappy_jl_b200 import device, synth

nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
raw = synth.mix(nfrag, seed=2026)
d = torch.from_numpy(raw).cuda()
stream, index = device.compress_device(d, want_index=True)
print("compressed", raw.size, "->", stream.numel())
for occ in (8, 10, 12):
  device.set_option("decode_occupancy", occ)
  print("occupancy", occ)
  for name, kw in (("indexed", dict(index=index)),):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        back = device.uncompress_device(stream, claimed=raw.size, **kw)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(' ',name, rep, "%.2f ms" % (dt * 1e3), "launches", device.last_launch_count(1),
              "kernel_ms %.2f" % device.last_kernel_ms(1), "ok", bool(torch.equal(back, d)))
