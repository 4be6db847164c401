"""GPU: indexed decoder time on the 1 GiB mix (min of 4), per decode_occupancy."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snappy_jl_b200 import device, synth
raw = synth.mix(16384, seed=2026)
d = torch.from_numpy(raw).cuda()
stream, index = device.compress_device(d, want_index=True)
for occ in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "12").split(",")]:
    device.set_option("decode_occupancy", occ)
    ts = []
    for _ in range(4):
        back = device.uncompress_device(stream, index=index, claimed=raw.size)
        ts.append(device.last_kernel_ms(1))
    print("occ", occ, "decode ms min %.2f avg %.2f" % (min(ts), sum(ts) / len(ts)), torch.equal(back, d), flush=True)
