"""A/B one library option on the 1 GiB mix: bit-exactness against the default and kernel time.
    python tools/ab_option.py slowcont=1 [nfrag]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snappy_jl_b200 import device, synth
name, val = sys.argv[1].split("=")
nfrag = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
d = torch.from_numpy(synth.mix(nfrag, seed=2026)).cuda()


def run():
    best = 1e9
    for it in range(4):
        s, idx = device.compress_device(d, want_index=True)
        if it:
            best = min(best, device.last_kernel_ms(0))
        device.uncompress_device(s, index=idx, claimed=d.numel())
    return s.clone(), best


a, ta = run()
device.set_option(name, int(val))
b, tb = run()
device.set_option(name, 0)
print("default %.2f ms, %s=%s %.2f ms, identical=%s" % (ta, name, val, tb, bool(torch.equal(a, b))), flush=True)
