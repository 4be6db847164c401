"""Small workload for ncu: index-free uncompress (config 3) of `nfrag` fragments of the mix."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snappy_jl_b200 import device, synth
nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
raw = synth.mix(nfrag, seed=2026)
d = torch.from_numpy(raw).cuda()
stream, index = device.compress_device(d, want_index=False)
for _ in range(2):
    back = device.uncompress_device(stream, claimed=raw.size)
torch.cuda.synchronize()
assert torch.equal(back, d)
print("ok", device.last_kernel_ms(1))
