"""Small fixed workload for ncu captures: compress + uncompress `nfrag` fragments of the mix once."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snappy_jl_b200 import device, synth
nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = synth.mix(nfrag, seed=2026)
d = torch.from_numpy(raw).cuda()
device.set_option("compress_variant", variant)
for kv in sys.argv[3:]:  # extra library options, e.g. smem_chains=0 l2_ctas=3
    k, v = kv.split("=")
    device.set_option(k, int(v))
for _ in range(2):
    stream, index = device.compress_device(d, want_index=True)
    back = device.uncompress_device(stream, index=index, claimed=raw.size)
torch.cuda.synchronize()
assert torch.equal(back, d)
print("ok", raw.size, stream.numel(), "compress_ms", device.last_kernel_ms(0), "uncompress_ms", device.last_kernel_ms(1))
