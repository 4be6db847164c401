"""GPU: BASELINE config 5's corpus on one GPU -- 1 GiB of the source-code-like token stream
(every fragment is text-like, the expensive class), device-resident compress + uncompress."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snappy_jl_b200 import device, synth
n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
raw = synth.source_like(n, seed=2026)
d = torch.from_numpy(raw).cuda()
kc, ku = [], []
for it in range(4):
    stream, index = device.compress_device(d, want_index=True)
    kc.append(device.last_kernel_ms(0))
    back = device.uncompress_device(stream, index=index, claimed=n)
    ku.append(device.last_kernel_ms(1))
assert torch.equal(back, d)
print("source-like %d MiB: ratio %.3f compress kernel %.2f ms (%.1f GB/s) uncompress kernel %.2f ms (%.1f GB/s)" % (
    n >> 20, stream.numel() / n, min(kc[1:]), n / min(kc[1:]) / 1e6, min(ku[1:]), n / min(ku[1:]) / 1e6))
