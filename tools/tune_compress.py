"""GPU tuning sweep: chain-kernel occupancy knobs (smem-table warps / global-table warps per SM)."""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snappy_jl_b200 import device, synth
nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
raw = synth.mix(nfrag, seed=2026)
d = torch.from_numpy(raw).cuda()
ref = None
for sm, l2, pf in [(int(a), int(b), int(c)) for a, b, c in (x.split(',') for x in (sys.argv[2] if len(sys.argv) > 2 else '6,0,8;6,6,8').split(';'))]:
    device.set_option("l2_reserve", pf)
    device.set_option("smem_chains", sm)
    device.set_option("l2_chains", l2)
    best = 1e9
    for _ in range(3):
        stream, index = device.compress_device(d, want_index=False)
        best = min(best, device.last_kernel_ms(0))
    s = stream.clone()
    if ref is None:
        ref = s
    ok = bool(torch.equal(s, ref))
    print("smem_chains=%d l2_chains=%d reserve=%d kernel_ms=%.2f GB/s=%.1f same=%s" % (sm, l2, pf, best, raw.size / best / 1e6, ok), flush=True)
