"""GPU tuning sweep of the chain-kernel knobs.  Usage:
    python tools/tune_compress.py NFRAG "smem,l2,reserve[,spec_smem,spec_l2[,l2_ctas[,ring_smem,ring_l2]]];..."
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from snappy_jl_b200 import device, synth
nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
raw = synth.mix(nfrag, seed=2026)
d = torch.from_numpy(raw).cuda()
ref = None
for kv in sys.argv[3:]:  # global library options, e.g. window=1
    k_, v_ = kv.split('=')
    device.set_option(k_, int(v_))
for cfg in (sys.argv[2] if len(sys.argv) > 2 else '6,0,2;6,14,2').split(';'):
    v = [int(x) for x in cfg.split(',')]
    sm, l2, rs = v[:3]
    ss, sl = (v[3], v[4]) if len(v) >= 5 else (32, 16)
    nc = v[5] if len(v) >= 6 else 1
    device.set_option("l2_ctas", nc)
    rs_, rl_ = (v[6], v[7]) if len(v) >= 8 else (2048, 1024)  # the library defaults
    device.set_option("ring_smem", rs_)
    device.set_option("ring_l2", rl_)
    device.set_option("l2_reserve", rs)
    device.set_option("smem_chains", sm)
    device.set_option("l2_chains", l2)
    device.set_option("spec_smem", ss)
    device.set_option("spec_l2", sl)
    best, tot, cnt = 1e9, 0.0, 0
    for it in range(6):  # like bench.py: compress and uncompress alternate; the first pass is warm-up
        stream, index = device.compress_device(d, want_index=True)
        ms = device.last_kernel_ms(0)
        back = device.uncompress_device(stream, index=index, claimed=raw.size)
        if it:
            best = min(best, ms)
            tot += ms
            cnt += 1
    avg = tot / cnt
    s = stream.clone()
    if ref is None:
        ref = s
    ok = bool(torch.equal(s, ref))
    print("smem=%d l2=%dx%d reserve=%d spec=%d/%d ring=%d/%d min_ms=%.2f avg_ms=%.2f GB/s=%.1f same=%s" % (sm, nc, l2, rs, ss, sl, rs_, rl_, best, avg, raw.size / avg / 1e6, ok), flush=True)
