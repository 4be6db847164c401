"""ONE GPU: what the multi-GPU data path costs besides NVLink and NCCL.  A loopback world of W ranks (all in this
process, peer pointers are local) compresses / uncompresses W streams of 1 GiB / W each -- the kernels, the byte-count
matrix, k_assemble, k_pull and the sharded decode are the production ones; the collectives disappear -- next to the
plain single-stream calls on the same 1 GiB.

  python tools/time_loopback.py [W ...]        (default 8)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from snappy_jl_b200 import device, multi, synth

FR = 65536
raw = synth.mix(16384, seed=2026)
d = torch.from_numpy(raw).cuda()


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[len(ts[2:]) // 2], r


tc, (stream, index) = timed(lambda: device.compress_device(d, want_index=True))
kc = device.last_kernel_ms(0)
tu, back = timed(lambda: device.uncompress_device(stream, index=index, claimed=raw.size))
ku = device.last_kernel_ms(1)
print("plain, one 1 GiB stream : compress call %6.2f ms (kernel %5.2f) | uncompress call %5.2f ms (kernel %5.2f)" % (tc, kc, tu, ku), flush=True)
for W in [int(x) for x in (sys.argv[1:] or ["8"])]:
    per = 16384 // W
    raws = [raw[s * per * FR:(s + 1) * per * FR] for s in range(W)]
    totals = [r.size for r in raws]
    comm = multi.LibComm(loopback_world=W)
    shards = []
    for r in range(W):
        for s in range(W):
            lo, hi = multi.shard_bounds(totals[s], W)[r]
            shards.append(d[s * per * FR + lo: s * per * FR + hi])
    tc2, (streams, indexes, lens) = timed(lambda: comm.compress(shards, totals))
    kc2 = device.last_kernel_ms(0)
    outs = [torch.empty_like(x) for x in shards]
    tu2, _ = timed(lambda: comm.uncompress(streams, indexes, totals, outs=outs))
    ku2 = device.last_kernel_ms(1)
    ok = all(torch.equal(a, b) for a, b in zip(outs, shards))
    print("loopback world %d, %d streams of %4d MiB: compress call %6.2f ms (kernel %5.2f) | uncompress call %5.2f ms (kernel %5.2f) | round trip %s | bytes %d vs %d" % (
        W, W, totals[0] >> 20, tc2, kc2, tu2, ku2, ok, sum(lens), stream.numel()), flush=True)
    comm.close()
