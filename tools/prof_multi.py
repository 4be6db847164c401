"""Small fixed workload for ncu captures of the multi-GPU kernels on ONE GPU: a loopback world of 4 ranks compresses
and uncompresses 4 streams (k_scan_shards, k_assemble, k_check_header, k_pull and the sharded decode), then a foreign
stream that needs the bounded serial walk (k_decode_serial_tiles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
from snappy_jl_b200 import device, multi, synth
W, per, FR = 4, 256, 65536
raw = synth.mix(W * per, seed=2026)
d = torch.from_numpy(raw).cuda()
totals = [per * FR] * W
comm = multi.LibComm(loopback_world=W)
shards = []
for r in range(W):
    for s in range(W):
        lo, hi = multi.shard_bounds(totals[s], W)[r]
        shards.append(d[s * per * FR + lo: s * per * FR + hi])
for _ in range(2):
    streams, indexes, lens = comm.compress(shards, totals)
    outs = comm.uncompress(streams, indexes, totals)
torch.cuda.synchronize()
assert all(torch.equal(a, b) for a, b in zip(outs, shards))
comm.close()
import pyoracle as oracle
small = synth.mix(48, seed=17, tail=4321)
parts = [oracle.encode32(small.size)]
for o in range(0, small.size, 50000):
    s = oracle.compress(small[o: o + 50000].tobytes())
    _, k = oracle.parse32(s, 0)
    parts.append(s[k:])
st = torch.from_numpy(np.frombuffer(b"".join(parts), dtype=np.uint8).copy()).cuda()
for _ in range(2):
    got = device.uncompress_device(st)
torch.cuda.synchronize()
assert np.array_equal(got.cpu().numpy(), small)
print("ok")
