"""GPU: decode throughput of streams this library did not write, 256 MiB of the mix.

  block = 32768 : every 32 KiB compressed on its own (what a 32 KiB-block encoder emits, like tests/data/alice29.snappy):
                  elements never cross the 64 KiB tiles -> parse + parallel tiles (path 1)
  block = 50000 : blocks that do not tile 64 KiB: literals straddle tile boundaries, copies reach into the previous tile
                  -> parse + parallel tiles + bounded serial walk over the rejected tiles (path 2)
  corrupt       : one garbled spot in the middle -> the reference's status from the parse alone
The streams are built by the CPU oracle (test infrastructure); the timed call is device-resident uncompress without an index."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import pyoracle as oracle
import snappy_jl_b200 as Snappy
from snappy_jl_b200 import device, synth

tile = synth.mix(512, seed=23)  # 32 MiB
reps = 8
raw = np.tile(tile, reps)
d_raw = torch.from_numpy(raw).cuda()


def blocky(block):
    parts = [np.frombuffer(oracle.encode32(raw.size), dtype=np.uint8)]
    one = []
    for o in range(0, tile.size, block):
        s = np.frombuffer(oracle.compress(tile[o: o + block].tobytes()), dtype=np.uint8)
        _, k = oracle.parse32(s.tobytes()[:8], 0)
        one.append(s[k:])
    return one, parts


def path():
    return Snappy._abi.lib().snappy_b200_get_option(b"last_decode_path")


for block in (32768, 65536, 50000, 7777):
    if (tile.size % block) != 0:  # blocks run across the repetitions of the tile: compress the whole buffer block by block
        pieces = []
        for o in range(0, raw.size, block):
            s = oracle.compress(raw[o: o + block].tobytes())
            _, k = oracle.parse32(s[:8], 0)
            pieces.append(np.frombuffer(s[k:], dtype=np.uint8))
        stream = np.concatenate([np.frombuffer(oracle.encode32(raw.size), dtype=np.uint8)] + pieces)
    else:
        one, hdr = blocky(block)
        stream = np.concatenate(hdr + one * reps)
    d = torch.from_numpy(stream).cuda()
    out = torch.empty(raw.size, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        device.uncompress_device(d, out=out, claimed=raw.size)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ok = torch.equal(out, d_raw)
    print("block %6d: %8.1f MB stream, path %d, %6.2f ms (min of 4) = %6.1f GB/s of output, identical %s" % (
        block, stream.size / 1e6, path(), min(ts) * 1e3, raw.size / min(ts) / 1e9, ok), flush=True)
