"""Small end-to-end workload that touches every kernel of the default paths (meant for
`compute-sanitizer --tool memcheck python tools/sanitize_run.py`; the tool is closed on this GPU pool, so
here it only serves as a quick all-kernels smoke run).
Exercises every kernel of the default paths on a few fragments: device compress (both table
placements), indexed decode, index-free parse + decode, batched pages, shard API, host-buffer API."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import snappy_jl_b200 as Snappy
from snappy_jl_b200 import device, synth
raw = synth.mix(12, seed=3, tail=777)
d = torch.from_numpy(raw).cuda()
for opts in ({}, {"l2_chains": 0}, {"smem_chains": 0}, {"window": 0}):
    for k, v in opts.items():
        device.set_option(k, v)
    stream, index = device.compress_device(d, want_index=True)
    back = device.uncompress_device(stream, index=index, claimed=raw.size)
    assert torch.equal(back, d)
    for k, v in {"l2_chains": 14, "smem_chains": 6, "window": 1}.items():
        device.set_option(k, v)
back = device.uncompress_device(stream, claimed=raw.size)       # index-free parse
assert torch.equal(back, d)
pages = synth.pages(64, 4096, seed=5)
dp = torch.from_numpy(pages.reshape(-1)).cuda()
offs = torch.arange(64, dtype=torch.int64, device="cuda") * 4096
sizes = torch.full((64,), 4096, dtype=torch.int32, device="cuda")
out, oo, osz = device.compress_batched_device(dp, offs, sizes)
bk = torch.empty_like(dp)
device.uncompress_batched_device(out, oo, osz, bk, offs, sizes)
assert torch.equal(bk, dp)
got = Snappy.compress_np(raw)
assert np.array_equal(Snappy.uncompress_np(got), raw)
torch.cuda.synchronize()
print("sanitize workload ok")
