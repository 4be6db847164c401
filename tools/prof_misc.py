"""Workload for the ncu captures of the kernels besides the compress kernel and the indexed decoder
(north_star: every kernel ships with a capture): batched pages (k_compress_pages_window / k_compress_pages /
k_decode_pages), cost estimate + order (schedule.cuh), scan + compaction, index-free parse (k_parse_*, k_build_index).
Every kernel is launched ONCE (after an untimed warm-up pass that ncu skips with -s), pages first.

    python tools/prof_misc.py [nfrag=512] [pages=16384]
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from snappy_jl_b200 import device, synth

nfrag = int(sys.argv[1]) if len(sys.argv) > 1 else 512
npages = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
pages = synth.pages(npages, 4096, seed=7)
dp = torch.from_numpy(pages.reshape(-1)).cuda()
in_off = torch.arange(npages, dtype=torch.int64, device="cuda") * 4096
in_sz = torch.full((npages,), 4096, dtype=torch.int32, device="cuda")
dec = torch.empty(npages * 4096, dtype=torch.uint8, device="cuda")
out, out_off, out_sz = device.compress_batched_device(dp, in_off, in_sz)
sizes, st = device.uncompress_batched_device(out, out_off, out_sz, dec, in_off, in_sz)
device.set_option("pages_window", 0)   # the serial page kernel (pages above 8 KiB)
device.compress_batched_device(dp, in_off, in_sz)
device.set_option("pages_window", 1)
torch.cuda.synchronize()
assert int(st.abs().sum()) == 0 and torch.equal(dec, dp)
raw = synth.mix(nfrag, seed=2026)
d = torch.from_numpy(raw).cuda()
stream, index = device.compress_device(d, want_index=True)
back = device.uncompress_device(stream.clone(), index=None, claimed=raw.size)  # no index: parse kernels
torch.cuda.synchronize()
assert torch.equal(back, d)
print("ok pages+parse", raw.size, stream.numel(), int(out_sz.sum()))
