"""GPU: BASELINE config 4 -- batched independent 4 KiB pages, one stream per page."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from snappy_jl_b200 import device, synth
count = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
page = 4096
pages = synth.pages(count, page, seed=2026)
d = torch.from_numpy(pages.reshape(-1)).cuda()
offs = torch.arange(count, dtype=torch.int64, device="cuda") * page
sizes = torch.full((count,), page, dtype=torch.int32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, oo, osz = device.compress_batched_device(d, offs, sizes)
    torch.cuda.synchronize(); tc = time.perf_counter() - t0
    kc = device.last_kernel_ms(0)
    back = torch.empty_like(d)
    caps = sizes
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gs, st = device.uncompress_batched_device(out, oo, osz, back, offs, caps)
    torch.cuda.synchronize(); tu = time.perf_counter() - t0
    ku = device.last_kernel_ms(1)
    ok = bool(torch.equal(back, d)) and int(st.abs().sum()) == 0
    print("pages=%d compress %.2f ms (kernel %.2f, %.1f GB/s) uncompress %.2f ms (kernel %.2f, %.1f GB/s) ratio %.3f ok=%s" % (
        count, tc * 1e3, kc, d.numel() / kc / 1e6, tu * 1e3, ku, d.numel() / ku / 1e6, float(osz.sum()) / d.numel(), ok), flush=True)
