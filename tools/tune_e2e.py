"""GPU: host-buffer API (e2e) timing for different library options, e.g.
    python tools/tune_e2e.py host_pipeline=1 host_pipeline=2,pipe_chunk_frags=4096"""
import sys, os, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import snappy_jl_b200 as Snappy
from snappy_jl_b200 import device, synth
nfrag = 16384
raw = synth.mix(nfrag, seed=2026)
n = raw.size
host_in = torch.from_numpy(raw).pin_memory()
cap = Snappy.maxlength_compressed(n)
h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
h_back = torch.empty(n, dtype=torch.uint8).pin_memory()
lib = Snappy._abi.lib()
lib.snappy_b200_init(0)
for cfg in (sys.argv[1:] or ["host_pipeline=1"]):  # each argument: comma-separated name=value options
    for kv in cfg.split(","):
        k_, v_ = kv.split("=")
        device.set_option(k_, int(v_))
    chunk = cfg
    for rep in range(3):
        ol = ctypes.c_size_t(cap)
        t0 = time.perf_counter()
        rc = lib.snappy_b200_compress(host_in.data_ptr(), n, h_out.data_ptr(), ctypes.byref(ol))
        t1 = time.perf_counter()
        bl = ctypes.c_size_t(n)
        rc2 = lib.snappy_b200_uncompress(h_out.data_ptr(), ol.value, h_back.data_ptr(), ctypes.byref(bl))
        t2 = time.perf_counter()
    ok = rc == 0 and rc2 == 0 and bool(torch.equal(h_back, host_in))
    print("%s: compress %.1f ms uncompress %.1f ms total %.1f ms -> %.1f GB/s ok=%s" % (
        chunk, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, n / (t2 - t0) / 1e9, ok), flush=True)
