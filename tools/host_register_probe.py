"""How fast is cudaHostRegister on an ordinary numpy array?  (decides between pinning a pageable caller buffer in place
and bouncing it through pinned slots)"""
import ctypes
import time

import numpy as np
import torch

torch.cuda.init()
rt = torch.cuda.cudart()
for mib in (64, 256, 1024):
    a = np.ones(mib << 20, dtype=np.uint8)
    for rep in range(2):
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        t1 = time.perf_counter()
        rc2 = rt.cudaHostUnregister(a.ctypes.data)
        t2 = time.perf_counter()
        print("%5d MiB: register %.1f ms (%.1f GB/s, rc %s), unregister %.1f ms" % (mib, (t1 - t0) * 1e3, a.nbytes / (t1 - t0) / 1e9, rc, (t2 - t1) * 1e3), flush=True)
