import sys, os
sys.path.insert(0, "/root/repo")
import torch
from snappy_jl_b200 import device, synth
raw = synth.mix(16384, seed=2026); d = torch.from_numpy(raw).cuda()
for skip in (0, 1):
    for cfg in ((6, 14), (6, 0), (0, 14)):
        device.set_option("dbg_skip_emit", skip); device.set_option("smem_chains", cfg[0]); device.set_option("l2_chains", cfg[1])
        ts = []
        for it in range(4):
            stream, index = device.compress_device(d, want_index=True); ts.append(device.last_kernel_ms(0))
            if not skip: back = device.uncompress_device(stream, index=index, claimed=raw.size)
        print("skip_emit", skip, cfg, "ms min %.2f" % min(ts[1:]), flush=True)
