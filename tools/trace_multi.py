"""2+ GPUs: where the time of multi.compress_streams goes (run under torchrun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from snappy_jl_b200 import device, multi, synth
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = 16384 // world
shards = [torch.from_numpy(synth.mix(per, 2026 + 1000 * s + rank)).to(dev) for s in range(world)]
totals = [per * world * 65536] * world
codec = multi.CudaCodec()
def sync():
    torch.cuda.synchronize()
    return time.perf_counter()
for it in range(4):
    dist.barrier(); t0 = sync()
    pairs = codec.compress_shards(shards, totals); t1 = sync()
    stream, index = multi.compress_streams(shards, totals, codec); t2 = sync()
    back = multi.uncompress_streams(stream, index, totals[rank], codec); t3 = sync()
    if rank == 0:
        print("codec.compress_shards %.1f ms (kernel %.1f) | compress_streams %.1f ms | uncompress_streams %.1f ms" % (
            (t1 - t0) * 1e3, device.last_kernel_ms(0), (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
dist.destroy_process_group()
