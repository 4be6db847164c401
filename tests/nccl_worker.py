"""One rank of the NCCL parity check (tests/test_gpu_multi.py::test_nccl_world_assembles_the_oracles_bytes; launched by
torch.distributed.run, one process per GPU).  Every rank generates the same small streams from fixed seeds, feeds its
own whole-fragment runs to snappy_b200_comm_compress, and the OWNER of each stream compares the assembled bytes
(NCCL size all-gather + NVLink peer stores) and the side index with the CPU oracle; then the inverse."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pyoracle  # noqa: E402
from snappy_jl_b200 import multi, synth  # noqa: E402

FRAGMENT = 65536


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    raws = [synth.mix(37, seed=21, tail=4321), np.zeros(0, dtype=np.uint8), synth.source_like(11 * FRAGMENT + 5, seed=3),
            synth.mix(2, seed=9)[: FRAGMENT + 1], synth.mix(64, seed=2)]
    totals = [r.size for r in raws]
    comm = multi.LibComm()
    for rep in range(2):  # the second pass reuses the arenas
        shards = []
        for raw in raws:
            lo, hi = multi.shard_bounds(raw.size, world)[rank]
            shards.append(torch.from_numpy(raw[lo:hi].copy()).to(dev) if hi > lo else None)
        streams, indexes, lens = comm.compress(shards, totals)
        for s, raw in enumerate(raws):
            want = pyoracle.compress_np(raw)
            assert lens[s] == want.size, (rank, s, lens[s], want.size)
            if comm.owns(s):
                got = streams[s].cpu().numpy()
                assert np.array_equal(got, want), "rank %d: stream %d differs from the oracle" % (rank, s)
                idx = indexes[s].cpu().numpy()
                assert idx[-1] == want.size and np.all(np.diff(idx) >= 0)
            else:
                assert streams[s] is None
        outs = comm.uncompress(streams, indexes, totals)
        for s, raw in enumerate(raws):
            lo, hi = multi.shard_bounds(raw.size, world)[rank]
            assert np.array_equal(outs[s].cpu().numpy(), raw[lo:hi]), "rank %d: run of stream %d" % (rank, s)
        # foreign buffers, no side index: the owner parses, everybody pulls
        ext = [torch.from_numpy(pyoracle.compress_np(r)).to(dev) if comm.owns(s) else None for s, r in enumerate(raws)]
        outs = comm.uncompress(ext, [None] * len(raws), totals)
        for s, raw in enumerate(raws):
            lo, hi = multi.shard_bounds(raw.size, world)[rank]
            assert np.array_equal(outs[s].cpu().numpy(), raw[lo:hi])
    # a corrupt stream: every rank gets the reference's status
    bad = bytearray(pyoracle.compress(raws[0].tobytes()))
    for k in range(10):
        bad[len(bad) // 2 + 5 * k] ^= 0x3C
    want_status = pyoracle.status_of_uncompress(bytes(bad))
    ext = [None] * len(raws)
    if comm.owns(0):
        ext[0] = torch.from_numpy(np.frombuffer(bytes(bad), dtype=np.uint8).copy()).to(dev)
    for s in range(1, len(raws)):
        if comm.owns(s):
            ext[s] = torch.from_numpy(pyoracle.compress_np(raws[s])).to(dev)
    st = [None] * len(raws)
    try:
        comm.uncompress(ext, [None] * len(raws), totals, statuses_out=st)
        raised = False
    except Exception:
        raised = True
    assert raised and st[0] == want_status and all(x == 0 for x in st[1:]), (rank, st, want_status)
    comm.close()
    dist.barrier()
    print("NCCL_PARITY_OK rank %d of %d" % (rank, world), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
