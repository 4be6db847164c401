"""GPU: the multi-GPU data path of the library (snappy_b200_comm_*, csrc/multi_host.inc + multi.cuh).

On ONE GPU the exact multi-rank path runs as a loopback world (all ranks in this process; the peer pointers are local
allocations, the collectives disappear): whole-fragment sharding, the byte-count matrix, k_assemble's stores into the
owners' arenas, k_pull and the sharded decode are all the production kernels.  The assembled streams must be
byte-identical to the oracle's compress of the whole stream (src/Snappy.jl:20-36) for every world size.
With >= 2 GPUs the same checks run over NCCL + cudaIpc peer mappings (tests/nccl_worker.py under torchrun)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, read_data

pytestmark = pytest.mark.gpu
FRAGMENT = 65536


def _streams():
    from snappy_jl_b200 import synth
    rng = np.random.default_rng(5)
    return [
        synth.mix(40, seed=3, tail=777),                                   # 40 fragments + a ragged one
        np.zeros(0, dtype=np.uint8),                                       # empty stream: header only
        np.frombuffer(b"a", dtype=np.uint8).copy(),                        # one byte
        synth.mix(1, seed=8)[:FRAGMENT],                                   # exactly one fragment
        np.frombuffer(read_data("urls.10K") + read_data("html_x_4"), dtype=np.uint8).copy(),
        rng.integers(0, 256, 3 * FRAGMENT + 1, dtype=np.uint8),            # incompressible, ragged by one byte
        synth.source_like(9 * FRAGMENT + 12345, seed=4),
    ]


def _shards(raws, world, ranks):
    import torch
    from snappy_jl_b200 import multi
    out = []
    for r in ranks:
        for raw in raws:
            lo, hi = multi.shard_bounds(raw.size, world)[r]
            out.append(torch.from_numpy(raw[lo:hi].copy()).cuda() if hi > lo else None)
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_loopback_world_assembles_the_oracles_bytes(oracle, world):
    import torch
    from snappy_jl_b200 import multi
    raws = _streams()
    totals = [r.size for r in raws]
    comm = multi.LibComm(loopback_world=world)
    try:
        streams, indexes, lens = comm.compress(_shards(raws, world, range(world)), totals)
        for s, raw in enumerate(raws):
            want = oracle.compress_np(raw)
            assert lens[s] == want.size, (s, lens[s], want.size)
            got = streams[s].cpu().numpy()
            assert np.array_equal(got, want), "stream %d differs from the oracle at world %d" % (s, world)
            idx = indexes[s].cpu().numpy()
            nfrag = (raw.size + FRAGMENT - 1) // FRAGMENT
            assert idx.size == nfrag + 1 and idx[-1] == want.size and np.all(np.diff(idx) >= 0)
            _, sizes = oracle.compress_fragments(raw, raw.size, 0, nfrag) if nfrag else (None, np.zeros(0))
            hdr = len(oracle.encode32(raw.size))
            assert np.array_equal(np.diff(idx), sizes.astype(np.int64)) and idx[0] == hdr
        # inverse, with the side index (zero-copy: the streams are already in the owners' arenas)
        outs = comm.uncompress(streams, indexes, totals)
        for s, raw in enumerate(raws):
            back = np.concatenate([outs[r * len(raws) + s].cpu().numpy() for r in range(world)])
            assert np.array_equal(back, raw), "round trip of stream %d at world %d" % (s, world)
        # inverse from foreign buffers (the oracle's streams, copied into the arena) WITHOUT an index: owners parse
        ext = [torch.from_numpy(oracle.compress_np(r)).cuda() for r in raws]
        outs = comm.uncompress(ext, [None] * len(raws), totals)
        for s, raw in enumerate(raws):
            back = np.concatenate([outs[r * len(raws) + s].cpu().numpy() for r in range(world)])
            assert np.array_equal(back, raw)
    finally:
        comm.close()


def test_loopback_status_of_corrupt_and_unshardable_streams(snappy, oracle):
    """a failing stream gets the reference's status (the owner reruns the whole-stream path), others stay OK"""
    import torch
    from snappy_jl_b200 import multi, synth
    good = synth.mix(6, seed=12, tail=99)
    bad = bytearray(oracle.compress(read_data("alice29.txt")))
    pos = len(bad) // 2
    for k in range(12):  # garble a dozen bytes in the middle
        bad[pos + 7 * k] ^= 0x5A
    bad = bytes(bad)
    want_bad = oracle.status_of_uncompress(bad)
    assert want_bad != 0
    foreign = read_data("alice29.snappy")  # 32 KiB-block encoder: elements straddle 64 KiB boundaries
    streams = [torch.from_numpy(oracle.compress_np(good)).cuda(),
               torch.from_numpy(np.frombuffer(bad, dtype=np.uint8).copy()).cuda(),
               torch.from_numpy(np.frombuffer(foreign, dtype=np.uint8).copy()).cuda()]
    totals = [good.size, len(read_data("alice29.txt")), len(read_data("alice29.txt"))]
    comm = multi.LibComm(loopback_world=3)
    try:
        st = [None] * 3
        with pytest.raises(snappy.SnappyError) as e:
            comm.uncompress(streams, [None] * 3, totals, statuses_out=st)
        assert st[0] == 0 and st[1] == want_bad and st[2] == snappy._abi.BAD_ARGUMENT
        assert e.value.status == want_bad
        # a wrong length claim is rejected, not trusted
        with pytest.raises(snappy.SnappyError):
            comm.uncompress(streams[:1], [None], [good.size + 1])
        # and the communicator is still usable
        outs = comm.uncompress(streams[:1], [None], [good.size])
        assert np.array_equal(np.concatenate([o.cpu().numpy() for o in outs]), good)
    finally:
        comm.close()


def test_comm_rejects_wrong_shard_sizes(snappy):
    import torch
    from snappy_jl_b200 import multi, synth
    raw = synth.mix(5, seed=1)
    comm = multi.LibComm(loopback_world=2)
    try:
        wrong = [torch.from_numpy(raw[: 2 * FRAGMENT].copy()).cuda(), torch.from_numpy(raw[2 * FRAGMENT:].copy()).cuda()]
        with pytest.raises(snappy.SnappyError) as e:  # whole-fragment sharding gives rank 0 three fragments
            comm.compress(wrong, [raw.size])
        assert e.value.status == snappy._abi.BAD_ARGUMENT
    finally:
        comm.close()


def test_nccl_world_assembles_the_oracles_bytes():
    """>= 2 GPUs: the same parity over NCCL (size all-gather) and cudaIpc peer mappings (NVLink stores / loads)"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29741",
                        os.path.join(ROOT, "tests", "nccl_worker.py")], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert p.stdout.count("NCCL_PARITY_OK") == world, p.stdout[-2000:]
