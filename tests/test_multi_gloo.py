"""CPU suite, part 3: the multi-GPU host logic (fragment-boundary sharding, size all-gather,
segment assembly, sharded decode) on world_size-2 and -3 gloo groups with CPU tensors.  The codec
calls are injected, so here the oracle stands in for the CUDA shard kernels (test-only use)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, read_data


class OracleCodec:
    """test stand-in with the signature of snappy.jl_b200.multi.CudaCodec"""

    def compress_shard(self, shard, total_len):
        import pyoracle
        a = shard.numpy()
        nfrag = (a.size + 65535) // 65536
        # sjo_compress_fragments wants the whole stream; a shard starts on a fragment boundary and
        # only the total length matters for the table size, so present it as a stream prefix
        import ctypes
        out = np.empty(a.size + a.size // 6 + 64, dtype=np.uint8)
        sizes = np.zeros(max(nfrag, 1), dtype=np.uint32)
        L = pyoracle.lib()
        n = 0
        op = 0
        for f in range(nfrag):
            frag = np.ascontiguousarray(a[f * 65536:(f + 1) * 65536])
            table = np.full(16384, 0xFFFF, dtype=np.uint16)
            L.sjo_compress_fragment.restype = ctypes.c_size_t
            L.sjo_compress_fragment.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_uint32]
            c = L.sjo_compress_fragment(frag.ctypes.data, frag.size, out[op:].ctypes.data,
                                        table.ctypes.data, L.sjo_hashtable_entries(total_len))
            sizes[f] = c
            op += c
        return torch.from_numpy(out[:op].copy()), torch.from_numpy(sizes[:nfrag].astype(np.int32))

    def uncompress_shard(self, data, frag_offsets, out_len):
        import pyoracle
        d = data.numpy()
        fo = frag_offsets.numpy()
        out = np.zeros(out_len, dtype=np.uint8)
        for i in range(len(fo) - 1):
            n = min(65536, out_len - i * 65536)
            s = pyoracle.encode32(n) + d[fo[i]:fo[i + 1]].tobytes()
            out[i * 65536:i * 65536 + n] = pyoracle.uncompress_np(s)
        return torch.from_numpy(out)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pyoracle
        from snappy_jl_b200 import multi, synth
        streams = [synth.mix(5 + s, seed=100 + s, tail=(0 if s == 1 else 4321 * (s + 1))) for s in range(world)]
        if world == 2:
            streams[1] = np.frombuffer(read_data("urls.10K"), dtype=np.uint8)
        totals = [int(x.size) for x in streams]
        shards = []
        for s in range(world):
            lo, hi = multi.shard_bounds(totals[s], world)[rank]
            shards.append(torch.from_numpy(streams[s][lo:hi].copy()))
        stream, index = multi.compress_streams(shards, totals, OracleCodec())
        want = pyoracle.compress_np(streams[rank])
        assert np.array_equal(stream.numpy(), want), "assembled stream differs from single-process compress"
        idx = index.numpy()
        nfrag = (totals[rank] + 65535) // 65536
        assert idx.shape[0] == nfrag + 1 and idx[-1] == want.size
        runs = multi.uncompress_streams(stream, index, totals[rank], OracleCodec())
        for s in range(world):
            lo, hi = multi.shard_bounds(totals[s], world)[rank]
            assert np.array_equal(runs[s].numpy(), streams[s][lo:hi]), (rank, s)
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_streams_gloo(tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(tmp_path / ("ok%d" % r))


def test_shard_bounds():
    from snappy_jl_b200 import multi
    for total in (0, 1, 65536, 65537, 10 * 65536, 10 * 65536 + 5, 2 ** 30):
        for world in (1, 2, 3, 4, 8):
            b = multi.shard_bounds(total, world)
            assert b[0][0] == 0 and b[-1][1] == total
            for (lo, hi), (lo2, _) in zip(b[:-1], b[1:]):
                assert hi == lo2 and (lo % 65536 == 0 or lo == total) and (hi % 65536 == 0 or hi == total)
            nf = [(hi - lo + 65535) // 65536 for lo, hi in b]
            assert max(nf) - min(nf) <= 1
    assert multi.encode_header(2 ** 30) == bytes([0x80, 0x80, 0x80, 0x80, 0x04])
