"""Multi-GPU sharding of the path on whole-fragment boundaries.

One process per GPU.  A stream of `total_len` bytes is cut into contiguous runs of whole 64 KiB
fragments, one run per rank (fragments are independent: the hash table is reset per fragment,
src/Snappy.jl:30, and candidates never precede the fragment start, src/internal.jl:129,190).  The
only global facts a rank needs are the TOTAL length (table size and header, src/Snappy.jl:26-27)
and where its bytes go.  The one real exchange step of the path is therefore:

    sizes   = all_gather(compressed byte count of my run)            (8 bytes per rank)
    offsets = header_len + exclusive_scan(sizes)
    stream  = segments placed at `offsets` in the owner's buffer      (over NVLink)

`compress_streams` does this for W streams at once (stream s is owned by rank s), which turns the
assembly into one all_to_all_single with split sizes taken from the gathered size matrix.

The codec calls are injected (`codec`), so the same host logic runs under gloo on CPU tensors in
the tests (with the oracle as the stand-in codec) and under NCCL with the CUDA path in production.
"""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

_TRACE = bool(os.environ.get("SNAPPY_B200_TRACE_MULTI"))


class _Trace:
    """SNAPPY_B200_TRACE_MULTI=1: rank 0 prints where the time of a call went (synchronising marks)."""

    def __init__(self, name):
        self.name, self.marks = name, []
        if _TRACE:
            torch.cuda.synchronize()
            self.t = time.perf_counter()

    def mark(self, what):
        if _TRACE:
            torch.cuda.synchronize()
            now = time.perf_counter()
            self.marks.append("%s %.2f" % (what, (now - self.t) * 1e3))
            self.t = now

    def done(self):
        if _TRACE and dist.get_rank() == 0:
            print("[multi] %s: %s ms" % (self.name, " | ".join(self.marks)), flush=True)

FRAGMENT = 65536


def shard_bounds(total_len, world):
    """Byte ranges [lo, hi) of each rank's run of whole fragments; the last run may be ragged."""
    nfrag = (total_len + FRAGMENT - 1) // FRAGMENT
    base, extra = divmod(nfrag, world)
    bounds, f = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        lo = min(f * FRAGMENT, total_len)
        hi = min((f + k) * FRAGMENT, total_len)
        bounds.append((lo, hi))
        f += k
    return bounds


def encode_header(total_len):
    """varint.jl:46-69 (host)."""
    out = bytearray()
    v = int(total_len)
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


class CudaCodec:
    """Production codec: the shard entry points of libsnappy_b200.so."""

    def compress_shard(self, shard, total_len):
        from . import device
        out, sizes = device.compress_shard_device(shard, total_len, want_sizes=True)
        return out, sizes

    def uncompress_shard(self, data, frag_offsets, out_len):
        from . import device
        return device.uncompress_shard_device(data, frag_offsets, out_len)

    # batched forms: every shard a rank holds goes through ONE kernel pass, which keeps the GPU
    # full when the rank's work is many small shards (W streams x 1/W each)
    def compress_shards(self, shards, total_lens):
        from . import device
        return device.compress_shards_device(shards, total_lens)

    def uncompress_shards(self, datas, frag_offsets, out_lens):
        from . import device
        return device.uncompress_shards_device(datas, frag_offsets, out_lens)


def compress_streams(shards, total_lens, codec, group=None):
    """shards[s]: this rank's run of stream s (uint8 tensor), total_lens[s]: stream s's total length.
    len(shards) == world size; stream s is assembled on rank s.

    Returns (stream, frag_index): the complete stream owned by this rank (header + all ranks'
    segments) and its side index ((nfrag+1) int64 offsets into the stream)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(shards) == world == len(total_lens)
    dev = shards[0].device
    tr = _Trace("compress_streams")
    if hasattr(codec, "compress_shards"):
        pairs = codec.compress_shards(shards, total_lens)
    else:
        pairs = [codec.compress_shard(shards[s], total_lens[s]) for s in range(world)]
    tr.mark("codec")
    segs = [p[0] for p in pairs]
    frag_sizes = [p[1].to(torch.int64) for p in pairs]
    # exchange 1: compressed byte counts (the path's only data-dependent global fact)
    mine = torch.tensor([int(x.numel()) for x in segs], dtype=torch.int64, device=dev)
    matrix = torch.empty(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(matrix, mine, group=group)
    matrix = matrix.view(world, world).cpu()          # matrix[r][s] = bytes rank r made for stream s
    tr.mark("sizes")
    in_splits = [int(x) for x in matrix[rank]]
    out_splits = [int(matrix[r][rank]) for r in range(world)]
    hdr = encode_header(total_lens[rank])
    stream = torch.empty(len(hdr) + sum(out_splits), dtype=torch.uint8, device=dev)
    stream[: len(hdr)] = torch.frombuffer(bytearray(hdr), dtype=torch.uint8).to(dev)
    # exchange 2: segment assembly -- every segment lands at header + exclusive-scan offset
    send = torch.cat(segs) if world > 1 else segs[0]
    tr.mark("alloc+cat")
    dist.all_to_all_single(stream[len(hdr):], send, out_splits, in_splits, group=group)
    tr.mark("all_to_all")
    # side index of my stream: fragment sizes of every rank's run, in rank order
    nf_mine = torch.tensor([int(x.numel()) for x in frag_sizes], dtype=torch.int64, device=dev)
    nf_matrix = torch.empty(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(nf_matrix, nf_mine, group=group)
    nf_matrix = nf_matrix.view(world, world).cpu()
    fs_in = [int(x) for x in nf_matrix[rank]]
    fs_out = [int(nf_matrix[r][rank]) for r in range(world)]
    all_sizes = torch.empty(sum(fs_out), dtype=torch.int64, device=dev)
    dist.all_to_all_single(all_sizes, torch.cat(frag_sizes), fs_out, fs_in, group=group)
    index = torch.empty(all_sizes.numel() + 1, dtype=torch.int64, device=dev)
    index[0] = len(hdr)
    index[1:] = len(hdr) + torch.cumsum(all_sizes, 0)
    tr.mark("index")
    tr.done()
    return stream, index


def uncompress_streams(stream, index, total_len, codec, group=None):
    """Inverse: the owner of each stream scatters every rank's compressed range; each rank decodes
    its own run of every stream.  The uncompressed output stays sharded (SURVEY.md section 5:
    gathering it on one GPU would be NVLink-ingest bound).  Returns the list of decoded runs, one
    per stream."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = stream.device
    lens = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens, torch.tensor([total_len], dtype=torch.int64, device=dev), group=group)
    lens = [int(x) for x in lens.cpu()]
    idx = index.cpu().numpy()
    bounds = shard_bounds(total_len, world)
    frag_lo = [lo // FRAGMENT for lo, _ in bounds] + [(total_len + FRAGMENT - 1) // FRAGMENT]
    in_splits = [int(idx[frag_lo[r + 1]] - idx[frag_lo[r]]) for r in range(world)]
    splits_t = torch.tensor(in_splits, dtype=torch.int64, device=dev)
    matrix = torch.empty(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(matrix, splits_t, group=group)
    matrix = matrix.view(world, world).cpu()
    out_splits = [int(matrix[s][rank]) for s in range(world)]
    hdr = int(idx[0])
    recv = torch.empty(sum(out_splits), dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv, stream[hdr:], out_splits, in_splits, group=group)
    # the per-fragment offsets of my run of every stream travel the same way
    rel = []
    for r in range(world):
        part = idx[frag_lo[r]: frag_lo[r + 1] + 1] - idx[frag_lo[r]]
        rel.append(torch.from_numpy(np.ascontiguousarray(part)).to(dev))
    n_in = [int(x.numel()) for x in rel]
    n_mat = torch.empty(world * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(n_mat, torch.tensor(n_in, dtype=torch.int64, device=dev), group=group)
    n_mat = n_mat.view(world, world).cpu()
    n_out = [int(n_mat[s][rank]) for s in range(world)]
    offs = torch.empty(sum(n_out), dtype=torch.int64, device=dev)
    dist.all_to_all_single(offs, torch.cat(rel), n_out, n_in, group=group)
    datas, fos, out_lens, a, b = [], [], [], 0, 0
    for s in range(world):
        lo, hi = shard_bounds(lens[s], world)[rank]
        datas.append(recv[a: a + out_splits[s]])
        fos.append(offs[b: b + n_out[s]])
        out_lens.append(hi - lo)
        a += out_splits[s]
        b += n_out[s]
    if hasattr(codec, "uncompress_shards"):
        return codec.uncompress_shards(datas, fos, out_lens)
    return [codec.uncompress_shard(datas[s], fos[s].contiguous(), out_lens[s]) for s in range(world)]


# ------------------------------------------------------------------------------------------------------
# Library path (CUDA): the exchange steps live in libsnappy_b200.so (csrc/multi_host.inc, csrc/multi.cuh):
# ncclAllGather of the byte counts on the compute stream, fragments and side-index entries stored straight into
# the owner's buffer through cudaIpc-mapped peer pointers (NVLink), the inverse by peer loads.  torch.distributed
# only carries the 128-byte NCCL id to the other ranks.  The functions above remain the host-logic reference that
# the gloo tests run on CPU tensors.
# ------------------------------------------------------------------------------------------------------
import ctypes


class _DevMem:
    """A device pointer + length as a __cuda_array_interface__ object (zero-copy torch view of arena memory)."""

    def __init__(self, ptr, nbytes, typestr="|u1", itemsize=1):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2, "strides": None}


def _view_u8(ptr, n, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevMem(ptr, n), device=device)


def _view_i64(ptr, n, device):
    if n == 0 or not ptr:
        return torch.empty(0, dtype=torch.int64, device=device)
    return torch.as_tensor(_DevMem(ptr, n * 8, "<i8", 8), device=device)


class LibComm:
    """snappy_b200_comm: `world` ranks, `nlocal` of them hosted by this process (1 under torchrun; all of them in a
    loopback world on one GPU).  Stream s is owned by rank s % world.

    compress(shards, total_lens) / uncompress(streams, indexes, total_lens) are collective; the per-rank argument
    lists are rank-major over the local ranks: entry [lr * nstreams + s]."""

    def __init__(self, group=None, loopback_world=None, device=None):
        from . import _abi
        from .api import _check
        self._abi, self._check = _abi, _check
        self.handle = ctypes.c_void_p(0)
        if loopback_world is not None:
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
            with torch.cuda.device(self.device):
                _check(_abi.lib().snappy_b200_comm_create_loopback(int(loopback_world), ctypes.byref(self.handle)))
            self.world, self.nlocal, self.rank0 = int(loopback_world), int(loopback_world), 0
            return
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        ident = (ctypes.c_uint8 * 128)()
        if world > 1:
            t = torch.zeros(128, dtype=torch.uint8, device=self.device)
            if rank == 0:
                _check(_abi.lib().snappy_b200_comm_unique_id(ident))
                t.copy_(torch.frombuffer(bytearray(bytes(ident)), dtype=torch.uint8))
            dist.broadcast(t, src=0, group=group)  # plumbing: the id is all torch.distributed carries
            raw = bytes(t.cpu().numpy().tobytes())
            ident = (ctypes.c_uint8 * 128).from_buffer_copy(raw)
        with torch.cuda.device(self.device):
            _check(_abi.lib().snappy_b200_comm_create(ident, rank, world, ctypes.byref(self.handle)))
        self.world, self.nlocal, self.rank0 = world, 1, rank

    def close(self):
        self.__dict__.pop("_views", None)  # views over arena memory die with the arenas
        if self.handle:
            self._abi.lib().snappy_b200_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def owns(self, s):
        return self.rank0 <= s % self.world < self.rank0 + self.nlocal

    def _view(self, make, ptr, n):
        """zero-copy tensor over arena memory; the same (address, length) comes back call after call while the arena
        keeps its size, and building a tensor from a foreign pointer costs ~0.1 ms: keep the last few"""
        cache = self.__dict__.setdefault("_views", {})
        key = (make.__name__, int(ptr), int(n))
        t = cache.get(key)
        if t is None:
            if len(cache) > 256:
                cache.clear()
            t = cache[key] = make(ptr, n, self.device)
        return t

    def compress(self, shards, total_lens):
        """shards: nlocal * nstreams uint8 CUDA tensors (or None for an empty run).  Returns (streams, indexes,
        lens): per stream a zero-copy view of the assembled stream and of its side index (None for streams owned
        elsewhere) and the stream length (known everywhere).  Views stay valid until the next compress()."""
        S = len(total_lens)
        assert len(shards) == self.nlocal * S
        n = len(shards)
        ptrs = (ctypes.c_void_p * n)(*[(t.data_ptr() if t is not None and t.numel() else 0) for t in shards])
        lens = (ctypes.c_size_t * n)(*[(t.numel() if t is not None else 0) for t in shards])
        totals = (ctypes.c_uint64 * S)(*[int(x) for x in total_lens])
        o_streams = (ctypes.c_void_p * S)()
        o_lens = (ctypes.c_size_t * S)()
        o_index = (ctypes.c_void_p * S)()
        with torch.cuda.device(self.device):
            self._check(self._abi.lib().snappy_b200_comm_compress(
                self.handle, ptrs, lens, totals, S, o_streams, o_lens, o_index,
                ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        streams, indexes = [], []
        for s in range(S):
            if o_streams[s]:
                nfrag = (int(total_lens[s]) + FRAGMENT - 1) // FRAGMENT
                streams.append(self._view(_view_u8, o_streams[s], int(o_lens[s])))
                indexes.append(self._view(_view_i64, o_index[s], nfrag + 1))
            else:
                streams.append(None)
                indexes.append(None)
        return streams, indexes, [int(x) for x in o_lens]

    def uncompress(self, streams, indexes, total_lens, outs=None, statuses_out=None):
        """streams / indexes: per stream, given where a local rank owns it (None elsewhere; index may be None:
        the owner parses).  Returns the local ranks' decoded runs, nlocal * nstreams tensors (rank-major)."""
        S = len(total_lens)
        sp = (ctypes.c_void_p * S)(*[(t.data_ptr() if t is not None and t.numel() else 0) for t in streams])
        sl = (ctypes.c_size_t * S)(*[(t.numel() if t is not None else 0) for t in streams])
        ip = (ctypes.c_void_p * S)(*[(t.data_ptr() if t is not None and t.numel() else 0) for t in indexes])
        totals = (ctypes.c_uint64 * S)(*[int(x) for x in total_lens])
        if outs is None:
            outs = []
            for lr in range(self.nlocal):
                for s in range(S):
                    lo, hi = shard_bounds(int(total_lens[s]), self.world)[self.rank0 + lr]
                    outs.append(torch.empty(hi - lo, dtype=torch.uint8, device=self.device))
        op = (ctypes.c_void_p * len(outs))(*[(t.data_ptr() if t.numel() else 0) for t in outs])
        st = (ctypes.c_int * S)()
        with torch.cuda.device(self.device):
            rc = self._abi.lib().snappy_b200_comm_uncompress(
                self.handle, sp, sl, ip, totals, S, op, st,
                ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if statuses_out is not None:
            statuses_out[:] = [int(x) for x in st]
        self._check(rc)
        return outs
