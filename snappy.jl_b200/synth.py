"""Deterministic synthetic inputs for the parity tests and bench.py.

All randomness comes from a counter-based splitmix64 evaluated with numpy uint64 arithmetic, so a
(seed, size) pair names the same bytes on every machine.

`mix(nfrag, seed)`      -- BASELINE config 2: `nfrag` x 64 KiB fragments, the class of fragment f is
                           splitmix64(seed, f) mod 4:
    0  uniform random bytes                                   (incompressible, ratio ~1.0)
    1  the reference's own fuzz generator, test/runtests.jl:37-43: 64 dictionary words of
       1..16 random bytes, uniform picks                      (ratio ~0.27)
    2  Zipf-weighted lower-case word text, 4096-word vocabulary (ratio ~0.55)
    3  structured records: byte runs, short-period repeats and a few random bytes
       (exercises offset < length copies and 64-byte copy splitting)   (ratio ~0.1)
`source_like(nbytes, seed)` -- BASELINE config 5: identifier / keyword / indent / newline token stream.
`pages(...)`            -- BASELINE config 4: fixed-size pages cut from `mix`.
"""
import numpy as np

FRAGMENT = 65536
_U64 = np.uint64
_GOLD = _U64(0x9E3779B97F4A7C15)
_M1 = _U64(0xBF58476D1CE4E5B9)
_M2 = _U64(0x94D049BB133111EB)


def splitmix64(x):
    """Vectorised splitmix64 finaliser of (x + 1) * golden; x: uint64 array or int."""
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=_U64) + _U64(1)) * _GOLD
        z = (z ^ (z >> _U64(30))) * _M1
        z = (z ^ (z >> _U64(27))) * _M2
        return z ^ (z >> _U64(31))


def _stream(seed, lane, count, chunk_base=0):
    """`count` uint64 values of the stream (seed, lane)."""
    with np.errstate(over="ignore"):
        key = splitmix64(_U64(seed) * _U64(0x100000001B3) + _U64(lane))
        idx = np.arange(chunk_base, chunk_base + count, dtype=_U64)
        return splitmix64(idx ^ key)


def random_bytes(nbytes, seed, lane=0):
    words = (nbytes + 7) // 8
    return _stream(seed, lane, words).view(np.uint8)[:nbytes]


def _gather_segments(blob, seg_base, seg_len, seg_period, total):
    """out[j] = blob[base_i + (k mod period_i)] for the k-th byte of segment i, first `total` bytes."""
    ends = np.cumsum(seg_len)
    nseg = int(np.searchsorted(ends, total, side="left")) + 1
    seg_len = seg_len[:nseg]
    starts = ends[:nseg] - seg_len
    k = np.arange(int(ends[nseg - 1]), dtype=np.int64) - np.repeat(starts, seg_len)
    if seg_period is not None:
        k %= np.repeat(seg_period[:nseg], seg_len)
    k += np.repeat(seg_base[:nseg], seg_len)
    return blob[k[:total]]


def _dictionary_bytes(nbytes, seed):
    """test/runtests.jl:37-43: 64 words of 1..16 random bytes, uniform picks, concatenated."""
    wl = (_stream(seed, 11, 64) % _U64(16)).astype(np.int64) + 1
    wbase = np.cumsum(wl) - wl
    blob = random_bytes(int(wl.sum()), seed, 12)
    out = np.empty(nbytes, dtype=np.uint8)
    done, chunk, c = 0, 1 << 24, 0
    while done < nbytes:
        want = min(chunk, nbytes - done)
        picks = (_stream(seed, 13, want // 8 + 64, c * (chunk // 8 + 64)) % _U64(64)).astype(np.int64)
        lens = wl[picks]
        # mean word length 8.5 -> want/8+64 picks always cover `want` bytes? not guaranteed: top up
        while int(lens.sum()) < want:
            picks = np.concatenate([picks, picks])
            lens = wl[picks]
        out[done:done + want] = _gather_segments(blob, wbase[picks], lens, None, want)
        done += want
        c += 1
    return out


def _zipf_vocab(seed, nwords, lane, alphabet, min_len, span, sep):
    r = _stream(seed, lane, nwords)
    wl = (r % _U64(span)).astype(np.int64) + min_len
    letters = alphabet[(random_bytes(int(wl.sum()), seed, lane + 1) % len(alphabet))]
    wbase = np.cumsum(wl) - wl
    # append the separator to every word so that a pick is "word + sep"
    blob = np.empty(int(wl.sum()) + nwords, dtype=np.uint8)
    nb = wbase + np.arange(nwords)
    idx = np.arange(int(wl.sum()), dtype=np.int64) + np.repeat(np.arange(nwords), wl)
    blob[idx] = letters
    blob[nb + wl] = sep
    weights = 1.0 / np.arange(1, nwords + 1) ** 1.07
    cdf = np.cumsum(weights / weights.sum())
    return blob, nb, wl + 1, cdf


def _zipf_pick_bytes(nbytes, seed, lane, blob, base, lens, cdf):
    out = np.empty(nbytes, dtype=np.uint8)
    mean = float((lens * np.diff(np.concatenate([[0.0], cdf]))).sum())
    done, chunk, c = 0, 1 << 24, 0
    while done < nbytes:
        want = min(chunk, nbytes - done)
        npick = int(want / mean * 1.15) + 256
        u = (_stream(seed, lane, npick, c * (1 << 26)) >> _U64(11)).astype(np.float64) / float(1 << 53)
        picks = np.minimum(np.searchsorted(cdf, u), len(lens) - 1)
        l = lens[picks]
        while int(l.sum()) < want:
            picks = np.concatenate([picks, picks])
            l = lens[picks]
        out[done:done + want] = _gather_segments(blob, base[picks], l, None, want)
        done += want
        c += 1
    return out


def _text_bytes(nbytes, seed):
    alphabet = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    blob, base, lens, cdf = _zipf_vocab(seed, 4096, 21, alphabet, 2, 10, 32)
    return _zipf_pick_bytes(nbytes, seed, 23, blob, base, lens, cdf)


def _record_bytes(nbytes, seed):
    """runs (period 1), short-period repeats (2..7) and short random gaps"""
    blob = random_bytes(1 << 16, seed, 31)
    out = np.empty(nbytes, dtype=np.uint8)
    done, chunk, c = 0, 1 << 24, 0
    while done < nbytes:
        want = min(chunk, nbytes - done)
        nseg = want // 64 + 256
        r = _stream(seed, 32, nseg, c * (1 << 26))
        kind = (r % _U64(8)).astype(np.int64)          # 0-2 run, 3-5 pattern, 6-7 random gap
        length = ((r >> _U64(8)) % _U64(400)).astype(np.int64) + 8
        period = np.where(kind < 3, 1, np.where(kind < 6, ((r >> _U64(24)) % _U64(6)).astype(np.int64) + 2, 1 << 30))
        length = np.where(kind >= 6, ((r >> _U64(8)) % _U64(12)).astype(np.int64) + 1, length)
        base = ((r >> _U64(32)) % _U64((1 << 16) - 512)).astype(np.int64)
        while int(length.sum()) < want:
            kind, length, period, base = (np.concatenate([a, a]) for a in (kind, length, period, base))
        out[done:done + want] = _gather_segments(blob, base, length, period, want)
        done += want
        c += 1
    return out


def fragment_classes(nfrag, seed):
    return (splitmix64(np.arange(nfrag, dtype=_U64) + _U64(seed) * _U64(1000003)) % _U64(4)).astype(np.int64)


def mix(nfrag, seed=2026, tail=0):
    """`nfrag` whole fragments (+ `tail` extra bytes of class-2 text) of mixed compressibility."""
    cls = fragment_classes(nfrag, seed)
    out = np.empty(nfrag * FRAGMENT + tail, dtype=np.uint8)
    frames = out[: nfrag * FRAGMENT].reshape(nfrag, FRAGMENT)
    gens = (lambda n: random_bytes(n, seed, 1), lambda n: _dictionary_bytes(n, seed),
            lambda n: _text_bytes(n, seed), lambda n: _record_bytes(n, seed))
    for k in range(4):
        sel = np.nonzero(cls == k)[0]
        if sel.size:
            frames[sel] = gens[k](sel.size * FRAGMENT).reshape(sel.size, FRAGMENT)
    if tail:
        out[nfrag * FRAGMENT:] = _text_bytes(tail, seed + 1)
    return out


def source_like(nbytes, seed=2026):
    """Source-code-like token stream: Zipf identifiers/keywords, indentation, newlines."""
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz_ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)
    blob, base, lens, cdf = _zipf_vocab(seed, 2048, 41, alphabet, 1, 14, 32)
    # punctuation / layout tokens take the most probable slots
    layout = [b"\n    ", b"\n        ", b" = ", b"(", b")", b"{\n", b"}\n", b", ", b";\n", b"->", b"    ", b"\n"]
    extra = np.frombuffer(b"".join(layout), dtype=np.uint8)
    lb = np.cumsum([0] + [len(t) for t in layout[:-1]]) + blob.size
    blob = np.concatenate([blob, extra])
    slots = np.arange(len(layout)) * 2
    base = base.copy()
    lens = lens.copy()
    base[slots] = lb
    lens[slots] = [len(t) for t in layout]
    return _zipf_pick_bytes(nbytes, seed, 43, blob, base, lens, cdf)


def pages(count, page_size=4096, seed=2026):
    """`count` pages of `page_size` bytes cut from the mix generator (config 4)."""
    per_frag = FRAGMENT // page_size
    nfrag = (count + per_frag - 1) // per_frag
    return mix(nfrag, seed)[: count * page_size].reshape(count, page_size)


# ---- several buffers at once: worker processes fill one shared-memory segment (bench.py: 1 GiB of mix takes ~26 s
# and 1 GiB of source-like text ~43 s on one core) -----------------------------------------------------------------
def _gen_piece(job):
    """worker process: one piece straight into the shared file mapping"""
    kind, seed, nfrag, path, offset = job
    a = source_like(nfrag * FRAGMENT, seed=seed) if kind == "source" else mix(nfrag, seed=seed)
    m = np.memmap(path, dtype=np.uint8, mode="r+", offset=offset, shape=(a.size,))
    m[:] = a
    m.flush()
    del m
    return a.size


class ParallelGenerator:
    """add(kind, seed, nfrag) -> (offset, nbytes) of a piece (kind "source": source_like, anything else: mix);
    run() generates all pieces side by side in spawned workers (no CUDA state is forked) into one file mapping
    (/dev/shm when it has the room, else the temp directory); view() reads them."""

    def __init__(self, workers):
        self.jobs, self.total, self.workers = [], 0, max(1, int(workers))
        self.path, self.map = None, None

    def add(self, kind, seed, nfrag):
        off = self.total
        self.jobs.append([kind, int(seed), int(nfrag), None, off])
        self.total += nfrag * FRAGMENT
        return off, nfrag * FRAGMENT

    def run(self):
        import os
        import shutil
        import tempfile
        from concurrent.futures import ProcessPoolExecutor
        from multiprocessing import get_context
        if not self.jobs:
            return
        where = None
        try:
            if shutil.disk_usage("/dev/shm").free > self.total + (1 << 30):
                where = "/dev/shm"
        except Exception:
            pass
        fd, self.path = tempfile.mkstemp(prefix="sb200_gen_", dir=where)
        os.ftruncate(fd, max(self.total, 1))
        os.close(fd)
        for j in self.jobs:
            j[3] = self.path
        order = sorted(self.jobs, key=lambda j: -j[2])  # longest first; the pieces are independent
        with ProcessPoolExecutor(max_workers=min(self.workers, len(order)), mp_context=get_context("spawn")) as ex:
            list(ex.map(_gen_piece, [tuple(j) for j in order]))
        self.map = np.memmap(self.path, dtype=np.uint8, mode="r", shape=(max(self.total, 1),))

    def view(self, off, nbytes):
        return self.map[off: off + nbytes]

    def close(self):
        import os
        self.map = None
        if self.path is not None:
            try:
                os.unlink(self.path)
            except Exception:
                pass
            self.path = None
