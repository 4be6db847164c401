"""Lane-level emulation (numpy, 32 lanes) of the warp-speculative compress_fragment used by
snappy.jl_b200/csrc/compress.cuh.  Development aid: validates the *algorithm* (probe speculation,
intra-warp forwarding, bail-out rules, record emission) against the oracle on CPU, since CUDA
cannot run in the build container.  Not part of the product path."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))

MUL = 0x1E35A7BD
LANES = np.arange(32)


def probe_offsets(limit=70000):
    po, skip, off = [0], 32, 0
    while off <= limit:
        b = skip >> 5
        skip += b
        off += b
        po.append(off)
    while len(po) % 32 != 1 or len(po) < 321:
        po.append(po[-1] + 4096)
    return np.array(po, dtype=np.int64)


PO = probe_offsets()


class Frag:
    def __init__(self, data, entries):
        self.n = len(data)
        self.F = np.concatenate([np.frombuffer(data, dtype=np.uint8), np.zeros(64, np.uint8)]).astype(np.int64)
        self.shift = 32 - (entries.bit_length() - 1)
        self.T = np.zeros(entries, dtype=np.int64)
        self.records = []  # (next_emit, ip, offset, M) ; M == 0 -> literal only

    def w(self, q):
        q = np.asarray(q)
        F = self.F
        return F[q] | (F[q + 1] << 8) | (F[q + 2] << 16) | (F[q + 3] << 24)

    def h(self, w):
        return ((w * MUL) & 0xFFFFFFFF) >> self.shift


def warp_match_length(fr, a, b):
    n, F, total = fr.n, fr.F, 0
    while True:
        pb = b + 4 * LANES
        x = fr.w(np.minimum(a + 4 * LANES, n)) ^ fr.w(np.minimum(pb, n))
        cnt = np.where(x == 0, 4, 3)
        for k in (2, 1, 0):  # first differing byte
            cnt = np.where((x & (0xFF << (8 * k))) != 0, k, cnt)
        cnt = np.where(x == 0, 4, cnt)
        cnt = np.where(pb < n, np.minimum(cnt, np.maximum(n - pb, 0)), 0)
        stop = cnt < 4
        if stop.any():
            f = int(np.argmax(stop))
            return total + 4 * f + int(cnt[f])
        total += 128; a += 128; b += 128


def compress_fragment(data, entries):
    fr = Frag(data, entries)
    n, T = fr.n, fr.T
    lim = n - 16
    ip = next_emit = 0
    if n >= 15:
        done = False
        while not done:
            # ---------------- scan
            s = ip + 1
            base = 0
            hit = False
            while True:
                i = base + LANES
                p = s + PO[i]
                pn = s + PO[i + 1]
                valid = pn <= lim
                ps = np.where(valid, p, 0)
                w = fr.w(ps)
                hh = fr.h(w)
                c = T[hh].copy()
                # forwarding: largest j < i, valid, same hash
                for li in range(32):
                    for lj in range(li - 1, -1, -1):
                        if valid[lj] and hh[lj] == hh[li]:
                            c[li] = p[lj]
                            break
                eq = valid & (fr.w(c) == w)
                first_invalid = int(np.argmax(~valid)) if (~valid).any() else 32
                first_hit = int(np.argmax(eq)) if eq.any() else 32
                if first_hit < first_invalid:
                    for li in range(first_hit + 1):
                        T[hh[li]] = p[li]
                    ip, cand, hit = int(p[first_hit]), int(c[first_hit]), True
                    break
                if first_invalid < 32:
                    break
                for li in range(32):
                    T[hh[li]] = p[li]
                base += 32
            if not hit:
                break
            # ---------------- copy chain
            lit_from = next_emit
            while True:
                # speculative probes for e_l = ip + 4 + l (table state BEFORE this step's inserts)
                e = ip + 4 + LANES
                ok = e < lim
                es = np.where(ok, e, 1)
                we = fr.w(es)
                he = fr.h(we)
                hm = fr.h(fr.w(es - 1))
                ce = np.where(hm == he, es - 1, T[he])
                eqe = fr.w(ce) == we
                M = 4 + warp_match_length(fr, cand + 4, ip + 4)
                fr.records.append((lit_from, ip, ip - cand, M))
                ip += M
                next_emit = lit_from = ip
                if ip >= lim:
                    done = True
                    break
                L = M - 4
                if L < 32:
                    assert e[L] == ip and ok[L]
                    c, same, hprev, hcur = int(ce[L]), bool(eqe[L]), int(hm[L]), int(he[L])
                else:
                    hprev = int(fr.h(fr.w(ip - 1))); wc = fr.w(ip); hcur = int(fr.h(wc))
                    c = ip - 1 if hprev == hcur else int(T[hcur])
                    same = bool(fr.w(c) == wc)
                T[hprev] = ip - 1
                T[hcur] = ip
                if not same:
                    break
                cand = c
    if next_emit < n:
        fr.records.append((next_emit, n, 0, 0))
    return fr


# ---------------- emitter (batch of records -> bytes), mirrors the emitter warp
def op_size(off, ln):
    return 2 if (ln < 12 and off < 2048) else 3


def copy_ops(off, M):
    ops = []
    if M == 0:
        return ops
    if M >= 12:
        while M >= 68:
            ops.append(64); M -= 64
        if M > 64:
            ops.append(60); M -= 60
    ops.append(M)
    return ops


def emit(fr):
    out = bytearray()
    F = fr.F
    for (a, b, off, M) in fr.records:
        ln = b - a
        if ln > 0:
            nm1 = ln - 1
            if ln < 60:
                out.append((nm1 << 2) & 0xFF)
            else:
                cnt = 1 if nm1 <= 0xFF else (2 if nm1 <= 0xFFFF else 3)
                out.append((59 + cnt) << 2)
                for k in range(cnt):
                    out.append((nm1 >> (8 * k)) & 0xFF)
            out += bytes(F[a:b].astype(np.uint8))
        for l in copy_ops(off, M):
            if l < 12 and off < 2048:
                out.append((1 + ((l - 4) << 2) + ((off >> 3) & 0xE0)) & 0xFF)
                out.append(off & 0xFF)
            else:
                u = 2 + ((l - 1) << 2) + (off << 8)
                out += bytes([u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF])
    return bytes(out)


def compress(data):
    import pyoracle
    n = len(data)
    entries = 256
    while entries < 16384 and entries < n:
        entries <<= 1
    out = bytearray(pyoracle.encode32(n))
    for s in range(0, n, 65536):
        out += emit(compress_fragment(data[s:s + 65536], entries))
    return bytes(out)


if __name__ == "__main__":
    import pyoracle
    from snappy_jl_b200 import synth
    rng = np.random.default_rng(1)
    cases = [open(os.path.join(ROOT, "tests/data", f), "rb").read()[:140000]
             for f in ("html", "alice29.txt", "urls.10K", "geo.protodata", "fireworks.jpeg", "kppkn.gtb")]
    cases += [synth.mix(4, seed=k).tobytes() for k in (1, 2)]
    cases += [b"A" * 100000, b"xyz" * 21845 + b"x", bytes(range(60)), b"a" * 70 + bytes(range(100, 130)), b"", b"abc"]
    cases += [rng.integers(0, a, sz, dtype=np.uint8).tobytes() for a in (2, 3, 256) for sz in (15, 16, 17, 31, 32, 33, 40, 100, 300, 5000, 65536, 65551)]
    for i, d in enumerate(cases):
        got, want = compress(d), pyoracle.compress(d)
        print(i, len(d), len(want), "OK" if got == want else "MISMATCH")
        assert got == want
