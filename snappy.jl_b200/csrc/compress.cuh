// compress.cuh -- Snappy.jl-exact fragment compressor for sm_100a.
//
// One CTA (one warp) per independent <=64 KiB fragment (src/Snappy.jl:29-33).  The fragment is
// staged in shared memory by a TMA bulk copy next to the u16 hash table (64 KiB + 32 KiB), so every
// probe / candidate load of the serial decision chain is a shared-memory access.  The decisions
// (skip heuristic, probes, match extension, tag choice) are exactly those of
// src/internal.jl:127-329, so the bytes equal Snappy.jl's.
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr u32 kFragPad = 32;  // zeroed bytes after the fragment so word loads may run past n
constexpr u32 kCompressSmemBytes = kBlockSize + kFragPad + kMaxTableEntries * 2 + 16;

struct FragEmitter {
    u8* out;   // scratch slot of this fragment (global)
    u32 op;    // bytes written so far (warp-uniform)
    u32 lane;
    u32 lit_short = 60;  // literals below this take the one-byte header: 60 = Snappy.jl (:271), 61 = libsnappy (option `rules`)

    // src/internal.jl:252-287 -- tag, then the bytes.  lanes copy the literal cooperatively.
    __device__ __forceinline__ void literal(const u8* F, u32 from, u32 len) {
        u32 n = len - 1;
        u32 hdr;
        if (len < lit_short) {  // :271 (a 60-byte literal takes the 2-byte header, like the reference)
            hdr = 1;
            if (lane == 0) out[op] = (u8)(n << 2);
        } else {
            u32 count = (n > 0xff) ? ((n > 0xffff) ? 3u : 2u) : 1u;  // :279-282, n < 65536 here
            hdr = 1 + count;
            if (lane == 0) {
                out[op] = (u8)((59 + count) << 2);
                out[op + 1] = (u8)n;
                if (count > 1) out[op + 2] = (u8)(n >> 8);
                if (count > 2) out[op + 3] = (u8)(n >> 16);
            }
        }
        u8* dst = out + op + hdr;
        const u8* src = F + from;
        for (u32 i = lane; i < len; i += 32) dst[i] = src[i];
        op += hdr + len;
    }

    // src/internal.jl:289-304
    __device__ __forceinline__ void copy_upto_64(u32 offset, u32 len) {
        if (len < 12 && offset < 2048) {
            if (lane == 0) {
                out[op] = (u8)(1 + ((len - 4) << 2) + ((offset >> 3) & 0xe0));
                out[op + 1] = (u8)offset;
            }
            op += 2;
        } else {
            if (lane == 0) {
                u32 u = 2 + ((len - 1) << 2) + (offset << 8);
                out[op] = (u8)u;
                out[op + 1] = (u8)(u >> 8);
                out[op + 2] = (u8)(u >> 16);
            }
            op += 3;
        }
    }

    // src/internal.jl:306-329
    __device__ __forceinline__ void copy(u32 offset, u32 len) {
        if (len >= 12) {
            while (len >= 68) {
                copy_upto_64(offset, 64);
                len -= 64;
            }
            if (len > 64) {
                copy_upto_64(offset, 60);
                len -= 60;
            }
        }
        copy_upto_64(offset, len);
    }
};

// Longest common prefix of F[a..) and F[b..n), a < b, all 32 lanes cooperating: lane l compares
// the 4 bytes at +4l; __ballot_sync/__ffs pick the first mismatching lane.  Same value as
// find_match_length (src/internal.jl:344-387) which is bounded by the fragment end only.
__device__ __forceinline__ u32 warp_match_length(const u8* F, u32 a, u32 b, u32 n, u32 lane) {
    u32 total = 0;
    for (;;) {
        u32 pb = b + 4 * lane;
        u32 cnt = 0;
        if (pb < n) {
            u32 x = lds32u(F, a + 4 * lane) ^ lds32u(F, pb);
            cnt = x ? ((u32)(__ffs((int)x) - 1) >> 3) : 4u;
            u32 room = n - pb;
            cnt = cnt < room ? cnt : room;
        }
        u32 stop = __ballot_sync(kFullMask, cnt < 4);
        if (stop) {
            u32 first = (u32)__ffs((int)stop) - 1;
            return total + 4 * first + __shfl_sync(kFullMask, cnt, first);
        }
        total += 128;
        a += 128;
        b += 128;
    }
}

// Stage one fragment (n bytes at g) into shared memory: TMA bulk copy for the 16-byte-aligned
// body, plain loads for a ragged tail / unaligned source; zero the pad.  `phase` is the number of
// bulk copies already waited for on `bar` (0 on first use: the barrier is initialised here);
// returns the updated count.
__device__ __forceinline__ u32 load_fragment(u8* F, u64* bar, const u8* g, u32 n, u32 lane,
                                             u32 phase) {
#ifdef SB200_CPU_EMU  // tools/cpu_warp: no TMA, no mbarrier -- the lanes copy
    (void)bar;
    __syncwarp();
    for (u32 i = lane; i < n; i += 32) F[i] = g[i];
    if (lane < kFragPad) F[n + lane] = 0;
    __syncwarp();
    return phase;
#else
    u32 body = 0;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) body = n & ~15u;
    if (phase == 0) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        phase = 1;  // bit 0: initialised; bits 1..: completed phases
    }
    __syncwarp();
    if (body && lane == 0) {
        // order earlier generic-proxy reads of F before the async-proxy overwrite
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(bar, body);
        tma_bulk_g2s(F, g, body, bar);
    }
    for (u32 i = body + lane; i < n; i += 32) F[i] = g[i];
    if (lane < kFragPad) F[n + lane] = 0;
    if (body) {
        mbar_wait(bar, (phase >> 1) & 1);
        phase += 2;
    }
    __syncwarp();
    return phase;
#endif
}

// Zero the hash table (position per hash, 0 == empty; the reference stores pos-1 with
// 0xffff == empty, src/internal.jl:177-191 -- the candidate positions are the same).
__device__ __forceinline__ void reset_table(u16* T, u32 shift, u32 lane) {
    const u32 entries = 1u << (32 - shift);
    uint4* t4 = reinterpret_cast<uint4*>(T);
    for (u32 i = lane; i < entries / 8; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
}

// Serial form of compress_fragment! (src/internal.jl:127-250), executed warp-uniformly: every
// lane follows the same decisions; lanes cooperate on match extension and literal copies.
// F: fragment in shared memory (n bytes + zero pad), T: zeroed table, shift = 32 - log2(entries).
// kLib (option `rules`): libsnappy's ip_limit = n - 15 and bucket = ((w * mul) >> shift) & hmask.
template <bool kLib = false>
__device__ __forceinline__ void compress_fragment_serial(const u8* F, u16* T, const u32 n,
                                                         const u32 shift, FragEmitter& em, const u32 hmask = ~0u) {
    const u32 lane = em.lane;
    const int lim = (int)n - (kLib ? 15 : 16);  // ip_limit, 0-based (src/internal.jl:131)
    auto bucket = [&](u32 w) { return kLib ? (((w * kHashMul) >> shift) & hmask) : ((w * kHashMul) >> shift); };
    u32 ip = 0, next_emit = 0;

    if (n >= kInputMargin) {
        for (;;) {
            // ---- scan for a 4-byte match, src/internal.jl:162-194
            u32 skip = 32;
            ip += 1;
            u32 next_ip = ip;
            u32 next_hash = bucket(lds32u(F, ip));
            u32 cand = 0;
            bool bail = false;
            for (;;) {
                ip = next_ip;
                u32 h = next_hash;
                u32 between = skip >> 5;
                skip += between;
                next_ip = ip + between;
                if ((int)next_ip > lim) { bail = true; break; }  // :175
                next_hash = bucket(lds32u(F, next_ip));
                cand = T[h];
                __syncwarp();
                if (lane == 0) T[h] = (u16)ip;
                __syncwarp();
                if (lds32u(F, cand) == lds32u(F, ip)) break;
            }
            if (bail) break;
            em.literal(F, next_emit, ip - next_emit);  // :200
            // ---- copy chain, src/internal.jl:211-239
            for (;;) {
                u32 matched = 4 + warp_match_length(F, cand + 4, ip + 4, n, lane);  // :216
                em.copy(ip - cand, matched);
                ip += matched;
                next_emit = ip;
                if ((int)ip >= lim) { bail = true; break; }  // :222
                u32 w = lds32u(F, ip);
                u32 hp = bucket(lds32u(F, ip - 1));
                u32 hc = bucket(w);
                __syncwarp();
                if (lane == 0) T[hp] = (u16)(ip - 1);  // :233
                __syncwarp();
                cand = T[hc];                          // :234
                __syncwarp();
                if (lane == 0) T[hc] = (u16)ip;        // :235
                __syncwarp();
                if (w != lds32u(F, cand)) break;       // :238
            }
            if (bail) break;
        }
    }
    if (next_emit < n) em.literal(F, next_emit, n - next_emit);  // :242-248
}

#ifdef SB200_EXPERIMENTS  // the first two designs (74 ms and 115 ms per GiB): kept bit-exact-tested, not shipped
// =============================================================================================
// K1 v2 (experiment): warp-specialised, lane-speculative fragment compressor.
//
// CTA = 2 warps per fragment.  Warp 0 ("decider") walks the reference's serial decision chain
// (src/internal.jl:162-239) but evaluates up to 32 table probes per round in its lanes:
//   * scan: lane i takes the i-th probe position of the skip sequence (:167-194); a probe whose
//     hash equals an earlier lane's sees that lane's position (intra-warp forwarding == the table
//     insert the reference would have made, :191); the first hit wins and only lanes up to it
//     commit their inserts (later lane wins on equal hashes).
//   * copy chain: while the match length is being measured, the lanes pre-evaluate the post-copy
//     probe (:228-238) for every possible end position ip+4 .. ip+35; the lane owning the real
//     end position supplies candidate + verdict.
//   Positions are owned by lanes (lane = position & 31) with a two-slot register window of the
//   bytes around them, refilled one step ahead, so hashes need no shared-memory load on the chain.
// Warp 1 ("emitter") turns the (literal, copy) records the decider pushes through a shared-memory
// ring into tag bytes (src/internal.jl:252-329), 32 records per batch with a warp prefix sum for
// the output positions.  Output bytes are identical to the serial form.
// =============================================================================================
constexpr u32 kRing = 128;               // records in the decider -> emitter ring
constexpr u32 kPoEntries = 352;          // probe-offset table (skip sequence of :162-172)
constexpr u32 kRecDone = 0xffffffffu;
constexpr u32 kNoTag = 0xffffffffu;
constexpr u32 kShortLiteral = 16;
constexpr u32 kCompress2SmemBytes =
    kBlockSize + 64 /*pad*/ + kMaxTableEntries * 2 + kPoEntries * 4 + kRing * 16 + 64;

__device__ __forceinline__ uint4 lds_volatile_v4(const uint4* p) {
    uint4 r;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ void sts_volatile_v4(uint4* p, uint4 v) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ u32 lds_volatile_u32(const u32* p) {
    u32 r;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ void sts_volatile_u32(u32* p, u32 v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ u32 rotr32(u32 x, u32 r) { return __funnelshift_r(x, x, r); }
__device__ __forceinline__ u32 mask_le(u32 i) { return (i >= 31) ? kFullMask : ((2u << i) - 1); }

struct Decider {
    const u8* F;
    u16* T;
    const u32* PO;
    uint4* ring;
    u32* tail_ptr;
    u32 n, shift, lane;
    int lim;
    // ring producer state
    u32 head, tail_cache;
    // register window: slot s holds the two aligned words covering bytes [tag-1, tag+4)
    u32 tag0, a0, b0, tag1, a1, b1, ptag, pa, pb;

    __device__ __forceinline__ u32 hash(u32 w) const { return (w * kHashMul) >> shift; }

    __device__ __forceinline__ void push(u32 lit_from, u32 lit_len, u32 off, u32 M) {
        if (head - tail_cache >= kRing) {
            do { tail_cache = lds_volatile_u32(tail_ptr); } while (head - tail_cache >= kRing);
        }
        if (lane == 0)
            sts_volatile_v4(&ring[head % kRing], make_uint4(lit_from | (((head / kRing) & 1u) << 31), lit_len, off, M));
        head++;
    }

    // words around position q (q >= 1): W = bytes [q, q+4), Wm = bytes [q-1, q+3)
    __device__ __forceinline__ void lookup(u32 q, u32& W, u32& Wm) const {
        const u32 s = (q >> 5) & 1;
        u32 lo, hi;
        if ((s ? tag1 : tag0) == q) {
            lo = s ? a1 : a0;
            hi = s ? b1 : b0;
        } else {
            const u32* w = reinterpret_cast<const u32*>(F + ((q - 1) & ~3u));
            lo = w[0];
            hi = w[1];
        }
        const u32 sh = ((q - 1) & 3u) * 8u;
        Wm = __funnelshift_r(lo, hi, sh);
        W = __funnelshift_rc(lo, hi, sh + 8u);
    }

    // make last step's prefetch visible, then prefetch one missing position of [lo, lo+64)
    __device__ __forceinline__ void refill(u32 lo) {
        if (ptag != kNoTag) {
            if ((ptag >> 5) & 1) { tag1 = ptag; a1 = pa; b1 = pb; }
            else { tag0 = ptag; a0 = pa; b0 = pb; }
            ptag = kNoTag;
        }
        const u32 q0 = lo + ((lane - lo) & 31u), q1 = q0 + 32;
        const u32 s0 = (q0 >> 5) & 1;
        const u32 t0 = s0 ? tag1 : tag0, t1 = s0 ? tag0 : tag1;
        const u32 need = (t0 != q0) ? q0 : ((t1 != q1) ? q1 : kNoTag);
        if (need <= n) {  // kNoTag > n
            const u32* w = reinterpret_cast<const u32*>(F + ((need - 1) & ~3u));
            ptag = need;
            pa = w[0];
            pb = w[1];
        }
    }

    // One scan round over 32 probes.  Lane's probe: index i (0..31 in probe order), position p,
    // successor position pn, word W at p.  r = lane of probe 0 (probe order = lane order rotated).
    // Returns 1 = hit (ip/cand set), 2 = bail to remainder, 0 = no hit (all 32 inserted).
    template <bool kStride1>
    __device__ __forceinline__ int scan_round(u32 i, u32 r, u32 p, u32 pn, u32 W, u32& ip, u32& cand) {
        const bool valid = (int)pn <= lim;  // src/internal.jl:175
        const u32 H = hash(W);
        const u32 mp = rotr32(__match_any_sync(kFullMask, valid ? H : (0x80000000u | lane)), r);
        u32 c = valid ? (u32)T[H] : 0u;
        const u32 prior = mp & ((1u << i) - 1u);
        // forwarding: the latest earlier probe with the same hash is what the table would hold (:191)
        const u32 j = 31u - (u32)__clz((int)(prior | 1u));
        if (kStride1) {
            if (prior) c = p - i + j;
        } else {
            const u32 fp = __shfl_sync(kFullMask, p, prior ? ((j + r) & 31u) : lane);
            if (prior) c = fp;
        }
        const bool eq = valid && (lds32u(F, c) == W);  // :193
        const u32 hitm = rotr32(__ballot_sync(kFullMask, eq), r);
        const u32 invm = rotr32(__ballot_sync(kFullMask, !valid), r);
        const u32 fh = hitm ? (u32)__ffs((int)hitm) - 1u : 32u;
        const u32 fi = invm ? (u32)__ffs((int)invm) - 1u : 32u;
        if (fh >= fi && fi < 32) return 2;
        const u32 last = (fh < 32) ? fh : 31u;
        // commit inserts of probes 0..last; on equal hashes the later probe wins (:191)
        if (i <= last && (mp & ~mask_le(i) & mask_le(last)) == 0) T[H] = (u16)p;
        __syncwarp();
        if (fh < 32) {
            const u32 src = (fh + r) & 31u;
            ip = __shfl_sync(kFullMask, p, src);
            cand = __shfl_sync(kFullMask, c, src);
            return 1;
        }
        return 0;
    }

    __device__ __forceinline__ void run() {
        head = 0;
        tail_cache = 0;
        tag0 = tag1 = ptag = kNoTag;
        a0 = b0 = a1 = b1 = pa = pb = 0;
        lim = (int)n - 16;  // ip_limit, src/internal.jl:131
        u32 ip = 0, lit_from = 0;
        if (n >= kInputMargin) {
            bool finished = false;
            refill(1);
            while (!finished) {
                // ---------------- scan, src/internal.jl:162-194
                const u32 s = ip + 1;
                u32 cand = 0;
                int res;
                {
                    const u32 i = (lane - s) & 31u, q = s + i;
                    u32 W = 0, Wm;
                    if ((int)(q + 1) <= lim) lookup(q, W, Wm);
                    res = scan_round<true>(i, s & 31u, q, q + 1, W, ip, cand);
                }
                for (u32 base = 32; res == 0; base += 32) {
                    const u32 p = s + PO[base + lane], pn = s + PO[base + lane + 1];
                    const u32 W = ((int)pn <= lim) ? lds32u(F, p) : 0u;
                    res = scan_round<false>(lane, 0, p, pn, W, ip, cand);
                }
                if (res == 2) break;
                refill(ip + 1);
                // ---------------- copy chain, src/internal.jl:211-239
                for (;;) {
                    // speculative post-copy probes for every end position e in [ip+4, ip+36)
                    const u32 e = ip + 4 + ((lane - (ip + 4)) & 31u);
                    const bool ok = (int)e < lim;
                    u32 We = 0, Wme = 0;
                    if (ok) lookup(e, We, Wme);
                    const u32 He = hash(We), Hme = hash(Wme);
                    const u32 ce = (Hme == He) ? (e - 1) : (ok ? (u32)T[He] : 0u);
                    const bool eqe = ok && (lds32u(F, ce) == We);
                    const u32 hitm = __ballot_sync(kFullMask, eqe);
                    // the real thing
                    const u32 M = 4 + warp_match_length(F, cand + 4, ip + 4, n, lane);  // :216
                    push(lit_from, ip - lit_from, ip - cand, M);                       // :200,:217
                    ip += M;
                    lit_from = ip;
                    if ((int)ip >= lim) { finished = true; break; }  // :222
                    bool same;
                    u32 c2;
                    if (M - 4 < 32) {
                        const u32 owner = ip & 31u;
                        c2 = __shfl_sync(kFullMask, ce, owner);
                        same = (hitm >> owner) & 1u;
                        if (lane == owner) {
                            T[Hme] = (u16)(ip - 1);  // :233
                            T[He] = (u16)ip;         // :235
                        }
                    } else {
                        const u32 w = lds32u(F, ip);
                        const u32 hp = hash(lds32u(F, ip - 1)), hc = hash(w);
                        c2 = (hp == hc) ? (ip - 1) : (u32)T[hc];  // :233-234
                        __syncwarp();
                        if (lane == 0) {
                            T[hp] = (u16)(ip - 1);
                            T[hc] = (u16)ip;
                        }
                        same = (lds32u(F, c2) == w);  // :238
                    }
                    __syncwarp();
                    refill(ip + 1);
                    if (!same) break;
                    cand = c2;
                }
            }
        }
        if (lit_from < n) push(lit_from, n - lit_from, 0, 0);  // :242-248
        push(0, 0, kRecDone, 0);
    }
};

__device__ __forceinline__ u32 copy_tag_bytes(u32 off, u32 M) {  // size of emit_copy!, :306-329
    if (M == 0) return 0;
    u32 bytes = 0;
    if (M >= 12) {
        if (M >= 68) {
            const u32 k = (M - 4) / 64;
            bytes = 3 * k;
            M -= 64 * k;
        }
        if (M > 64) {
            bytes += 3;
            M -= 60;
        }
    }
    return bytes + ((M < 12 && off < 2048) ? 2u : 3u);
}

__device__ __forceinline__ u32 put_copy_op(u8* out, u32 p, u32 off, u32 len) {  // :289-304
    if (len < 12 && off < 2048) {
        out[p] = (u8)(1 + ((len - 4) << 2) + ((off >> 3) & 0xe0));
        out[p + 1] = (u8)off;
        return p + 2;
    }
    const u32 u = 2 + ((len - 1) << 2) + (off << 8);
    out[p] = (u8)u;
    out[p + 1] = (u8)(u >> 8);
    out[p + 2] = (u8)(u >> 16);
    return p + 3;
}

// Emitter warp: ring records -> tag bytes in the fragment's scratch slot.  Returns bytes written.
__device__ __forceinline__ u32 emit_records(const u8* F, u8* out, const uint4* ring, u32* tail_ptr,
                                            const u32 lane) {
    u32 tail = 0, op = 0;
    bool finished = false;
    while (!finished) {
        const u32 idx = tail + lane;
        const uint4 r = lds_volatile_v4(&ring[idx % kRing]);
        const bool valid = (r.x >> 31) == ((idx / kRing) & 1u);
        const u32 vm = __ballot_sync(kFullMask, valid);
        u32 cnt = (vm == kFullMask) ? 32u : (u32)__ffs((int)~vm) - 1u;
        if (cnt == 0) {
            __nanosleep(40);
            continue;
        }
        u32 consumed = cnt;
        const u32 dm = __ballot_sync(kFullMask, lane < cnt && r.z == kRecDone);
        if (dm) {
            finished = true;
            cnt = (u32)__ffs((int)dm) - 1u;
            consumed = cnt + 1;
        }
        const bool mine = lane < cnt;
        const u32 lf = r.x & 0x7fffffffu, ll = mine ? r.y : 0u, off = r.z, M = mine ? r.w : 0u;
        // src/internal.jl:271-283: header bytes of the literal
        const u32 lh = (ll == 0) ? 0u : (ll < 60 ? 1u : ((ll - 1) <= 0xffu ? 2u : ((ll - 1) <= 0xffffu ? 3u : 4u)));
        const u32 sz = lh + ll + copy_tag_bytes(off, M);
        u32 incl = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= (u32)d) incl += t;
        }
        const u32 pos = op + incl - sz;
        op += __shfl_sync(kFullMask, incl, 31);
        if (ll) {
            const u32 nm1 = ll - 1;
            if (ll < 60) {
                out[pos] = (u8)(nm1 << 2);
            } else {
                out[pos] = (u8)((59 + (lh - 1)) << 2);
                out[pos + 1] = (u8)nm1;
                if (lh > 2) out[pos + 2] = (u8)(nm1 >> 8);
                if (lh > 3) out[pos + 3] = (u8)(nm1 >> 16);
            }
            if (ll <= kShortLiteral) {
                for (u32 k = 0; k < ll; k++) out[pos + lh + k] = F[lf + k];
            }
        }
        // long literals: the whole warp copies them one after the other
        u32 lm = __ballot_sync(kFullMask, ll > kShortLiteral);
        while (lm) {
            const u32 j = (u32)__ffs((int)lm) - 1u;
            lm &= lm - 1;
            const u32 src = __shfl_sync(kFullMask, lf, j);
            const u32 len = __shfl_sync(kFullMask, ll, j);
            const u32 dst = __shfl_sync(kFullMask, pos + lh, j);
            for (u32 k = lane; k < len; k += 32) out[dst + k] = F[src + k];
        }
        if (M) {  // src/internal.jl:306-329
            u32 p = pos + lh + ll, len = M;
            if (len >= 12) {
                while (len >= 68) {
                    p = put_copy_op(out, p, off, 64);
                    len -= 64;
                }
                if (len > 64) {
                    p = put_copy_op(out, p, off, 60);
                    len -= 60;
                }
            }
            put_copy_op(out, p, off, len);
        }
        tail += consumed;
        if (lane == 0) sts_volatile_u32(tail_ptr, tail);
    }
    return op;
}

__global__ void __launch_bounds__(64)
k_compress_fragments(const u8* __restrict__ g_in, u64 shard_len, u32 shift, u8* __restrict__ scratch,
                     u32* __restrict__ frag_sizes) {
    extern __shared__ __align__(128) u8 smem[];
    u8* F = smem;
    u16* T = reinterpret_cast<u16*>(smem + kBlockSize + 64);
    u32* PO = reinterpret_cast<u32*>(smem + kBlockSize + 64 + kMaxTableEntries * 2);
    uint4* ring = reinterpret_cast<uint4*>(PO + kPoEntries);
    u64* bar = reinterpret_cast<u64*>(ring + kRing);
    u32* tail_ptr = reinterpret_cast<u32*>(bar + 1);

    const u32 lane = lane_id();
    const u32 warp = threadIdx.x >> 5;
    const u32 frag = blockIdx.x;
    const u64 start = (u64)frag * kBlockSize;
    const u32 n = (u32)((shard_len - start < kBlockSize) ? (shard_len - start) : kBlockSize);

    if (warp == 0) {
        load_fragment(F, bar, g_in + start, n, lane, 0);
        if (lane < 32) F[n + 32 + lane] = 0;  // load_fragment zeroes 32 pad bytes; the pad here is 64
    } else {
        reset_table(T, shift, lane);
        for (u32 i = lane; i < kRing; i += 32) ring[i] = make_uint4(0x80000000u, 0, 0, 0);
        if (lane == 0) {
            *tail_ptr = 0;
            // probe offsets of the skip heuristic: skip starts at 32, step = skip >> 5 (:162-172)
            u32 skip = 32, off = 0;
            PO[0] = 0;
            for (u32 i = 1; i < kPoEntries; i++) {
                const u32 b = skip >> 5;
                skip += b;
                off = (off + b > 0x100000u) ? 0x100000u : off + b;
                PO[i] = off;
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        Decider d;
        d.F = F; d.T = T; d.PO = PO; d.ring = ring; d.tail_ptr = tail_ptr;
        d.n = n; d.shift = shift; d.lane = lane;
        d.run();
    } else {
        const u32 total = emit_records(F, scratch + (u64)frag * kSlotStride, ring, tail_ptr, lane);
        if (lane == 0) frag_sizes[frag] = total;
    }
}

// K1: one CTA (one warp) per fragment of the shard.
//   g_in        : first byte of the shard (a multiple of 65536 inside the stream)
//   shard_len   : bytes in the shard; fragment f covers [f*65536, min((f+1)*65536, shard_len))
//   shift       : 32 - log2(table entries), entries derived from the TOTAL stream length
//   scratch     : per-fragment output slots of kSlotStride bytes
//   frag_sizes  : compressed size of each fragment
__global__ void __launch_bounds__(32)
k_compress_fragments_serial(const u8* __restrict__ g_in, u64 shard_len, u32 shift,
                            u8* __restrict__ scratch, u32* __restrict__ frag_sizes) {
    extern __shared__ __align__(128) u8 smem[];
    u8* F = smem;
    u16* T = reinterpret_cast<u16*>(smem + kBlockSize + kFragPad);
    u64* bar = reinterpret_cast<u64*>(smem + kBlockSize + kFragPad + kMaxTableEntries * 2);

    const u32 lane = lane_id();
    const u32 frag = blockIdx.x;
    const u64 start = (u64)frag * kBlockSize;
    const u32 n = (u32)((shard_len - start < kBlockSize) ? (shard_len - start) : kBlockSize);

    load_fragment(F, bar, g_in + start, n, lane, 0);
    reset_table(T, shift, lane);
    FragEmitter em{scratch + (u64)frag * kSlotStride, 0, lane};
    compress_fragment_serial(F, T, n, shift, em);
    if (lane == 0) frag_sizes[frag] = em.op;
}

#endif  // SB200_EXPERIMENTS

// K1b: batched pages -- one CTA (one warp) per independent stream (src/Snappy.jl:20-36 per page:
// own varint header, table sized from the page length).  Fragments of a page are compressed one
// after the other straight into the page's output slot, so no compaction pass is needed.
// Shared memory: frag_cap + kFragPad bytes of fragment, then table_cap u16 entries, then mbarrier.
// kLib: libsnappy's rules (option `rules`, 1 or 2), table sized per fragment up to table_cap.
template <bool kLib = false>
__global__ void __launch_bounds__(32)
k_compress_pages(const u8* __restrict__ g_in, const u64* __restrict__ in_off,
                 const u32* __restrict__ in_size, u8* __restrict__ g_out,
                 const u64* __restrict__ out_off, u32* __restrict__ out_size, u32 frag_cap,
                 u32 table_cap, u32 lib_rules = 0) {
    extern __shared__ __align__(128) u8 smem[];
    u8* F = smem;
    u16* T = reinterpret_cast<u16*>(smem + frag_cap + kFragPad);
    u64* bar = reinterpret_cast<u64*>(smem + frag_cap + kFragPad + table_cap * 2);

    const u32 lane = lane_id();
    const u32 pg = blockIdx.x;
    const u8* pin = g_in + in_off[pg];
    const u32 total = in_size[pg];
    u32 entries = 256;  // alloc_hashtable, src/internal.jl:107-113
    while (entries < kMaxTableEntries && entries < total) entries <<= 1;
    const u32 shift = 32 - (31 - __clz(entries));

    FragEmitter em{g_out + out_off[pg], 0, lane};
    if (kLib) em.lit_short = 61;
    {   // varint header, src/varint.jl:46-69
        u32 v = total, k = 0;
        while (v >= 0x80) {
            if (lane == 0) em.out[k] = (u8)(v | 0x80);
            v >>= 7;
            k++;
        }
        if (lane == 0) em.out[k] = (u8)v;
        em.op = k + 1;
    }
    u32 phase = 0;
    for (u32 s = 0; s < total; s += kBlockSize) {
        const u32 n = (total - s < kBlockSize) ? (total - s) : kBlockSize;
        __syncwarp();
        phase = load_fragment(F, bar, pin + s, n, lane, phase);
        if (kLib) {  // GetHashTable: sized from this fragment's length
            u32 e = 256;
            while (e < table_cap && e < n) e <<= 1;
            uint4* t4 = reinterpret_cast<uint4*>(T);
            for (u32 i = lane; i < e / 8; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            compress_fragment_serial<true>(F, T, n, lib_rules == 2u ? 17u : (u32)__clz((int)e) + 1u, em, e - 1u);
        } else {
            reset_table(T, shift, lane);
            compress_fragment_serial(F, T, n, shift, em);
        }
    }
    if (lane == 0) out_size[pg] = em.op;
}

// K2: exclusive scan of the fragment sizes (single CTA; nfrag <= 65536), plus the side index.
//   offsets[f]   : byte offset of fragment f behind `base` (base = varint header length, or 0)
//   offsets[nfrag] = base + total
__global__ void __launch_bounds__(1024)
k_scan_sizes(const u32* __restrict__ sizes, u32 nfrag, u64 base, u64* __restrict__ offsets,
             u64* running = nullptr) {
    __shared__ u64 warp_excl[32];
    __shared__ u64 carry_s;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = base + (running ? *running : 0ull);  // running: total carried from chunk to chunk
    __syncthreads();
    constexpr int kItems = 4;  // per thread and pass (16 was tried for the ~470 K chunk sizes of the index-free parse: the
                               // 64-byte stride between lanes costs more than the saved passes, 8.22 vs 7.99 ms for config 3)
    for (u32 blk = 0; blk < nfrag; blk += 1024 * kItems) {
        const u32 i0 = blk + tid * kItems;
        u32 v[kItems];
        u64 s = 0;
#pragma unroll
        for (int k = 0; k < kItems; k++) {
            v[k] = (i0 + k < nfrag) ? sizes[i0 + k] : 0;
            s += v[k];
        }
        u64 incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 t = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= (u32)d) incl += t;
        }
        if (lane == 31) warp_excl[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            u64 w = warp_excl[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 t = __shfl_up_sync(kFullMask, wi, d);
                if (lane >= (u32)d) wi += t;
            }
            warp_excl[lane] = wi - w;
        }
        __syncthreads();
        u64 ex = carry_s + warp_excl[wid] + (incl - s);
#pragma unroll
        for (int k = 0; k < kItems; k++) {
            if (i0 + k < nfrag) offsets[i0 + k] = ex;
            ex += v[k];
        }
        __syncthreads();
        if (tid == 1023) carry_s = ex;  // thread 1023 ends at carry + this block's total
        __syncthreads();
    }
    if (tid == 0) {
        offsets[nfrag] = carry_s;
        if (running) *running = carry_s;
    }
}

// K2s: the scan of one chunk of the streamed host-buffer path.  Waits until all `need` fragments of
// the chunk have been compressed (done counter of k_compress_window), then scans <= 1024 sizes behind
// the running stream length.  256 threads so that it fits next to the persistent compress CTAs.
__global__ void __launch_bounds__(256)
k_scan_chunk(const u32* __restrict__ sizes, u32 nf, u64* __restrict__ offsets, u64* __restrict__ running,
             const u32* done, u32 need, u64* host_total) {
    __shared__ u64 warp_excl[8];
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        while (*reinterpret_cast<const volatile u32*>(done) < need) __nanosleep(1000);
        __threadfence();
    }
    __syncthreads();
    u64 v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k] = (tid * 4 + k < nf) ? __ldcg(sizes + tid * 4 + k) : 0;
        s += v[k];
    }
    u64 incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= (u32)d) incl += t;
    }
    if (lane == 31) warp_excl[wid] = incl;
    __syncthreads();
    u64 before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        if ((u32)w < wid) before += warp_excl[w];
        total += warp_excl[w];
    }
    const u64 base = *running;
    u64 ex = base + before + (incl - s);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (tid * 4 + k < nf) offsets[tid * 4 + k] = ex;
        ex += v[k];
    }
    __syncthreads();
    if (tid == 0) {
        offsets[nf] = base + total;
        *running = base + total;
        // the stream length so far goes straight into pinned host memory (a device-to-host copy would
        // queue behind the output copies on the copy engine)
        *host_total = base + total;
        __threadfence_system();
    }
}


// K3: concatenate the per-fragment scratch slots into the contiguous stream.  One CTA per
// fragment; destination-aligned 16-byte stores, source read as aligned words + funnel shift.
__global__ void __launch_bounds__(256)
k_compact(const u8* __restrict__ scratch, const u32* __restrict__ sizes,
          const u64* __restrict__ offsets, u8* __restrict__ out) {
    const u32 frag = blockIdx.x;
    const u32 c = sizes[frag];
    const u8* src = scratch + (u64)frag * kSlotStride;
    u8* dst = out + offsets[frag];
    u32 head = (u32)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    if (head > c) head = c;
    for (u32 i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    const u32 nvec = (c - head) >> 4;
    const u32 sh = (head & 3) * 8;
    const u32* sw = reinterpret_cast<const u32*>(src + (head & ~3u));
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (u32 j = threadIdx.x; j < nvec; j += blockDim.x) {
        const u32* p = sw + 4 * j;
        u32 w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = p[4];
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh);
        v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh);
        v.w = __funnelshift_r(w3, w4, sh);
        d4[j] = v;
    }
    for (u32 i = head + (nvec << 4) + threadIdx.x; i < c; i += blockDim.x) dst[i] = src[i];
}

}  // namespace sb200
