// compress_chain.cuh -- K1 (default): lane-speculative fragment compressor, one warp per fragment.
//
// What bounds this kernel (measured, profiles/): the reference's algorithm is a serial decision
// chain of ~4-8 thousand dependent steps per 64 KiB fragment (probe -> candidate -> match length
// -> next probe, src/internal.jl:162-239), so a fragment advances at the speed of ONE warp's
// dependent instruction stream, and throughput = (independent fragments in flight per SM) /
// (cycles per step).  The only per-fragment state that needs low-latency random read/write is the
// 32 KiB u16 hash table; the fragment bytes are read-only.  So the table lives in shared memory and
// the fragment is read in place through L1/L2 (the active working set, resident warps x 64 KiB,
// sits in the 126 MB L2), which lets 6 fragments run per SM instead of the 2 that fit when the
// fragment is staged in shared memory next to its table.  Warps are persistent and pull fragments
// from a global counter, so extra warps whose table lives in global memory (L2) can join in.
//
// The decisions are exactly the reference's; the lanes only evaluate them ahead of time:
//   * scan (src/internal.jl:167-194): lane i takes the i-th probe position of the skip sequence; a
//     probe whose hash equals an earlier lane's sees that lane's position (== the table insert the
//     reference would have made, :191); the first hit wins; only lanes up to it commit inserts.
//   * copy chain (:211-239): while the match length is measured (32 bytes per ballot), lane l
//     pre-evaluates the post-copy probe (:228-238) for end position ip+4+l; the lane of the real
//     end position supplies candidate and verdict.
//   * emission (:252-329): (literal, copy) records are parked one per lane and turned into tag
//     bytes 32 at a time with a warp prefix sum for the output positions.
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr u32 kChainPoEntries = 352;  // probe offsets of the skip heuristic (:162-172)
constexpr u32 kPrefetchLanes = 8;       // post-copy candidates prefetched into L1 (ends ip+4 .. ip+11)
constexpr u32 kFirstProbes = 8;        // lanes in the first round of a scan
constexpr u32 kStreamAhead = 2048;     // ip-side bytes are pulled into L2 this far ahead
constexpr u32 kTailPad = 256;         // zero bytes behind the padded copy of the shard's last fragment

// unaligned little-endian 32-bit load through the read-only path (fastmemory.jl:4 load32u);
// touches the two aligned words around p, i.e. up to 7 bytes past p
__device__ __forceinline__ u32 ldg32u(const u8* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const u32* w = reinterpret_cast<const u32*>(a & ~(uintptr_t)3);
    return __funnelshift_r(__ldg(w), __ldg(w + 1), (u32)a << 3);
}

// probe offsets of the skip heuristic (skip starts at 32, step = skip >> 5, :162-172); filled once
// by k_init_probe_offsets.  Only scan rounds past the first 32 probes read it (incompressible data).
__device__ u32 g_probe_offsets[kChainPoEntries];
#ifdef SB200_EXPERIMENTS
__device__ u32 g_dbg_skip_emit = 0;  // upper-bound experiments: what would free emission buy
#endif

__global__ void k_init_probe_offsets() {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u32 skip = 32, off = 0;
    g_probe_offsets[0] = 0;
    for (u32 i = 1; i < kChainPoEntries; i++) {
        const u32 b = skip >> 5;
        skip += b;
        off = (off + b > 0x100000u) ? 0x100000u : off + b;
        g_probe_offsets[i] = off;
    }
}

// kLib: libsnappy emission rules instead of Snappy.jl's (option `rules`): ip_limit =
// n - 15, a 60-byte literal keeps the one-byte header, the table is sized per fragment and the bucket of a hash
// is ((w * mul) >> shift) & hmask (rules = 2: Google snappy >= 1.1.9, shift 17 and up to 32768 buckets).
// Output bytes of a fragment are written once and read once, much later, by the compaction kernel: with
// SB200_OUT_CS they are stored with the streaming (evict-first) policy, so that they do not push the fragment bytes
// the global-table warps gather their far candidates from out of L2.
#if defined(SB200_OUT_CS) && !defined(SB200_CPU_EMU)
#define SB200_ST8(p, v) __stcs((p), (u8)(v))
#else
#define SB200_ST8(p, v) (*(p) = (u8)(v))
#endif

// kSmemTable: 1 = the table lives in shared memory, 0 = in global memory (L2), 2 = either, decided per warp at run time
// (dyn_smem): ONE copy of the round's code serves both kinds of warp of the merged kernel (instruction cache).
#ifndef SB200_FLUSH_NOINLINE
#define SB200_FLUSH_NOINLINE 0
#endif
#ifndef SB200_PIN_OUT
#define SB200_PIN_OUT 0
#endif
template <int kSmemTable, bool kLib = false>
struct Chain {
    static constexpr u32 kLitShort = kLib ? 61u : 60u;   // literals below this take the one-byte header (:271)
    static constexpr int kLimMargin = kLib ? 15 : 16;    // ip_limit = n - margin (:131)
    u32 hmask;       // kLib only
    const u8* F;     // fragment bytes (global, >= 64 readable bytes past n)
    u16* T;          // hash table, position per hash, 0 == empty (global variant)
    u32 Ts;          // shared-space address of the table (shared variant)
    u8* out;         // scratch slot of this fragment
    u32 n, shift, lane, op, nrec, spec;  // spec: end positions pre-probed per copy (<= 32)
    int lim;
    u32 r_lit, r_cpy;  // lane k parks record k: (lit_from | ip << 16), (cand | M << 16)
    u64 pol;           // L2 cache policy of the global-table accesses (evict_last), make_policy()
    bool dyn_smem;     // kSmemTable == 2 only
    __device__ __forceinline__ bool in_smem() const { return kSmemTable == 2 ? dyn_smem : kSmemTable == 1; }

    __device__ __forceinline__ u32 hash(u32 w) const {
        return kLib ? (((w * kHashMul) >> shift) & hmask) : ((w * kHashMul) >> shift);
    }
#ifdef SB200_CPU_EMU
    __device__ __forceinline__ void make_policy() { pol = 0; }
    __device__ __forceinline__ u32 tget(u32 h) const {
        return in_smem() ? *reinterpret_cast<const u16*>(smem + Ts + 2u * h) : T[h];
    }
    __device__ __forceinline__ void tput(u32 h, u32 pos) const {
        if (in_smem()) *reinterpret_cast<u16*>(smem + Ts + 2u * h) = (u16)pos;
        else T[h] = (u16)pos;
    }
#else
    __device__ __forceinline__ u32 tget(u32 h) const {
        if (in_smem()) {
            u16 v;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(Ts + 2u * h) : "memory");
            return v;
        }
        // global tables: keep their lines in L2 ahead of everything that streams through it (evict_last):
        // -4 % kernel time, they are the randomly re-read state
        // (the policy is made once per fragment, make_policy(): as part of every access `createpolicy` cost six
        // uniform-datapath instructions, three times per round)
        u16 v;
        asm volatile("ld.global.cg.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(T + h), "l"(pol) : "memory");
        return v;
    }
    __device__ __forceinline__ void tput(u32 h, u32 pos) const {
        if (in_smem()) asm volatile("st.shared.u16 [%0], %1;" ::"r"(Ts + 2u * h), "h"((u16)pos) : "memory");
        else asm volatile("st.global.cg.L2::cache_hint.u16 [%0], %1, %2;" ::"l"(T + h), "h"((u16)pos), "l"(pol) : "memory");
    }
    __device__ __forceinline__ void make_policy() {
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    }
#endif

    // ---- emission ---------------------------------------------------------------------------
    static __device__ __forceinline__ u32 copy_bytes(u32 off, u32 M) {  // size of emit_copy!, :306-329
        if (M == 0) return 0;
        u32 bytes = 0;
        if (M >= 12) {
            if (M >= 68) {
                const u32 k = (M - 4) >> 6;
                bytes = 3 * k;
                M -= k << 6;
            }
            if (M > 64) {
                bytes += 3;
                M -= 60;
            }
        }
        return bytes + ((M < 12 && off < 2048) ? 2u : 3u);
    }
    static __device__ __forceinline__ u32 put_op(u8* o, u32 p, u32 off, u32 len) {  // :289-304
        if (len < 12 && off < 2048) {
            SB200_ST8(o + p, 1 + ((len - 4) << 2) + ((off >> 3) & 0xe0));
            SB200_ST8(o + p + 1, off);
            return p + 2;
        }
        const u32 u = 2 + ((len - 1) << 2) + (off << 8);
        SB200_ST8(o + p, u);
        SB200_ST8(o + p + 1, u >> 8);
        SB200_ST8(o + p + 2, u >> 16);
        return p + 3;
    }

#if SB200_FLUSH_NOINLINE && !defined(SB200_CPU_EMU)
    __device__ __noinline__ void flush() {
#else
    __device__ __forceinline__ void flush() {
#endif
        {   // pull the next few KiB of the fragment into L2 ahead of the ip-side loads (32 x 128 B)
            const u32 ahead = (r_lit >> 16) + kStreamAhead + lane * 128u;  // from this lane's ip
#ifndef SB200_CPU_EMU
            if (ahead < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(F + ahead));
#endif
        }
        const bool mine = lane < nrec;
        const u32 lf = r_lit & 0xffffu, ll = mine ? ((r_lit >> 16) - lf) : 0u;
        const u32 off = (r_lit >> 16) - (r_cpy & 0xffffu), M = mine ? (r_cpy >> 16) : 0u;
        // :271-283 header bytes of the literal (a 60-byte literal already takes the long form)
        const u32 lh = (ll == 0) ? 0u : (ll < kLitShort ? 1u : ((ll - 1) <= 0xffu ? 2u : ((ll - 1) <= 0xffffu ? 3u : 4u)));
        const u32 sz = lh + ll + copy_bytes(off, M);
        u32 incl = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= (u32)d) incl += t;
        }
        const u32 pos = op + incl - sz;
        op += __shfl_sync(kFullMask, incl, 31);
#if SB200_PIN_OUT && !defined(SB200_CPU_EMU)
        // build switch, off: keep the slot's base address in a register pair; ptxas otherwise rebuilds it (constant-bank
        // load, 64-bit multiply-add with the fragment number) in front of every group of byte stores below.  328 fewer
        // SASS instructions, 92 registers, and no faster: 12.28 vs 12.21 ms (profiles/r02zi_sweep_pin_out.txt)
        u8* out = this->out;
        asm volatile("" : "+l"(out));
#endif
#ifdef SB200_EXPERIMENTS
        if (g_dbg_skip_emit) {  // measurement only (option dbg_skip_emit): sizes stay right, bytes are not written
            nrec = 0;
            return;
        }
#endif
        if (ll) {
            const u32 nm1 = ll - 1;
            if (ll < kLitShort) {
                SB200_ST8(out + pos, nm1 << 2);
            } else {
                SB200_ST8(out + pos, (59 + (lh - 1)) << 2);
                SB200_ST8(out + pos + 1, nm1);
                if (lh > 2) SB200_ST8(out + pos + 2, nm1 >> 8);
                if (lh > 3) SB200_ST8(out + pos + 3, nm1 >> 16);
            }
            if (ll <= 16) {
                // <= 16 literal bytes through the aligned words that hold them (only words with a wanted byte are
                // read), then one predicated byte store per position: no per-lane loop
                const uintptr_t sa = reinterpret_cast<uintptr_t>(F + lf);
                const u32* w = reinterpret_cast<const u32*>(sa & ~(uintptr_t)3);
                const u32 sh = (u32)sa << 3;
                const u32 last = (((u32)sa & 3u) + ll - 1u) >> 2;  // index of the last word needed (0..4)
                const u32 w0 = __ldg(w), w1 = last >= 1 ? __ldg(w + 1) : 0u, w2 = last >= 2 ? __ldg(w + 2) : 0u,
                          w3 = last >= 3 ? __ldg(w + 3) : 0u, w4 = last >= 4 ? __ldg(w + 4) : 0u;
                const u32 b0 = __funnelshift_r(w0, w1, sh), b1 = __funnelshift_r(w1, w2, sh),
                          b2 = __funnelshift_r(w2, w3, sh), b3 = __funnelshift_r(w3, w4, sh);
                u8* d = out + pos + lh;
#pragma unroll
                for (u32 i = 0; i < 16; i++) {
                    const u32 word = i < 4 ? b0 : (i < 8 ? b1 : (i < 12 ? b2 : b3));
                    if (i < ll) SB200_ST8(d + i, word >> (8 * (i & 3)));
                }
            }
        }
        u32 lm = __ballot_sync(kFullMask, ll > 16);  // long literals: whole warp, one after the other
        while (lm) {
            const u32 j = (u32)__ffs((int)lm) - 1u;
            lm &= lm - 1;
            const u32 src = __shfl_sync(kFullMask, lf, j);
            const u32 len = __shfl_sync(kFullMask, ll, j);
            const u32 dst = __shfl_sync(kFullMask, pos + lh, j);
            for (u32 k = lane; k < len; k += 32) SB200_ST8(out + dst + k, __ldg(F + src + k));
        }
        if (M) {  // :306-329
            u32 p = pos + lh + ll, len = M;
            if (len >= 12) {
                while (len >= 68) {
                    p = put_op(out, p, off, 64);
                    len -= 64;
                }
                if (len > 64) {
                    p = put_op(out, p, off, 60);
                    len -= 60;
                }
            }
            put_op(out, p, off, len);
        }
        nrec = 0;
    }

    // literal [from, ip) followed by a copy of M bytes from cand (all < 65536)
    __device__ __forceinline__ void keep(u32 from, u32 ip, u32 cand, u32 M) {
        if (lane == nrec) {
            r_lit = from | (ip << 16);
            r_cpy = cand | (M << 16);
        }
        if (++nrec == 32) flush();
    }

    // ---- one scan round over 32 probes (lane i = i-th probe of the round) ------------------------
    // returns 1 = hit (ip/cand set), 2 = bail to the remainder, 0 = no hit (all 32 inserted)
    // `act`: lanes that take part (the first round of a scan uses only kFirstProbes lanes: most scans
    // hit within a few probes, and every probing lane costs a candidate-side cache line)
    __device__ __forceinline__ int scan_round(u32 p, u32 pn, bool act, u32& ip, u32& cand) {
        const bool valid = act && (int)pn <= lim;  // :175 (checked before probing p)
        const u32 W = ldg32u(F + (valid ? p : 0u));
        const u32 H = hash(W);
        const u32 mp = __match_any_sync(kFullMask, valid ? H : (0x80000000u | lane));
        u32 c = tget(valid ? H : 0u);
        // forwarding: the latest earlier probe with the same hash is what the table would hold (:191)
        const u32 prior = mp & ((1u << lane) - 1u);
        const u32 j = 31u - (u32)__clz((int)(prior | 1u));
        const u32 fp = __shfl_sync(kFullMask, p, j);
        if (prior) c = fp;
        const bool eq = valid && (ldg32u(F + c) == W);  // :193
        const u32 hitm = __ballot_sync(kFullMask, eq);
        const u32 invm = __ballot_sync(kFullMask, act && !valid);
        const u32 fh = hitm ? (u32)__ffs((int)hitm) - 1u : 32u;
        const u32 fi = invm ? (u32)__ffs((int)invm) - 1u : 32u;
        if (fh >= fi && fi < 32) return 2;
        const u32 last = (fh < 32) ? fh : 31u;
        const u32 upto = (last >= 31) ? kFullMask : ((2u << last) - 1u);
        const u32 after = (lane >= 31) ? 0u : ~((2u << lane) - 1u);
        // commit inserts of probes 0..last; on equal hashes the later probe wins (:191)
        if (valid && lane <= last && (mp & after & upto) == 0) tput(H, p);
        __syncwarp();
        if (fh < 32) {
            ip = __shfl_sync(kFullMask, p, fh);
            cand = __shfl_sync(kFullMask, c, fh);
            return 1;
        }
        return 0;
    }

    // ---- :242-248: pending records, then the remainder literal (up to the whole fragment: emitted directly)
    __device__ __forceinline__ void finish(u32 lit_from) {
        if (nrec) flush();
        if (lit_from < n) {
            const u32 ll = n - lit_from, nm1 = ll - 1;
            const u32 lh = ll < kLitShort ? 1u : (nm1 <= 0xffu ? 2u : (nm1 <= 0xffffu ? 3u : 4u));
            if (lane == 0) {
                if (ll < kLitShort) {
                    out[op] = (u8)(nm1 << 2);
                } else {
                    out[op] = (u8)((59 + (lh - 1)) << 2);
                    out[op + 1] = (u8)nm1;
                    if (lh > 2) out[op + 2] = (u8)(nm1 >> 8);
                    if (lh > 3) out[op + 3] = (u8)(nm1 >> 16);
                }
            }
            __syncwarp();
            warp_copy_forward<true>(out + op + lh, F + lit_from, ll, lane);  // 16-byte stores when long
            op += lh + ll;
        }
    }

    // ---- the fragment --------------------------------------------------------------------------
    __device__ __forceinline__ void run() {
        // keep these in registers: without the barrier ptxas re-derives them (64-bit min, S2R,
        // shared-window base) inside the copy loop, ~20 instructions per step
        asm volatile("" : "+r"(n), "+r"(Ts), "+r"(shift));
        op = 0;
        nrec = 0;
        r_lit = r_cpy = 0;
        lim = (int)n - kLimMargin;  // ip_limit, :131
        asm volatile("" : "+r"(lim));
        u32 ip = 0, lit_from = 0;
        if (n >= kInputMargin) {
            bool finished = false;
            while (!finished) {
                // ---------------- scan, :162-194
                const u32 s = ip + 1;
                u32 cand = 0;
                // the first 32 probes have stride 1; probe k >= 32 sits at s + g_probe_offsets[k]
                int res = scan_round(s + lane, s + lane + 1, lane < kFirstProbes, ip, cand);
                for (u32 base = kFirstProbes; res == 0; base += 32)
                    res = scan_round(s + g_probe_offsets[base + lane], s + g_probe_offsets[base + lane + 1], true,
                                     ip, cand);
                if (res == 2) break;
                // ---------------- copy chain, :211-239
                // One candidate-side memory round trip per copy: lane l compares byte l of the
                // candidate with byte l of ip, which verifies the 4-byte match of the previous
                // step's probe (:238) and measures the match length (:216) at once.  Meanwhile the
                // lanes pre-evaluate the post-copy probe (:228-235) for end position ip+4+lane from
                // ip-side bytes only (L1 hits).
                bool verified = true;  // a scan hit has compared its 4 bytes already (:193)
                for (;;) {
                    const u8* pa = F + cand + lane;
                    const u8* pb = F + ip + lane;
                    const u32 neq = __ballot_sync(kFullMask, __ldg(pa) != __ldg(pb));
                    const u32 e = ip + 4 + lane;
                    const uintptr_t ea = reinterpret_cast<uintptr_t>(F + e - 1);
                    const u32* ew = reinterpret_cast<const u32*>(ea & ~(uintptr_t)3);
                    const u32 elo = __ldg(ew), ehi = __ldg(ew + 1);
                    const u32 Wme = __funnelshift_r(elo, ehi, (u32)ea << 3);                 // bytes [e-1, e+3)
                    const u32 We = __funnelshift_rc(elo, ehi, (((u32)ea & 3u) << 3) + 8u);  // bytes [e, e+4)
                    const u32 He = hash(We), Hme = hash(Wme);
                    // only the `spec` shortest copies (M = 4 .. 3+spec) are pre-probed (the
                    // bulk of all copies): 32-lane gathers from L2 are what slows them down
                    const u32 te = (lane < spec) ? tget(He) : 0u;
                    const u32 ce = (Hme == He) ? (e - 1) : te;  // :233 is visible to :234
                    u32 M = neq ? (u32)__ffs((int)neq) - 1u : 32u;
                    if (!verified && M < 4) break;  // :238 no match at ip: back to scanning from ip+1
                    if (M == 32) {                  // long match: keep comparing, 32 bytes per round
                        while (ip + M < n) {
                            const u32 nq = __ballot_sync(kFullMask, __ldg(pa + M) != __ldg(pb + M));
                            if (nq) {
                                M += (u32)__ffs((int)nq) - 1u;
                                break;
                            }
                            M += 32;
                        }
                    }
                    if (ip + M > n) M = n - ip;  // find_match_length stops at the fragment end (:344-387)
                    keep(lit_from, ip, cand, M);  // :200,:217
                    ip += M;
                    lit_from = ip;
                    if ((int)ip >= lim) { finished = true; break; }  // :222
                    u32 c2;
                    if (M < 4 + spec) {
                        const u32 owner = M - 4;
                        c2 = __shfl_sync(kFullMask, ce, owner);
                        if (lane == owner) {
                            tput(Hme, ip - 1);  // :233
                            tput(He, ip);       // :235
                        }
                    } else {
                        const u32 hp = hash(ldg32u(F + ip - 1)), hc = hash(ldg32u(F + ip));
                        c2 = (hp == hc) ? (ip - 1) : tget(hc);  // :233-234
                        __syncwarp();
                        if (lane == 0) {
                            tput(hp, ip - 1);
                            tput(hc, ip);
                        }
                    }
                    __syncwarp();
                    cand = c2;
                    verified = false;
                }
            }
        }
        finish(lit_from);
    }
};

// One entry per shard for the batched shard API: fragments [frag_begin, frag_begin + nfrag) of the
// launch belong to this shard.
struct ShardDesc {
    const u8* ptr;   // first byte of the shard
    const u8* tail;  // padded copy of its last fragment
    u64 len;
    u32 frag_begin, nfrag, shift, pad_;
};

#ifdef SB200_EXPERIMENTS  // the step-wise kernel (option window=0, 20.9 ms per GiB); its Chain struct above is the slow path of K1w
// Persistent warps; fragments are pulled from *counter.  One CTA per SM with `blockDim.x / 32` warps:
// 7 shared-memory tables of 32 KiB fit one CTA (7 x 32 KiB + the 1 KiB the system reserves per CTA
// <= 227 KiB) where seven 1-warp CTAs would not.  Warps never synchronise with each other.
//   tail_copy : padded copy of the shard's LAST fragment (so reads may run past its end)
//   gtables   : kSmemTable == false: one 32 KiB table per warp in global memory
// The global-table variant is compiled for 3 CTAs of <= 14 warps per SM (<= 48 registers): many slow
// chains; the shared-table variant runs as one CTA per SM.
template <bool kSmemTable>
__global__ void __launch_bounds__(kSmemTable ? 224 : 448, kSmemTable ? 1 : 3)
k_compress_chain(const u8* __restrict__ g_in, u64 shard_len, u32 nfrag, u32 shift,
                 const u8* __restrict__ tail_copy, u8* __restrict__ scratch, u32* __restrict__ frag_sizes,
                 u32* __restrict__ counter, u16* __restrict__ gtables, u32 spec_lanes, u32 reserve,
                 const ShardDesc* __restrict__ descs, u32 ndesc) {
    extern __shared__ __align__(128) u8 smem[];
    const u32 warp = threadIdx.x >> 5;
    const u32 gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    u16* T = kSmemTable ? reinterpret_cast<u16*>(smem) + (size_t)warp * kMaxTableEntries
                        : gtables + (size_t)gwarp * kMaxTableEntries;
    const u32 lane = lane_id();
    for (;;) {
        // slow (global-table) warps leave the last `reserve` fragments to the fast ones, so that
        // the kernel does not end on a straggler
        if (reserve && *reinterpret_cast<volatile u32*>(counter) + reserve >= nfrag) break;
        u32 frag = 0;
        if (lane == 0) frag = atomicAdd(counter, 1u);
        frag = __shfl_sync(kFullMask, frag, 0);
        if (frag >= nfrag) break;
        // which shard (batched API: several shards share one launch), else the single shard
        const u8* sbase = g_in;
        const u8* stail = tail_copy;
        u64 slen = shard_len;
        u32 local = frag, lastf = nfrag - 1, fshift = shift;
        if (descs) {
            u32 k = 0;
            while (k + 1 < ndesc && descs[k + 1].frag_begin <= frag) k++;
            sbase = descs[k].ptr;
            stail = descs[k].tail;
            slen = descs[k].len;
            local = frag - descs[k].frag_begin;
            lastf = descs[k].nfrag - 1;
            fshift = descs[k].shift;
        }
        const u64 start = (u64)local * kBlockSize;
        const u32 n = (u32)((slen - start < kBlockSize) ? (slen - start) : kBlockSize);
        const u32 entries = 1u << (32 - fshift);
        uint4* t4 = reinterpret_cast<uint4*>(T);
        for (u32 i = lane; i < entries / 8; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        Chain<kSmemTable> ch;
        ch.F = (local == lastf) ? stail : sbase + start;
        ch.T = T;
        ch.Ts = kSmemTable ? smem_u32(T) : 0u;
        ch.out = scratch + (u64)frag * kSlotStride;
        ch.n = n;
        ch.shift = fshift;
        ch.lane = lane;
        ch.spec = spec_lanes;
        ch.make_policy();
        ch.run();
        if (lane == 0) frag_sizes[frag] = ch.op;
        __syncwarp();
    }
}

#endif  // SB200_EXPERIMENTS

}  // namespace sb200
