// compress_window.cuh -- K1w: window-parallel fragment compressor, one warp per fragment.
//
// Same serial decisions as the reference (src/internal.jl:127-250), evaluated 32 positions at a
// time instead of one decision per dependent memory round trip (compress_chain.cuh):
//
//   * A round looks at the window [a, a+32), lane l <-> position q = a + l.  Every lane evaluates
//     its position as if the chain arrived there: hash of the 4 bytes at q (:94), table lookup
//     against the table AS OF THE ROUND START, and the common prefix of candidate and q over 16
//     bytes (find_match_length, :344-387, capped).
//   * The real chain is then followed through the window with warp-uniform bit masks:
//       arrival after a copy (:228-238): insert q-1 and q, a match of >= 4 bytes is the next copy,
//         otherwise a scan starts at q+1;
//       scan (:167-194): every position up to the first hit is inserted; the hit emits the pending
//         literal and the copy.
//     A copy of m < 16 bytes jumps to lane l + m of the same window, so a round typically resolves
//     3-8 reference steps.
//   * Exactness: a lane's lookup equals the reference's iff no position inserted since the round
//     start has its hash.  All those positions lie in the window below it, so a lane whose hash
//     equals ANY lower lane's is never trusted: the round ends there and the next round starts at
//     that position as lane 0, which sees the committed table.  Inserts of the path are committed
//     at the end of the round, the highest position winning among equal hashes (:191).
//   * Left to the step-wise code of compress_chain.cuh: copies of >= 16 bytes (the whole warp
//     extends the match, 32 bytes per ballot) and scans that reach the stride > 1 part of the skip
//     heuristic (probe 32 onwards, :162-172) - incompressible data.
//
// Data path: the last `ring` bytes of the fragment up to ip + 64 live in a per-warp shared-memory
// ring (staged 512 bytes at a time with 16-byte loads), which serves every ip-side read and the
// candidates that are recent (the ring's first 32 bytes are mirrored behind its end, so a 20-byte read never
// wraps); older candidates are gathered from L1/L2, all 16 bytes in one round trip (SB200_FAR_ALL; fetching 4
// bytes first and the other 12 on a hit saves wavefronts but costs a second round trip: +6..9 %).
// tools/emulate_window.c is the CPU model of this file; it is checked against the oracle on every fixture.
// (The two compile-time switches below are measured and decided; the dead arms stay because removing them
// changed ptxas' schedule of the round for the worse: 26.1 vs 22.8 ms for the shared-table kernel in round 1, and
// again in round 2: 12.84 vs 12.19 ms for the merged kernel, profiles/r02zb_sweep_policy.txt.)
#pragma once
#include <type_traits>
#include "compress_chain.cuh"

namespace sb200 {

constexpr u32 kRingChunk = 512;  // bytes staged per step (one 16-byte load per lane)
constexpr u32 kRingAhead = 64;   // bytes past the window start that must be resident (+ 32 for the two-window round)
constexpr u32 kRingMirror = 32;  // the first bytes of the ring are repeated behind its end: a 20-byte read never wraps
#ifndef SB200_FAR_ALL
#define SB200_FAR_ALL 1
#endif
constexpr bool kFarAll = SB200_FAR_ALL != 0;
#ifdef SB200_CPU_EMU_STATS
static unsigned long g_emu_dist[17], g_emu_dist_hit[17];  // valid lanes by floor(log2(q - t)), and the hits among them
#endif
#ifdef SB200_CPU_EMU
static unsigned long g_emu_rounds = 0, g_emu_second = 0;  // tools/cpu_warp: rounds run / second windows entered
#endif
#ifndef SB200_M_BRANCHFREE
#define SB200_M_BRANCHFREE 1
#endif
// match length of a lane: the first differing word decides, one find-first-set instead of four (12.59 -> 12.40 ms
// per GiB of the mix, 13.59 -> 13.33 ms on source-like text, profiles/r02y_sweep_micro.txt).  Tried with it and
// dropped: no gather for the lanes whose hash equals a lower lane's (never trusted) -- slower, 14.6 ms on text.
#ifndef SB200_M_ONE_FFS
#define SB200_M_ONE_FFS 1
#endif

// kSlowCont (option `slowcont`, experimental, off): a copy of >= 16 bytes is extended inside the hop loop and the
// chain goes on in the same window when it lands there, instead of ending the round (tools/emulate_window.c
// SLOWCONT=1: 1-8 % fewer rounds).  Exact (tools/cpu_warp, and byte-identical on the B200); measured 13.58 vs 13.12 ms
// per GiB: slower, so off.
// kTwo (the two-window round): every round evaluates 64 positions, two per lane.  The first window is resolved as
// always; when the chain leaves it into the second one (a scan that runs off its end, a copy that lands up to 15
// bytes behind it) the second window is NOT evaluated again: its lanes re-read their table entry, a lane whose entry
// changed since the round start is untrusted (entries only grow, so any insert with its hash shows), and the chain is
// followed on.  The evaluation -- ring loads, hash, table lookup, candidate gather from L2 / DRAM, match length -- is
// the latency-bound part of a round, and the two windows' loads overlap.  (tools/emulate_window.c WW=2 is the model.)
template <int kSmemTable, bool kLib = false, bool kSlowCont = false, bool kTwo = false>
struct Win : Chain<kSmemTable, kLib> {
    using Base = Chain<kSmemTable, kLib>;
    using Base::F;
    using Base::lane;
    using Base::lim;
    using Base::n;
    // ring: positions [lo, hi) of the fragment at Rs + (position & rmask); hi is a multiple of 512
    u32 Rs, rmask, lo, hi, nstage;
    bool aligned16;
    uint4 pre;   // chunk [pre_at, pre_at + 512) loaded ahead of time (one 16-byte group per lane)
    u32 pre_at;

#ifdef SB200_CPU_EMU
    static u32 lds32(u32 a) { return *reinterpret_cast<const u32*>(smem + a); }
    template <int kOff>
    static u32 lds32o(u32 a) { return *reinterpret_cast<const u32*>(smem + a + kOff); }
    static void sts128(u32 a, uint4 v) { memcpy(smem + a, &v, 16); }
    static u32 lds8(u32 a) { return smem[a]; }
#else
    static __device__ __forceinline__ u32 lds32(u32 a) {
        u32 v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
        return v;
    }
    template <int kOff>
    static __device__ __forceinline__ u32 lds32o(u32 a) {  // word at a + kOff (immediate offset: no address math)
        u32 v;
        asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(kOff) : "memory");
        return v;
    }
#endif
    // unaligned 32-bit load at position p (lo <= p, p + 8 <= hi); reads may run into the mirror
    __device__ __forceinline__ u32 ring32u(u32 p) const {
        const u32 a = Rs + ((p & ~3u) & rmask);
        return __funnelshift_r(lds32o<0>(a), lds32o<4>(a), p << 3);
    }
    // one chunk (16 bytes per lane) of the fragment at position p
    __device__ __forceinline__ uint4 load_chunk(u32 p) const {
        uint4 v;
        if (aligned16) {
            v = __ldg(reinterpret_cast<const uint4*>(F + p));
        } else {
            v.x = ldg32u(F + p);
            v.y = ldg32u(F + p + 4);
            v.z = ldg32u(F + p + 8);
            v.w = ldg32u(F + p + 12);
        }
        return v;
    }
    // stage chunks until [.., upto) is resident or the fragment is exhausted.  The chunk behind the last
    // one staged is already on its way (pre, loaded when the previous one was stored), so that the load
    // latency is not exposed every 512 bytes.
    __device__ __forceinline__ void stage_to(u32 upto) {
        u32 from = lo;
        if (upto > hi + rmask + 1u) {  // a long copy ran past the ring: everything older is dropped
            hi = (upto - (rmask + 1u)) & ~(kRingChunk - 1u);
            from = hi;
            pre_at = 0xffffffffu;
        }
        while (hi < upto && hi < nstage) {
            const u32 p = hi + lane * 16u;
            const uint4 v = (pre_at == hi) ? pre : load_chunk(p);
            const u32 ro = p & rmask;
#ifdef SB200_CPU_EMU
            sts128(Rs + ro, v);
            if (ro < kRingMirror) sts128(Rs + rmask + 1u + ro, v);
#else
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(Rs + ro), "r"(v.x), "r"(v.y),
                         "r"(v.z), "r"(v.w)
                         : "memory");
            if (ro < kRingMirror)  // mirror of the ring's first bytes behind its end
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(Rs + rmask + 1u + ro), "r"(v.x),
                             "r"(v.y), "r"(v.z), "r"(v.w)
                             : "memory");
#endif
            hi += kRingChunk;
            if (hi < nstage) {  // the next chunk: issued now, stored when the window gets there
                pre = load_chunk(hi + lane * 16u);
                pre_at = hi;
            }
        }
        lo = hi > rmask + 1u ? hi - (rmask + 1u) : 0u;
        if (lo < from) lo = from;
        __syncwarp();
    }

    // whole-warp match extension from `M` known equal bytes (find_match_length, :344-387): 32 bytes per
    // ballot, from the ring while both sides are resident there (dictionary-like data: the candidate is
    // recent and the match ends within a few bytes), else from L1/L2
    __device__ __forceinline__ u32 extend(u32 ip, u32 cand, u32 M) const {
        while (ip + M < n) {
            u32 x, y;
            if (cand + M >= lo && ip + M + 32u <= hi) {
#ifdef SB200_CPU_EMU
                x = lds8(Rs + ((cand + M + lane) & rmask));
                y = lds8(Rs + ((ip + M + lane) & rmask));
#else
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(x) : "r"(Rs + ((cand + M + lane) & rmask)) : "memory");
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(y) : "r"(Rs + ((ip + M + lane) & rmask)) : "memory");
#endif
            } else {
                x = __ldg(F + cand + M + lane);
                y = __ldg(F + ip + M + lane);
            }
            const u32 nq = __ballot_sync(kFullMask, x != y);
            if (nq) {
                M += (u32)__ffs((int)nq) - 1u;
                break;
            }
            M += 32;
        }
        if (ip + M > n) M = n - ip;
        return M;
    }

    // One window: lane's position q as if the chain arrived there -- hash of its 4 bytes (:94), table lookup against
    // the table AS IT IS NOW, common prefix of candidate and position over 16 bytes (find_match_length, capped).
    //   H hash, t candidate position, m equal bytes (0 for an invalid lane), mp lanes with the same hash
    __device__ __forceinline__ void evaluate(const u32 q, const bool V, u32& H, u32& t, u32& m, u32& mp) const {
        u32 B0, B1, B2, B3;
        {
            const u32 qb = q & ~3u, sh = q << 3;
            const u32 ab = Rs + (qb & rmask);
            const u32 w0 = lds32o<0>(ab), w1 = lds32o<4>(ab), w2 = lds32o<8>(ab), w3 = lds32o<12>(ab),
                      w4 = lds32o<16>(ab);
            B0 = __funnelshift_r(w0, w1, sh);
            B1 = __funnelshift_r(w1, w2, sh);
            B2 = __funnelshift_r(w2, w3, sh);
            B3 = __funnelshift_r(w3, w4, sh);
        }
        H = this->hash(B0);
        t = V ? this->tget(H) : lo;
        mp = __match_any_sync(kFullMask, V ? H : (0x80000000u | lane));
        // candidate bytes, straight-line: recent candidates come from the ring (5 words), old
        // ones from L1/L2 (2 words = the 4 bytes that decide a hit; the other 3 only on a hit)
        const u32 nearp = (t >= lo) ? 1u : 0u;
        const uintptr_t ga = reinterpret_cast<uintptr_t>(F + t);  // F need not be 4-byte aligned
        const u32* g = reinterpret_cast<const u32*>(ga & ~(uintptr_t)3);
        const u32 tb = t & ~3u, tsh = (nearp ? t : (u32)ga) << 3;
        u32 c0, c1, c2 = 0, c3 = 0, c4 = 0;
        // far_all: old candidates fetch all 16 bytes at once (one L2 round trip, more wavefronts)
        // instead of 4 bytes first and the other 12 on a hit
        const u32 far_all = kFarAll ? 1u : 0u;
#ifdef SB200_CPU_EMU
        if (nearp) {
            const u32 ra = Rs + (tb & rmask);
            c0 = lds32o<0>(ra); c1 = lds32o<4>(ra); c2 = lds32o<8>(ra); c3 = lds32o<12>(ra); c4 = lds32o<16>(ra);
        } else {
            c0 = g[0]; c1 = g[1];
            if (far_all) { c2 = g[2]; c3 = g[3]; c4 = g[4]; }
        }
#else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.u32 p, %5, 0;\n"
            "@p ld.shared.u32 %0, [%6];\n"
            "@p ld.shared.u32 %1, [%6+4];\n"
            "@p ld.shared.u32 %2, [%6+8];\n"
            "@p ld.shared.u32 %3, [%6+12];\n"
            "@p ld.shared.u32 %4, [%6+16];\n"
            "@!p ld.global.nc.u32 %0, [%7];\n"
            "@!p ld.global.nc.u32 %1, [%7+4];\n"
            "setp.ne.and.u32 p, %8, 0, !p;\n"
            "@p ld.global.nc.u32 %2, [%7+8];\n"
            "@p ld.global.nc.u32 %3, [%7+12];\n"
            "@p ld.global.nc.u32 %4, [%7+16];\n"
            "}\n"
            : "=r"(c0), "=r"(c1), "+r"(c2), "+r"(c3), "+r"(c4)
            : "r"(nearp), "r"(Rs + (tb & rmask)), "l"(g), "r"(far_all)
            : "memory");
#endif
        u32 C0 = __funnelshift_r(c0, c1, tsh);
        const bool more = V && !nearp && !far_all && C0 == B0;
        if (__any_sync(kFullMask, more)) {
            if (more) {
                c2 = __ldg(g + 2);
                c3 = __ldg(g + 3);
                c4 = __ldg(g + 4);
            }
        }
        m = 0;
        {
            const u32 x0 = C0 ^ B0, x1 = __funnelshift_r(c1, c2, tsh) ^ B1,
                      x2 = __funnelshift_r(c2, c3, tsh) ^ B2, x3 = __funnelshift_r(c3, c4, tsh) ^ B3;
#if SB200_M_ONE_FFS
            // the first word that differs decides: one find-first-set instead of four
            const u32 w = x0 ? x0 : (x1 ? x1 : (x2 ? x2 : x3));
            const u32 base = x0 ? 0u : (x1 ? 4u : (x2 ? 8u : 12u));
            m = w ? base + (((u32)__ffs((int)w) - 1u) >> 3) : 16u;
#elif SB200_M_BRANCHFREE
            // equal bytes per word (0..4), then the length of the run of full words: no divergent branches
            const u32 m0 = x0 ? ((u32)__ffs((int)x0) - 1u) >> 3 : 4u, m1 = x1 ? ((u32)__ffs((int)x1) - 1u) >> 3 : 4u,
                      m2 = x2 ? ((u32)__ffs((int)x2) - 1u) >> 3 : 4u, m3 = x3 ? ((u32)__ffs((int)x3) - 1u) >> 3 : 4u;
            const bool f0 = m0 == 4u, f1 = f0 && m1 == 4u, f2 = f1 && m2 == 4u;
            m = m0 + (f0 ? m1 : 0u) + (f1 ? m2 : 0u) + (f2 ? m3 : 0u);
#else
            if (x0) m = ((u32)__ffs((int)x0) - 1u) >> 3;
            else if (x1) m = 4u + (((u32)__ffs((int)x1) - 1u) >> 3);
            else if (x2) m = 8u + (((u32)__ffs((int)x2) - 1u) >> 3);
            else if (x3) m = 12u + (((u32)__ffs((int)x3) - 1u) >> 3);
            else m = 16u;
#endif
            if (!V) m = 0;
        }
#ifdef SB200_CPU_EMU_STATS  // tools/cpu_warp: how far back the candidates are, and how many of them are hits
        if (V) {
            const u32 dist = q - t;
            u32 b = 0;
            while ((1u << (b + 1)) <= dist && b < 16) b++;
            g_emu_dist[b]++;
            if (m >= 4u) g_emu_dist_hit[b]++;
        }
#endif
    }

    enum : u32 { K_COPY = 0, K_SLOW = 1, K_FIN = 2, K_NEXTSCAN = 3, K_NEXTARR = 4, K_LEAVE = 5 };

    __device__ __forceinline__ void run_window() {
        asm volatile("" : "+r"(this->n), "+r"(this->Ts), "+r"(this->shift), "+r"(Rs), "+r"(rmask));
        this->op = 0;
        this->nrec = 0;
        this->r_lit = this->r_cpy = 0;
        lim = (int)n - Base::kLimMargin;  // ip_limit, :131
        u32 lit_from = 0;
        if (n >= kInputMargin) {
            bool arrival = false;  // round starts with a post-copy arrival at a (else: scanning)
            u32 a = 1, scan_s = 1;  // :162-163 the first scan starts at position 1
            for (;;) {
                // ------------- scans past 32 probes: stride > 1, step-wise (incompressible data)
                if (!arrival && a - scan_s >= 32u) {
                    u32 ip = 0, cand = 0;
                    int res = 0;
                    for (u32 base = 32; res == 0; base += 32)
                        res = this->scan_round(scan_s + g_probe_offsets[base + lane],
                                               scan_s + g_probe_offsets[base + lane + 1], true, ip, cand);
                    if (res == 2) break;
                    const u32 M = extend(ip, cand, 4);
                    this->keep(lit_from, ip, cand, M);  // :200,:217
                    a = ip + M;
                    lit_from = a;
                    if ((int)a >= lim) break;  // :222
                    arrival = true;
                    continue;
                }
                // ------------- lane evaluation against the table as of the round start
                if (a + kRingAhead + (kTwo ? 32u : 0u) > hi) stage_to(a + kRingAhead + (kTwo ? 32u : 0u));
                if (arrival) {  // :233 the position before an arrival is inserted first
                    if (lane == 0) this->tput(this->hash(ring32u(a - 1u)), a - 1u);
                    __syncwarp();
                }
#ifdef SB200_CPU_EMU
                if (lane == 0) g_emu_rounds++;
#endif
                u32 q = a + lane;
                bool V = (int)q < lim;
                u32 H, t, m, mp;
                evaluate(q, V, H, t, m, mp);
                // the second window of the round (kTwo): evaluated now, used only if the chain gets there
                u32 H1 = 0, t1 = 0, m1 = 0, mp1 = 0;
                bool V1 = false;
                const bool have_hi = kTwo && (int)(a + 32u) < lim;
                if (have_hi) {
                    V1 = (int)(q + 32u) < lim;
                    evaluate(q + 32u, V1, H1, t1, m1, mp1);
                }
                u32 changed = 0, entry = 0;  // second window: lanes whose table entry moved; lane the chain enters at
                u32 d, insacc, cur;
                for (u32 half = 0;; half++) {
                const u32 vmask = __ballot_sync(kFullMask, V);
                const u32 hitmask = __ballot_sync(kFullMask, m >= 4u);
                // untrusted lanes: hash equal to a lower lane's; in the second window also a table entry that moved
                const u32 dupmask = __ballot_sync(kFullMask, (mp & ((1u << lane) - 1u)) != 0u) | changed;
                const u32 tm = (t << 16) | (m << 8);
                // ------------- per-lane descriptor: what happens when the chain ARRIVES at this lane
                //   bits 0-2 kind, 3-7 lane e of the event, 8-12 copy length, 16-31 candidate, bit 13 = a scan started
                const u32 stop_all = ~vmask | dupmask | hitmask;
                u32 desc, ins;
                {
                    const u32 lbit = 1u << lane;
                    const u32 rest = (lane < 31u) ? (stop_all >> (lane + 1u)) : 0u;
                    const u32 es = lane + (u32)__ffs((int)rest);  // first event lane of a scan from lane + 1
                    const u32 ebit = 1u << (es & 31u);
                    // (selects, no branches: the lanes differ in all of these)
                    const bool hit = (hitmask & lbit) != 0u, none = rest == 0u;
                    const bool ev_invalid = (vmask & ebit) == 0u, ev_dup = (dupmask & ebit) != 0u;  // :175
                    // the next round starts at an untrusted lane (lane 0 of a round's first window looked the
                    // committed table up and is never one)
                    const bool untrusted = (lane != 0u || half != 0u) && (dupmask & lbit) != 0u;
                    const u32 above = (lane < 31u) ? (~0u << (lane + 1u)) : 0u;
                    const u32 kind_scan = none ? (u32)K_LEAVE
                                               : (ev_invalid ? (u32)K_FIN : (ev_dup ? (u32)K_NEXTSCAN : (u32)K_COPY));
                    // a scan inserts every position up to its first event, and the event itself if it is a hit (:191)
                    const u32 ins_scan = none ? above : (((ebit - 1u) & above) | ((ev_invalid || ev_dup) ? 0u : ebit));
                    u32 kind = hit ? (u32)K_COPY : kind_scan;
                    u32 e = hit ? lane : (none ? 0u : es);
                    ins = lbit | (lbit >> 1) | (hit ? 0u : ins_scan);  // :233,:235
                    kind = untrusted ? (u32)K_NEXTARR : kind;
                    e = untrusted ? lane : e;
                    ins = untrusted ? 0u : ins;
                    const u32 r = __shfl_sync(kFullMask, tm, e);  // candidate and length of the copy at e
                    if (kind == K_COPY && ((r >> 8) & 31u) == 16u) kind = K_SLOW;
                    desc = kind | (e << 3) | ((hitmask & lbit) ? 0u : (1u << 13)) | (r & 0xffff1f00u);
                }
                // a round that starts inside a scan: the same from "lane -1", and the scan also stops
                // where its probe count reaches 32 (:162-172)
                cur = entry;
                if (arrival) {
                    d = __shfl_sync(kFullMask, desc, entry);
                    insacc = __shfl_sync(kFullMask, ins, entry);
                } else {
                    const u32 klim = 32u - (a - scan_s);  // 1..32
                    const u32 limmask = klim < 32u ? ~((1u << klim) - 1u) : 0u;
                    const u32 rest = stop_all | limmask;
                    u32 kind, e = 0;
                    if (!rest) {
                        kind = K_LEAVE;
                        insacc = ~0u;
                    } else {
                        e = (u32)__ffs((int)rest) - 1u;
                        const u32 ebit = 1u << e;
                        insacc = ebit - 1u;
                        if (!(vmask & ebit)) kind = K_FIN;
                        else if ((dupmask | limmask) & ebit) kind = K_NEXTSCAN;
                        else {
                            kind = K_COPY;
                            insacc |= ebit;
                        }
                    }
                    const u32 r = __shfl_sync(kFullMask, tm, e);
                    if (kind == K_COPY && ((r >> 8) & 31u) == 16u) kind = K_SLOW;
                    d = kind | (e << 3) | (r & 0xffff1f00u);  // bit 13 clear: scan_s stays
                }
                // ------------- follow the chain through the window (warp-uniform)
                for (;;) {
                    if (d & (1u << 13)) scan_s = a + cur + 1u;  // :162 a new scan started behind lane cur
                    if (kSlowCont && (d & 7u) == K_SLOW) {
                        const u32 e = (d >> 3) & 31u, ip = a + e, cand = d >> 16;
                        const u32 M = extend(ip, cand, 16);
                        this->keep(lit_from, ip, cand, M);  // :200,:217
                        lit_from = ip + M;
                        if ((int)lit_from >= lim) {  // :222
                            d = K_FIN;
                            break;
                        }
                        if (e + M >= 32u) {
                            d = K_NEXTARR | (1u << 14);
                            break;
                        }
                        cur = e + M;
                        d = __shfl_sync(kFullMask, desc, cur);
                        insacc |= __shfl_sync(kFullMask, ins, cur);
                        continue;
                    }
                    if ((d & 7u) != K_COPY) break;
                    const u32 e = (d >> 3) & 31u, me = (d >> 8) & 31u;
                    this->keep(lit_from, a + e, d >> 16, me);  // :200,:217
                    cur = e + me;
                    lit_from = a + cur;
                    if ((int)lit_from >= lim) {  // :222
                        d = K_FIN;
                        break;
                    }
                    if (cur >= 32u) {
                        d = K_NEXTARR | (1u << 14);  // arrival beyond the window
                        break;
                    }
                    d = __shfl_sync(kFullMask, desc, cur);
                    insacc |= __shfl_sync(kFullMask, ins, cur);
                }
                const u32 ins_all = insacc;
                // ------------- commit the inserts of the path; the highest position wins (:191)
                if (((ins_all >> lane) & 1u) && (mp & ins_all & ~((2u << lane) - 1u)) == 0u) this->tput(H, q);
                __syncwarp();
                if (kTwo && half == 0u && have_hi) {
                    // does the chain go on in the second window?  a scan that ran off the first one (and has probes
                    // left at stride 1, :162-172), or a copy that landed behind it
                    const u32 k0 = d & 7u;
                    const bool leave = k0 == K_LEAVE && a + 32u - scan_s < 32u;
                    const bool landed = k0 == K_NEXTARR && (d & (1u << 14)) && lit_from - (a + 32u) < 32u;
                    if (leave || landed) {
#ifdef SB200_CPU_EMU
                        if (lane == 0) g_emu_second++;
#endif
                        entry = landed ? lit_from - (a + 32u) : 0u;
                        arrival = landed;
                        if (landed && entry == 0u && lane == 31u) this->tput(H, q);  // :233 position a + 31 goes in first
                        a += 32u;
                        __syncwarp();
                        // the second window's lookups are as old as the round: any insert since then with a lane's
                        // hash has replaced its entry (positions only grow)
                        const u32 tnow = V1 ? this->tget(H1) : t1;
                        changed = __ballot_sync(kFullMask, V1 && tnow != t1);
                        q += 32u;
                        V = V1;
                        H = H1;
                        t = t1;
                        m = m1;
                        mp = mp1;
                        continue;
                    }
                }
                break;
                }  // sub-rounds
                const u32 kind = d & 7u, ev = (d >> 3) & 31u;
                if (kind == K_SLOW) {  // copy of >= 16 bytes: the whole warp extends it
                    const u32 ip = a + ev, cand = d >> 16;
                    const u32 M = extend(ip, cand, 16);
                    this->keep(lit_from, ip, cand, M);
                    a = ip + M;
                    lit_from = a;
                    if ((int)a >= lim) break;
                    arrival = true;
                    continue;
                }
                if (kind == K_FIN) break;
                if (kind == K_LEAVE) {
                    arrival = false;
                    a += 32u;
                } else if (kind == K_NEXTSCAN) {
                    arrival = false;
                    a += ev;
                } else {  // K_NEXTARR: at an untrusted lane of this window, or beyond it
                    arrival = true;
                    a = (d & (1u << 14)) ? lit_from : a + ev;
                }
            }
        }
        this->finish(lit_from);
    }
};

// Arguments of a window-kernel launch (one struct: the two kernel forms below share them).
struct WindowArgs {
    const u8* g_in;        // the shard (nullptr with descs)
    u64 shard_len;
    u32 nfrag, shift;
    const u8* tail_copy;   // padded copy of the shard's last fragment (reads may run past a fragment's end)
    u8* scratch;           // per-fragment output slots of kSlotStride bytes
    u32* frag_sizes;
    u32* counter;          // fragments are pulled from here
    u16* gtables;          // one table per global-table warp
    u32 reserve;           // global-table warps stop pulling when fewer than this many fragments remain
    const ShardDesc* descs;
    u32 ndesc;
    const u32* ready;      // streamed input (host-buffer API): fragments resident so far
    u32* done;             // streamed output: finished fragments per chunk of done_div
    u32 done_div;
    u32 lib_rules;         // kLib: 1 = libsnappy <= 1.1.7 (hash >> shift, <= 16384 buckets), 2 = Google snappy >= 1.1.9
    const u32* order;      // the k-th fragment pulled is order[k] (schedule.cuh); the bytes do not depend on it
    u64* trace;            // option `trace`
};

// The life of one persistent warp: pull a fragment, clear the table, run the window rounds, record the size.
//   T / ring : this warp's table (shared or global memory) and ring (shared-space address)
//   kLib (option `rules`): libsnappy's rules, table sized per fragment; rules = 2: 64 KiB tables.
template <int kSmemTable, bool kLib>
struct Pipe;  // compress_pipe.cuh: the same round with its loads one round ahead (kPipe)
template <int kSmemTable, bool kLib, bool kSlowCont, bool kTwo = false, bool kPipe = false>
__device__ __forceinline__ void window_warp_loop(const WindowArgs& A, u16* T, u32 ring, u32 ring_bytes, u32 tab,
                                                 u32 reserve, bool dyn_smem = false) {
    const bool in_smem = kSmemTable == 2 ? dyn_smem : kSmemTable == 1;
    const u32 lane = lane_id();
    const u32 nfrag = A.nfrag;
    for (;;) {
        if (reserve && *reinterpret_cast<volatile u32*>(A.counter) + reserve >= nfrag) break;
        u32 frag = 0;
        if (lane == 0) frag = atomicAdd(A.counter, 1u);
        frag = __shfl_sync(kFullMask, frag, 0);
        if (frag >= nfrag) break;
        if (A.order) frag = A.order[frag];  // schedule.cuh: expensive fragments first (never together with `ready`)
        u64 t_begin = 0;
#ifndef SB200_CPU_EMU
        if (A.trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
#endif
        if (A.ready) {  // streamed input (host-buffer API): wait until this fragment and the one behind it
                        // (the kernel reads a few bytes past a fragment's end) have landed
            if (lane == 0) {
                const u32 need = frag + 2u < nfrag ? frag + 2u : nfrag;
                while (*reinterpret_cast<const volatile u32*>(A.ready) < need) __nanosleep(500);
            }
            __syncwarp();
        }
        const u8* sbase = A.g_in;
        const u8* stail = A.tail_copy;
        u64 slen = A.shard_len;
        u32 local = frag, lastf = nfrag - 1, fshift = A.shift;
        if (A.descs) {
            u32 k = 0;
            while (k + 1 < A.ndesc && A.descs[k + 1].frag_begin <= frag) k++;
            sbase = A.descs[k].ptr;
            stail = A.descs[k].tail;
            slen = A.descs[k].len;
            local = frag - A.descs[k].frag_begin;
            lastf = A.descs[k].nfrag - 1;
            fshift = A.descs[k].shift;
        }
        const u64 start = (u64)local * kBlockSize;
        const u32 n = (u32)((slen - start < kBlockSize) ? (slen - start) : kBlockSize);
        u32 entries = 1u << (32 - fshift);
        if (kLib) {  // GetHashTable: sized from THIS fragment's length
            entries = 256;
            while (entries < tab && entries < n) entries <<= 1;
            fshift = A.lib_rules == 2u ? 17u : (u32)__clz((int)entries) + 1u;  // 32 - log2(entries)
        }
        uint4* t4 = reinterpret_cast<uint4*>(T);
        for (u32 i = lane; i < entries / 8; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        typename std::conditional<kPipe, Pipe<kSmemTable, kLib>, Win<kSmemTable, kLib, kSlowCont, kTwo>>::type ch;
        ch.hmask = entries - 1u;
        ch.F = (local == lastf) ? stail : sbase + start;
        ch.T = T;
        ch.Ts = in_smem ? smem_u32(T) : 0u;
        ch.dyn_smem = dyn_smem;
        ch.out = A.scratch + (u64)frag * kSlotStride;
        ch.n = n;
        ch.shift = fshift;
        ch.lane = lane;
        ch.spec = 0;
        ch.make_policy();
        ch.Rs = ring;
        ch.rmask = ring_bytes - 1u;
        ch.lo = ch.hi = 0;
        ch.pre_at = 0xffffffffu;
        ch.nstage = (n + kRingChunk - 1u) & ~(kRingChunk - 1u);
        ch.aligned16 = (reinterpret_cast<uintptr_t>(ch.F) & 15u) == 0;
        if constexpr (kPipe) ch.run_pipe();
        else ch.run_window();
        if (lane == 0) A.frag_sizes[frag] = ch.op;
        if (A.trace && lane == 0) {  // option `trace`: [begin ns | table placement in bit 0, end ns | SM in the low 8 bits]
            u64 t_end = 0;
            u32 smid = 0;
#ifndef SB200_CPU_EMU
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
#endif
            A.trace[2 * (u64)frag] = (t_begin & ~(u64)1) | (in_smem ? 1u : 0u);
            A.trace[2 * (u64)frag + 1] = (t_end & ~(u64)0xff) | (smid & 0xffu);
        }
        if (A.done) {  // streamed output: per-chunk completion counts release the compaction of a chunk
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(A.done + frag / A.done_div, 1u);
        }
        __syncwarp();
    }
}

// One table placement per launch: every warp of the CTA owns a table (behind each other in shared memory, or in
// `gtables`) and `ring_bytes` of shared memory for its ring.  Used alone when the other placement is switched off.
template <bool kSmemTable, bool kLib = false, bool kSlowCont = false, bool kPipe = false>
__global__ void __launch_bounds__(kSmemTable ? 224 : 640, 1)
k_compress_window(const WindowArgs A, u32 ring_bytes) {
    extern __shared__ __align__(128) u8 smem[];
    const u32 warp = threadIdx.x >> 5;
    const u32 nwarp = blockDim.x >> 5;
    const u32 gwarp = blockIdx.x * nwarp + warp;
    const u32 tab = kLib ? (A.lib_rules == 2u ? 2u * kMaxTableEntries : kMaxTableEntries) : kMaxTableEntries;
    u16* T = kSmemTable ? reinterpret_cast<u16*>(smem) + (size_t)warp * tab : A.gtables + (size_t)gwarp * tab;
    const u32 ring = smem_u32(smem) + (kSmemTable ? nwarp * tab * 2u : 0u) + warp * (ring_bytes + kRingMirror);
    window_warp_loop<kSmemTable, kLib, kSlowCont, false, kPipe>(A, T, ring, ring_bytes, tab, kSmemTable ? 0u : A.reserve);
}

// Both table placements in ONE CTA per SM (the default): warps [0, wb) keep their table in global memory (L2),
// warps [wb, wb + wa) in shared memory.  The shared-table warps are the latency-bound ones (a lone dependent
// instruction stream each) and take the HIGHER warp numbers: the issue arbiter prefers higher warp slots, and as two
// separate kernels the shared-table CTA -- launched first, lower slots -- lost issue slots to the 14 global-table
// warps next to it (a text fragment took 3.6 ms on a shared-table warp with the other kernel running, 2.0 ms alone;
// profiles/r02b_trace_fragments.txt).  One kernel is also what lets ncu measure the pair as it really runs.
// Shared memory: wa tables, then wa rings of ring_a bytes, then wb rings of ring_b bytes.
// kTwoA / kTwoB: the two-window round (Win<.., kTwo>) for the shared-table / global-table warps.
// kPipe: bit 0 / bit 1 = the pipelined round (compress_pipe.cuh) for the shared-table / global-table warps.
// kUni: ONE copy of the round's code for both kinds of warp (the table placement is a run-time flag of the warp): half
// the instruction footprint of the hot loops, which is what the 32 KiB instruction cache of an SM sees.
template <bool kLib = false, bool kTwoA = false, bool kTwoB = false, int kPipe = 0, bool kUni = false>
__global__ void __launch_bounds__(640, 1)
k_compress_window_mixed(const WindowArgs A, u32 wa, u32 wb, u32 ring_a, u32 ring_b, u32 smem_first) {
    extern __shared__ __align__(128) u8 smem[];
    const u32 warp = threadIdx.x >> 5;
    const u32 tab = kLib ? (A.lib_rules == 2u ? 2u * kMaxTableEntries : kMaxTableEntries) : kMaxTableEntries;
    const u32 rings = smem_u32(smem) + wa * tab * 2u;
    // smem_first (experiment): the shared-table warps take the LOW warp numbers instead
    const bool is_smem = smem_first ? warp < wa : warp >= wb;
    if (kUni) {
        const u32 w = is_smem ? (smem_first ? warp : warp - wb) : (smem_first ? warp - wa : warp);
        u16* T = is_smem ? reinterpret_cast<u16*>(smem) + (size_t)w * tab : A.gtables + ((size_t)blockIdx.x * wb + w) * tab;
        const u32 ring = is_smem ? rings + w * (ring_a + kRingMirror)
                                 : rings + wa * (ring_a + kRingMirror) + w * (ring_b + kRingMirror);
        window_warp_loop<2, kLib, false, false, kPipe != 0>(A, T, ring, is_smem ? ring_a : ring_b, tab,
                                                            is_smem ? 0u : A.reserve, is_smem);
        return;
    }
    if (is_smem) {
        const u32 w = smem_first ? warp : warp - wb;
        u16* T = reinterpret_cast<u16*>(smem) + (size_t)w * tab;
        window_warp_loop<true, kLib, false, kTwoA, (kPipe & 1) != 0>(A, T, rings + w * (ring_a + kRingMirror), ring_a, tab, 0u);
    } else {
        const u32 w = smem_first ? warp - wa : warp;
        u16* T = A.gtables + ((size_t)blockIdx.x * wb + w) * tab;
        window_warp_loop<false, kLib, false, kTwoB, (kPipe & 2) != 0>(A, T, rings + wa * (ring_a + kRingMirror) + w * (ring_b + kRingMirror),
                                                    ring_b, tab, A.reserve);
    }
}

// K1bw: batched pages (one independent stream per page, src/Snappy.jl:20-36 per page) with the window round.
// Persistent warps pull pages from a counter; a warp holds its page's hash table AND the whole page in shared
// memory (the ring is as large as the largest page, so every candidate is "near": no global gathers at all), which
// limits this kernel to pages of at most ring_bytes <= 8 KiB (larger pages keep k_compress_pages).  The page's own
// varint header goes in front (varint.jl:46-69), the table is sized from the page length (Snappy.jl:27).
// 4 KiB pages: 8 KiB table + 4 KiB ring per warp, 16 warps per SM.
template <bool kLib = false>
__global__ void __launch_bounds__(512, 1)
k_compress_pages_window(const u8* __restrict__ g_in, const u64* __restrict__ in_off, const u32* __restrict__ in_size,
                        u32 count, u8* __restrict__ g_out, const u64* __restrict__ out_off, u32* __restrict__ out_size,
                        u32 ring_bytes, u32 table_cap, u32* __restrict__ counter, u32 lib_rules = 0) {
    extern __shared__ __align__(128) u8 smem[];
    const u32 warp = threadIdx.x >> 5;
    const u32 nwarp = blockDim.x >> 5;
    const u32 lane = lane_id();
    u16* T = reinterpret_cast<u16*>(smem) + (size_t)warp * table_cap;
    u8* R = smem + (size_t)nwarp * table_cap * 2u + (size_t)warp * (ring_bytes + kRingMirror);
    for (;;) {
        u32 pg = 0;
        if (lane == 0) pg = atomicAdd(counter, 1u);
        pg = __shfl_sync(kFullMask, pg, 0);
        if (pg >= count) break;
        const u8* pin = g_in + in_off[pg];
        const u32 n = in_size[pg];
        u8* pout = g_out + out_off[pg];
        u32 hdr = 0;
        {   // varint header, src/varint.jl:46-69
            u32 v = n;
            while (v >= 0x80) {
                if (lane == 0) pout[hdr] = (u8)(v | 0x80);
                v >>= 7;
                hdr++;
            }
            if (lane == 0) pout[hdr] = (u8)v;
            hdr++;
        }
        u32 entries = 256;  // alloc_hashtable, src/internal.jl:107-113 (kLib: GetHashTable, the same for one fragment)
        while (entries < table_cap && entries < n) entries <<= 1;
        u32 fshift = (u32)__clz((int)entries) + 1u;  // 32 - log2(entries)
        if (kLib && lib_rules == 2u) fshift = 17u;
        uint4* t4 = reinterpret_cast<uint4*>(T);
        for (u32 i = lane; i < entries / 8; i += 32) t4[i] = make_uint4(0, 0, 0, 0);
        // the whole page into the ring (zero behind its end), the ring's first bytes again behind the ring
        const bool al = (reinterpret_cast<uintptr_t>(pin) & 15u) == 0;
        for (u32 i = lane * 16u; i < ring_bytes; i += 512u) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (i + 16u <= n && al) {
                v = __ldg(reinterpret_cast<const uint4*>(pin + i));
            } else if (i < n) {
                u32 w[4] = {0, 0, 0, 0};
                for (u32 k = 0; k < 16u && i + k < n; k++) w[k >> 2] |= (u32)__ldg(pin + i + k) << (8u * (k & 3u));
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4*>(R + i) = v;
            if (i < kRingMirror) *reinterpret_cast<uint4*>(R + ring_bytes + i) = v;
        }
        __syncwarp();
        Win<true, kLib> ch;
        ch.hmask = entries - 1u;
        ch.F = pin;
        ch.T = T;
        ch.Ts = smem_u32(T);
        ch.out = pout + hdr;
        ch.n = n;
        ch.shift = fshift;
        ch.lane = lane;
        ch.spec = 0;
        ch.make_policy();
        ch.Rs = smem_u32(R);
        ch.rmask = ring_bytes - 1u;
        ch.lo = 0;
        ch.hi = 0x7fffff00u;  // everything is resident: the round never stages, every candidate is near
        ch.pre_at = 0xffffffffu;
        ch.nstage = 0;
        ch.aligned16 = false;
        ch.run_window();
        if (lane == 0) out_size[pg] = hdr + ch.op;
        __syncwarp();
    }
}

}  // namespace sb200
