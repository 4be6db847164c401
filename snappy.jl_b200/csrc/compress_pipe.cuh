// compress_pipe.cuh -- K1p: the window round of compress_window.cuh with its loads one round ahead.
//
// What the window kernel waits for (profiles/r02c_prof_compress_mixed.txt, per source line): 42 % of all warp time
// is the long scoreboard of ONE spot, the first use of the far candidates' bytes (L2 / DRAM gathers, 32 random lines)
// and of the global tables' entries.  A round cannot start its lookups before the previous round has committed its
// inserts -- unless a lookup made too early can be recognised: table entries only ever grow (a later position
// replaces an earlier one), so an evaluation made against an OLDER table is still the reference's iff the entry
// read then is the entry now.  So:
//
//   * lane l owns the positions q = l (mod 32); it keeps the evaluations (hash, candidate, equal bytes) of its two
//     positions in [a, a + 64) in two register slots (slot = bit 5 of q);
//   * every round ISSUES the evaluation loads of the positions that enter [f, a + 64) (f = frontier of what has been
//     evaluated), and only reduces them at the start of the NEXT round, behind a whole round of resolve work;
//   * the window [a, a + 32) is taken from the slots; each lane re-reads its table entry (the only exposed load of
//     a round), and a lane whose entry moved is evaluated again on the spot, so that after this step every lane holds
//     exactly what compress_window.cuh's evaluate() returns: the table as of the round start.  The resolve part
//     (descriptors, hop loop, commit) is that file's, unchanged;
//   * a copy may land up to 15 bytes behind the window, so the next window can be evaluated only up to f: the lanes
//     above are marked untrusted (a scan that reaches them, or a copy that lands there, ends the round exactly like
//     at a lane with an equal hash below it); after a long jump (copies of >= 16 bytes, stride > 1 scans) nothing
//     is held and the window is evaluated on the spot (a "cold" round, the window kernel's normal round).
//
// tools/cpu_warp runs this source against the oracle (tests/test_kernel_on_cpu_warp.py).
#pragma once
#include "compress_window.cuh"

namespace sb200 {

#ifndef SB200_PIPE_MINW
#define SB200_PIPE_MINW 12
#endif
constexpr u32 kPipeMinW = SB200_PIPE_MINW;  // fewer evaluated positions than this in front of a: a cold round instead
#ifdef SB200_CPU_EMU
static unsigned long g_emu_pipe_rounds = 0, g_emu_pipe_cold = 0, g_emu_pipe_fix = 0, g_emu_pipe_w = 0;
#endif

template <int kSmemTable, bool kLib = false>
struct Pipe : Win<kSmemTable, kLib> {
    using W = Win<kSmemTable, kLib>;
    using Base = Chain<kSmemTable, kLib>;
    using Base::F;
    using Base::lane;
    using Base::lim;
    using Base::n;
    using W::hi;
    using W::lo;
    using W::rmask;
    using W::Rs;

    // the loads of one evaluation: the 16 + 4 bytes at q from the ring, hash, table entry, candidate bytes (ring or
    // L1/L2).  tp = candidate | (1 << 16 when it was read from the ring); invalid lanes read harmless ring bytes.
    __device__ __forceinline__ void issue(const u32 q, const bool V, u32& H, u32& tp, u32& B0, u32& B1, u32& B2, u32& B3,
                                          u32& c0, u32& c1, u32& c2, u32& c3, u32& c4) const {
        {
            const u32 qb = q & ~3u, sh = q << 3;
            const u32 ab = Rs + (qb & rmask);
            const u32 w0 = W::template lds32o<0>(ab), w1 = W::template lds32o<4>(ab), w2 = W::template lds32o<8>(ab),
                      w3 = W::template lds32o<12>(ab), w4 = W::template lds32o<16>(ab);
            B0 = __funnelshift_r(w0, w1, sh);
            B1 = __funnelshift_r(w1, w2, sh);
            B2 = __funnelshift_r(w2, w3, sh);
            B3 = __funnelshift_r(w3, w4, sh);
        }
        H = this->hash(B0);
        const u32 t = V ? this->tget(H) : lo;
        gather(t, tp, c0, c1, c2, c3, c4);
    }
    __device__ __forceinline__ void gather(const u32 t, u32& tp, u32& c0, u32& c1, u32& c2, u32& c3, u32& c4) const {
        const u32 nearp = (t >= lo) ? 1u : 0u;
        tp = t | (nearp << 16);
        const uintptr_t ga = reinterpret_cast<uintptr_t>(F + t);  // F need not be 4-byte aligned
        const u32* g = reinterpret_cast<const u32*>(ga & ~(uintptr_t)3);
        const u32 ra = Rs + ((t & ~3u) & rmask);
#ifdef SB200_CPU_EMU
        if (nearp) {
            c0 = W::template lds32o<0>(ra); c1 = W::template lds32o<4>(ra); c2 = W::template lds32o<8>(ra);
            c3 = W::template lds32o<12>(ra); c4 = W::template lds32o<16>(ra);
        } else {
            c0 = g[0]; c1 = g[1]; c2 = g[2]; c3 = g[3]; c4 = g[4];
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
            "@!p ld.global.nc.u32 %2, [%7+8];\n"
            "@!p ld.global.nc.u32 %3, [%7+12];\n"
            "@!p ld.global.nc.u32 %4, [%7+16];\n"
            "}\n"
            : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4)
            : "r"(nearp), "r"(ra), "l"(g)
            : "memory");
#endif
    }
    // equal bytes of candidate and position over 16 bytes (find_match_length, src/internal.jl:344-387, capped)
    __device__ __forceinline__ u32 reduce(const u32 tp, const u32 B0, const u32 B1, const u32 B2, const u32 B3,
                                          const u32 c0, const u32 c1, const u32 c2, const u32 c3, const u32 c4) const {
        const u32 t = tp & 0xffffu;
        const u32 tsh = ((tp >> 16) ? t : (u32)reinterpret_cast<uintptr_t>(F + t)) << 3;
        const u32 x0 = __funnelshift_r(c0, c1, tsh) ^ B0, x1 = __funnelshift_r(c1, c2, tsh) ^ B1,
                  x2 = __funnelshift_r(c2, c3, tsh) ^ B2, x3 = __funnelshift_r(c3, c4, tsh) ^ B3;
        const u32 m0 = x0 ? ((u32)__ffs((int)x0) - 1u) >> 3 : 4u, m1 = x1 ? ((u32)__ffs((int)x1) - 1u) >> 3 : 4u,
                  m2 = x2 ? ((u32)__ffs((int)x2) - 1u) >> 3 : 4u, m3 = x3 ? ((u32)__ffs((int)x3) - 1u) >> 3 : 4u;
        const bool f0 = m0 == 4u, f1 = f0 && m1 == 4u, f2 = f1 && m2 == 4u;
        return m0 + (f0 ? m1 : 0u) + (f1 ? m2 : 0u) + (f2 ? m3 : 0u);
    }

    enum : u32 { K_COPY = 0, K_SLOW = 1, K_FIN = 2, K_NEXTSCAN = 3, K_NEXTARR = 4, K_LEAVE = 5 };

    __device__ __forceinline__ void run_pipe() {
        asm volatile("" : "+r"(this->n), "+r"(this->Ts), "+r"(this->shift), "+r"(Rs), "+r"(rmask));
        this->op = 0;
        this->nrec = 0;
        this->r_lit = this->r_cpy = 0;
        lim = (int)n - Base::kLimMargin;  // ip_limit, :131
        u32 lit_from = 0;
        if (n >= kInputMargin) {
            bool arrival = false;   // round starts with a post-copy arrival at a (else: scanning)
            u32 a = 1, scan_s = 1;  // :162-163 the first scan starts at position 1
            u32 f = 0;              // the slots hold the evaluations of the positions in [a, f), f <= a + 64
            u32 sH0 = 0, sT0 = 0, sH1 = 0, sT1 = 0;  // slot (q >> 5) & 1 of lane q & 31: hash, (candidate << 16) | (equal bytes << 8)
            // evaluations issued in the previous round, reduced at the start of this one: positions [f, pf)
            u32 pf = 0, pq = 0, pH = 0, ptp = 0, pB0 = 0, pB1 = 0, pB2 = 0, pB3 = 0, pc0 = 0, pc1 = 0, pc2 = 0, pc3 = 0, pc4 = 0;
            bool pV = false;
            for (;;) {
                // ------------- scans past 32 probes: stride > 1, step-wise (incompressible data)
                if (!arrival && a - scan_s >= 32u) {
                    u32 ip = 0, cand = 0;
                    int res = 0;
                    for (u32 base = 32; res == 0; base += 32)
                        res = this->scan_round(scan_s + g_probe_offsets[base + lane],
                                               scan_s + g_probe_offsets[base + lane + 1], true, ip, cand);
                    if (res == 2) break;
                    const u32 M = this->extend(ip, cand, 4);
                    this->keep(lit_from, ip, cand, M);  // :200,:217
                    a = ip + M;
                    lit_from = a;
                    if ((int)a >= lim) break;  // :222
                    arrival = true;
                    continue;
                }
                if (a + kRingAhead + 32u > hi) this->stage_to(a + kRingAhead + 32u);
                if (arrival) {  // :233 the position before an arrival is inserted first
                    if (lane == 0) this->tput(this->hash(this->ring32u(a - 1u)), a - 1u);
                    __syncwarp();
                }
                // ------------- last round's loads become evaluations
                if (pf) {
                    const u32 m = reduce(ptp, pB0, pB1, pB2, pB3, pc0, pc1, pc2, pc3, pc4);
                    const u32 T = (ptp << 16) | (m << 8);
                    if (pV) {
                        if (pq & 32u) {
                            sH1 = pH;
                            sT1 = T;
                        } else {
                            sH0 = pH;
                            sT0 = T;
                        }
                    }
                    f = pf;
                }
                // ------------- this lane's position of the window, from its slot; a cold round has none
                const bool cold = (int)(f - a) < (int)kPipeMinW;
                const u32 w = cold ? 32u : (f - a < 32u ? f - a : 32u);
                const u32 k = (lane - a) & 31u, q0 = a + k;
                const bool Vo = (int)q0 < lim && k < w;
                u32 Hc = (q0 & 32u) ? sH1 : sH0, Tc = (q0 & 32u) ? sT1 : sT0;
                u32 tnow = 0;
                if (!cold && Vo) tnow = this->tget(Hc);  // the one exposed load of a round
                // ------------- next round's loads: the positions that enter [fl, a + 64), one per lane at most
                {
                    const u32 fl = cold ? a + 32u : f;
                    pf = fl + 32u < a + 64u ? fl + 32u : a + 64u;
                    pq = fl + ((lane - fl) & 31u);
                    pV = pq < pf && (int)pq < lim;
                    issue(pq, pV, pH, ptp, pB0, pB1, pB2, pB3, pc0, pc1, pc2, pc3, pc4);
                    f = fl;
                }
#ifdef SB200_CPU_EMU
                if (lane == 0) {
                    g_emu_pipe_rounds++;
                    g_emu_pipe_cold += cold;
                    g_emu_pipe_w += w;
                }
#endif
                // ------------- lanes whose table entry moved since their evaluation (and all of a cold round): again
                const bool need = cold ? (int)q0 < lim : (Vo && tnow != (Tc >> 16));
                if (__any_sync(kFullMask, need)) {
#ifdef SB200_CPU_EMU
                    if (lane == 0 && !cold) g_emu_pipe_fix++;
#endif
                    u32 H, tp, B0, B1, B2, B3, c0, c1, c2, c3, c4;
                    issue(q0, need, H, tp, B0, B1, B2, B3, c0, c1, c2, c3, c4);
                    const u32 m = reduce(tp, B0, B1, B2, B3, c0, c1, c2, c3, c4);
                    if (need) {
                        Hc = H;
                        Tc = (tp << 16) | (m << 8);
                        if (q0 & 32u) {
                            sH1 = Hc;
                            sT1 = Tc;
                        } else {
                            sH0 = Hc;
                            sT0 = Tc;
                        }
                    }
                }
                // ------------- into window order: lane l <-> position a + l
                const u32 q = a + lane;
                const bool V = (int)q < lim;
                const bool have = V && lane < w;
                u32 H, tm;
                {
                    const u32 src = q & 31u;
                    H = __shfl_sync(kFullMask, Hc, src);
                    tm = __shfl_sync(kFullMask, Tc, src) & 0xffff1f00u;
                    if (!have) tm = 0;
                }
                const u32 mp = __match_any_sync(kFullMask, have ? H : (0x80000000u | lane));
                const u32 vmask = __ballot_sync(kFullMask, V);
                const u32 hitmask = __ballot_sync(kFullMask, (tm & 0x1f00u) >= 0x400u);
                // untrusted lanes: hash equal to a lower lane's, and everything above the evaluated part of the window
                const u32 dupmask = __ballot_sync(kFullMask, (mp & ((1u << lane) - 1u)) != 0u) | (w < 32u ? ~0u << w : 0u);
                // ------------- per-lane descriptor: what happens when the chain ARRIVES at this lane
                //   bits 0-2 kind, 3-7 lane e of the event, 8-12 copy length, 16-31 candidate, bit 13 = a scan started
                const u32 stop_all = ~vmask | dupmask | hitmask;
                u32 desc, ins;
                {
                    const u32 lbit = 1u << lane;
                    const u32 rest = (lane < 31u) ? (stop_all >> (lane + 1u)) : 0u;
                    const u32 es = lane + (u32)__ffs((int)rest);  // first event lane of a scan from lane + 1
                    const u32 ebit = 1u << (es & 31u);
                    const bool hit = (hitmask & lbit) != 0u, none = rest == 0u;
                    const bool ev_invalid = (vmask & ebit) == 0u, ev_dup = (dupmask & ebit) != 0u;  // :175
                    const bool untrusted = lane != 0u && (dupmask & lbit) != 0u;
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
                    desc = kind | (e << 3) | ((hitmask & lbit) ? 0u : (1u << 13)) | r;
                }
                // a round that starts inside a scan: the same from "lane -1", and the scan also stops
                // where its probe count reaches 32 (:162-172)
                u32 d, insacc, cur = 0;
                if (arrival) {
                    d = __shfl_sync(kFullMask, desc, 0);
                    insacc = __shfl_sync(kFullMask, ins, 0);
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
                    d = kind | (e << 3) | r;  // bit 13 clear: scan_s stays
                }
                // ------------- follow the chain through the window (warp-uniform)
                for (;;) {
                    if (d & (1u << 13)) scan_s = a + cur + 1u;  // :162 a new scan started behind lane cur
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
                // ------------- commit the inserts of the path; the highest position wins (:191)
                if (((insacc >> lane) & 1u) && (mp & insacc & ~((2u << lane) - 1u)) == 0u) this->tput(H, q);
                __syncwarp();
                const u32 kind = d & 7u, ev = (d >> 3) & 31u;
                if (kind == K_SLOW) {  // copy of >= 16 bytes: the whole warp extends it
                    const u32 ip = a + ev, cand = d >> 16;
                    const u32 M = this->extend(ip, cand, 16);
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

}  // namespace sb200
