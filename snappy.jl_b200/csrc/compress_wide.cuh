// compress_wide.cuh -- K1x: the window-parallel compressor (compress_window.cuh) with kW warps per
// fragment.
//
// One warp advances a fragment at ~8 cycles per instruction, whatever the instruction: the chain of
// a fragment is a single dependent instruction stream and the SM's issue slots stay ~85 % idle
// (profiles/).  The expensive part of a window round, the evaluation of 32 positions (ring reads,
// hash, table gather, candidate gather, 16-byte compare), does not depend on the chain, so kW warps
// evaluate kW consecutive windows at once and only the cheap part stays serial:
//
//   round:  every warp k evaluates window [a + 32k, a + 32k + 32) against whatever the table holds;
//           then the chain passes through the windows in order (a token travels warp 0 -> kW-1
//           through mbarriers): the holder re-reads the table for its 32 positions - a lane is
//           trusted iff the table STILL holds the value it looked up (every insert of the windows
//           before it has been committed by then) and no lower lane of its own window has its hash
//           - follows the chain through its window exactly as compress_window.cuh does, commits
//           its inserts and hands the state on.  A lane that cannot be trusted ends the round; the
//           next round starts at that position as lane 0 of warp 0.
//
// The records of the chain go through a 32-entry queue in shared memory and are turned into bytes by
// whichever warp holds the token when the queue fills.  tools/emulate_window.c (WW=2|4) is the CPU
// model; it is checked against the oracle on every fixture.
#pragma once
#include "compress_window.cuh"

namespace sb200 {

struct WideRound {  // state at the start of a round (double-buffered by round parity)
    u32 a, arrival, scan_s, lit_from, fin, pad_[3];
};

template <int kW>
struct WideCtl {
    u64 tok[4];        // tok[k]: warp k-1 -> warp k hand-off (mbarrier, one arrival per round)
    WideRound st[2];
    u32 cur, scanning, done, fin;   // chain state inside a round
    u32 scan_s, lit_from, op, nrec;
    u32 hi, lo, frag, pad_;
    u64 rec[32];       // record queue: (lit_from | ip << 16), (cand | M << 16)
};

template <int kW>
struct Wide : Win<true> {
    using Base = Win<true>;
    u32 Qs;  // shared address of the record queue

    __device__ __forceinline__ void keepq(u32 from, u32 ip, u32 cand, u32 M) {
        if (lane == 0)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(Qs + 8u * this->nrec), "r"(from | (ip << 16)),
                         "r"(cand | (M << 16))
                         : "memory");
        if (++this->nrec == 32) drain();
    }
    // the queued records become bytes (one record per lane)
    __device__ __forceinline__ void drain() {
        __syncwarp();
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(this->r_lit), "=r"(this->r_cpy) : "r"(Qs + 8u * lane) : "memory");
        this->flush();
        __syncwarp();
    }
};

enum : u32 { W_COPY = 0, W_SLOW = 1, W_FIN = 2, W_NEXTSCAN = 3, W_NEXTARR = 4, W_LEAVE = 5 };

__device__ __forceinline__ void chain_barrier(u32 id, u32 nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(u32 addr, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WIDE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WIDE_DONE;\n"
        "bra WIDE_WAIT;\n"
        "WIDE_DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}

// One CTA per SM: `chains` fragments in flight, kW warps each.  Shared memory: tables, rings, control blocks.
template <int kW>
__global__ void __launch_bounds__(768, 1)
k_compress_wide(const u8* __restrict__ g_in, u64 shard_len, u32 nfrag, u32 shift,
                const u8* __restrict__ tail_copy, u8* __restrict__ scratch, u32* __restrict__ frag_sizes,
                u32* __restrict__ counter, const ShardDesc* __restrict__ descs, u32 ndesc, u32 ring_bytes) {
    extern __shared__ __align__(128) u8 smem[];
    const u32 wi = threadIdx.x >> 5, lane = lane_id();
    const u32 chains = (blockDim.x >> 5) / kW;
    const u32 c = wi / kW, k = wi % kW;
    const u32 bar_id = 1u + c, bar_n = 32u * kW;
    u16* T = reinterpret_cast<u16*>(smem) + (size_t)c * kMaxTableEntries;
    const u32 ring = smem_u32(smem) + chains * kMaxTableEntries * 2u + c * (ring_bytes + kRingMirror);
    WideCtl<kW>* ctl = reinterpret_cast<WideCtl<kW>*>(smem + (size_t)chains * (kMaxTableEntries * 2u + ring_bytes + kRingMirror)) + c;
    volatile WideCtl<kW>* vc = ctl;
    if (k == 0 && lane == 0) {
        for (int j = 0; j < 4; j++) mbar_init(&ctl->tok[j], 1);
        fence_mbar_init();
    }
    u32 tok_parity = 0;
    chain_barrier(bar_id, bar_n);
    for (;;) {
        if (k == 0 && lane == 0) vc->frag = atomicAdd(counter, 1u);
        chain_barrier(bar_id, bar_n);
        const u32 frag = vc->frag;
        if (frag >= nfrag) break;
        const u8* sbase = g_in;
        const u8* stail = tail_copy;
        u64 slen = shard_len;
        u32 local = frag, lastf = nfrag - 1, fshift = shift;
        if (descs) {
            u32 j = 0;
            while (j + 1 < ndesc && descs[j + 1].frag_begin <= frag) j++;
            sbase = descs[j].ptr;
            stail = descs[j].tail;
            slen = descs[j].len;
            local = frag - descs[j].frag_begin;
            lastf = descs[j].nfrag - 1;
            fshift = descs[j].shift;
        }
        const u64 start = (u64)local * kBlockSize;
        const u32 n = (u32)((slen - start < kBlockSize) ? (slen - start) : kBlockSize);
        const u32 entries = 1u << (32 - fshift);
        uint4* t4 = reinterpret_cast<uint4*>(T);
        for (u32 i = k * 32u + lane; i < entries / 8; i += 32u * kW) t4[i] = make_uint4(0, 0, 0, 0);
        Wide<kW> ch;
        ch.F = (local == lastf) ? stail : sbase + start;
        ch.T = T;
        ch.Ts = smem_u32(T);
        ch.out = scratch + (u64)frag * kSlotStride;
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
        ch.Qs = smem_u32(&ctl->rec[0]);
        ch.op = 0;
        ch.nrec = 0;
        ch.r_lit = ch.r_cpy = 0;
        ch.lim = (int)n - 16;  // ip_limit, :131
        const int lim = ch.lim;
        u32 rp = 0;  // round parity: st[rp] is this round's start state
        if (k == 0 && lane == 0) {
            vc->st[0].a = 1;  // :162-163 the first scan starts at position 1
            vc->st[0].arrival = 0;
            vc->st[0].scan_s = 1;
            vc->st[0].lit_from = 0;
            vc->st[0].fin = (n >= kInputMargin) ? 0u : 1u;
            vc->op = 0;
            vc->nrec = 0;
            vc->hi = 0;
            vc->lo = 0;
        }
        chain_barrier(bar_id, bar_n);
        u32 lit_from = 0;
        for (;;) {  // rounds
            const u32 a = vc->st[rp].a, arrival = vc->st[rp].arrival;
            u32 scan_s = vc->st[rp].scan_s;
            lit_from = vc->st[rp].lit_from;
            if (vc->st[rp].fin) break;
            volatile WideRound* nx = &vc->st[rp ^ 1u];
            rp ^= 1u;
            // ------------- scans past 32 probes: stride > 1, step-wise, warp 0 (incompressible data)
            if (!arrival && a - scan_s >= 32u) {
                if (k == 0) {
                    ch.op = vc->op;
                    ch.nrec = vc->nrec;
                    u32 ip = 0, cand = 0, fin = 0, na = a;
                    int res = 0;
                    for (u32 base = 32; res == 0; base += 32)
                        res = ch.scan_round(scan_s + g_probe_offsets[base + lane],
                                            scan_s + g_probe_offsets[base + lane + 1], true, ip, cand);
                    if (res == 2) {
                        fin = 1;
                    } else {
                        const u32 M = ch.extend(ip, cand, 4);
                        ch.keepq(lit_from, ip, cand, M);  // :200,:217
                        na = ip + M;
                        lit_from = na;
                        if ((int)na >= lim) fin = 1;  // :222
                    }
                    if (lane == 0) {
                        nx->a = na;
                        nx->arrival = 1;
                        nx->scan_s = scan_s;
                        nx->lit_from = lit_from;
                        nx->fin = fin;
                        vc->op = ch.op;
                        vc->nrec = ch.nrec;
                    }
                }
                chain_barrier(bar_id, bar_n);
                continue;
            }
            // ------------- ring: [.., a + 32 kW + 64) must be resident (warp 0 stages, everybody waits)
            {
                const u32 need = a + 32u * kW + kRingAhead;
                if (need > vc->hi && vc->hi < ch.nstage) {
                    if (k == 0) {
                        ch.hi = vc->hi;
                        ch.lo = vc->lo;
                        ch.stage_to(need + kRingChunk);
                        if (lane == 0) {
                            vc->hi = ch.hi;
                            vc->lo = ch.lo;
                        }
                    }
                    chain_barrier(bar_id, bar_n);
                }
                ch.hi = vc->hi;
                ch.lo = vc->lo;
            }
            // ------------- every warp evaluates its window against whatever the table holds now
            const u32 base = a + 32u * k;
            if (k == 0 && arrival) {  // :233 the position before an arrival is inserted first
                if (lane == 0) ch.tput(ch.hash(ch.ring32u(a - 1u)), a - 1u);
                __syncwarp();
            }
            const u32 q = base + lane;
            const bool V = (int)q < lim;
            u32 B0, B1, B2, B3;
            {
                const u32 qb = q & ~3u, sh = q << 3;
                const u32 w0 = ch.lds32(ch.Rs + (qb & ch.rmask)), w1 = ch.lds32(ch.Rs + ((qb + 4u) & ch.rmask)),
                          w2 = ch.lds32(ch.Rs + ((qb + 8u) & ch.rmask)), w3 = ch.lds32(ch.Rs + ((qb + 12u) & ch.rmask)),
                          w4 = ch.lds32(ch.Rs + ((qb + 16u) & ch.rmask));
                B0 = __funnelshift_r(w0, w1, sh);
                B1 = __funnelshift_r(w1, w2, sh);
                B2 = __funnelshift_r(w2, w3, sh);
                B3 = __funnelshift_r(w3, w4, sh);
            }
            const u32 H = ch.hash(B0);
            const u32 t = V ? ch.tget(H) : ch.lo;
            const u32 mp = __match_any_sync(kFullMask, V ? H : (0x80000000u | lane));
            u32 m = 0;
            {
                const u32 nearp = (t >= ch.lo) ? 1u : 0u;
                const uintptr_t ga = reinterpret_cast<uintptr_t>(ch.F + t);
                const u32* g = reinterpret_cast<const u32*>(ga & ~(uintptr_t)3);
                const u32 tb = t & ~3u, tsh = (nearp ? t : (u32)ga) << 3;
                u32 c0, c1, c2 = 0, c3 = 0, c4 = 0;
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "setp.ne.u32 p, %5, 0;\n"
                    "@p ld.shared.u32 %0, [%6];\n"
                    "@p ld.shared.u32 %1, [%7];\n"
                    "@p ld.shared.u32 %2, [%8];\n"
                    "@p ld.shared.u32 %3, [%9];\n"
                    "@p ld.shared.u32 %4, [%10];\n"
                    "@!p ld.global.nc.u32 %0, [%11];\n"
                    "@!p ld.global.nc.u32 %1, [%11+4];\n"
                    "}\n"
                    : "=r"(c0), "=r"(c1), "+r"(c2), "+r"(c3), "+r"(c4)
                    : "r"(nearp), "r"(ch.Rs + (tb & ch.rmask)), "r"(ch.Rs + ((tb + 4u) & ch.rmask)),
                      "r"(ch.Rs + ((tb + 8u) & ch.rmask)), "r"(ch.Rs + ((tb + 12u) & ch.rmask)),
                      "r"(ch.Rs + ((tb + 16u) & ch.rmask)), "l"(g)
                    : "memory");
                const u32 C0 = __funnelshift_r(c0, c1, tsh);
                const bool more = V && !nearp && C0 == B0;
                if (__any_sync(kFullMask, more)) {
                    if (more) {
                        c2 = __ldg(g + 2);
                        c3 = __ldg(g + 3);
                        c4 = __ldg(g + 4);
                    }
                }
                const u32 x0 = C0 ^ B0, x1 = __funnelshift_r(c1, c2, tsh) ^ B1,
                          x2 = __funnelshift_r(c2, c3, tsh) ^ B2, x3 = __funnelshift_r(c3, c4, tsh) ^ B3;
                if (x0) m = ((u32)__ffs((int)x0) - 1u) >> 3;
                else if (x1) m = 4u + (((u32)__ffs((int)x1) - 1u) >> 3);
                else if (x2) m = 8u + (((u32)__ffs((int)x2) - 1u) >> 3);
                else if (x3) m = 12u + (((u32)__ffs((int)x3) - 1u) >> 3);
                else m = 16u;
                if (!V) m = 0;
            }
            const u32 vmask = __ballot_sync(kFullMask, V);
            const u32 hitmask = __ballot_sync(kFullMask, m >= 4u);
            const u32 dupmask = __ballot_sync(kFullMask, (mp & ((1u << lane) - 1u)) != 0u);
            const u32 tm = (t << 16) | (m << 8);
            // ------------- the chain passes through the windows in order
            if (k) mbar_wait_s(smem_u32(&ctl->tok[k]), tok_parity);
            u32 cur = k ? vc->cur : a;
            u32 scanning = k ? vc->scanning : (arrival ? 0u : 1u);
            u32 done = k ? vc->done : 0u;
            u32 fin = k ? vc->fin : 0u;
            if (k) {
                scan_s = vc->scan_s;
                lit_from = vc->lit_from;
            }
            if (!done && !fin && cur < base + 32u) {
                ch.op = vc->op;
                ch.nrec = vc->nrec;
                const u32 l0 = cur - base;
                u32 unsafe = dupmask;
                if (k) {
                    if (!scanning && l0 == 0) {  // :233 the position before the arrival lies in the window before
                        if (lane == 0) ch.tput(ch.hash(ch.ring32u(base - 1u)), base - 1u);
                        __syncwarp();
                    }
                    // trusted iff the table still holds what the lane looked up
                    unsafe |= __ballot_sync(kFullMask, V && ch.tget(H) != t);
                }
                // per-lane descriptor: what happens when the chain ARRIVES at this lane
                //   bits 0-2 kind, 3-7 lane e of the event, 8-12 copy length, 13 a scan started, 16-31 candidate
                const u32 stop_all = ~vmask | unsafe | hitmask;
                u32 desc, ins;
                {
                    const u32 lbit = 1u << lane;
                    const u32 rest = (lane < 31u) ? (stop_all >> (lane + 1u)) : 0u;
                    const u32 es = lane + (u32)__ffs((int)rest);
                    const u32 ebit = 1u << (es & 31u);
                    u32 kind, e;
                    ins = lbit | (lbit >> 1);  // :233,:235
                    if (hitmask & lbit) {
                        kind = W_COPY;
                        e = lane;
                    } else if (!rest) {
                        kind = W_LEAVE;
                        e = 0;
                        if (lane < 31u) ins |= ~0u << (lane + 1u);
                    } else {
                        e = es;
                        ins |= (ebit - 1u) & (~0u << (lane + 1u));
                        if (!(vmask & ebit)) kind = W_FIN;  // :175
                        else if (unsafe & ebit) kind = W_NEXTSCAN;
                        else {
                            kind = W_COPY;
                            ins |= ebit;  // :191
                        }
                    }
                    // an untrusted lane ends the round (lane 0 of warp 0 is always exact)
                    if ((unsafe & lbit) && (lane || k)) {
                        kind = W_NEXTARR;
                        e = lane;
                        ins = 0;
                    }
                    const u32 r = __shfl_sync(kFullMask, tm, e);
                    if (kind == W_COPY && ((r >> 8) & 31u) == 16u) kind = W_SLOW;
                    desc = kind | (e << 3) | ((hitmask & lbit) ? 0u : (1u << 13)) | (r & 0xffff1f00u);
                }
                u32 d, insacc, curl = l0;
                if (!scanning) {
                    d = __shfl_sync(kFullMask, desc, l0);
                    insacc = __shfl_sync(kFullMask, ins, l0);
                } else {  // entering inside a scan: it also stops where its probe count reaches 32 (:162-172)
                    const int klim = (int)(scan_s + 32u) - (int)base;  // first lane past the 32nd probe
                    const u32 limmask = klim >= 32 ? 0u : (klim <= 0 ? ~0u : ~((1u << klim) - 1u));
                    const u32 rest = (stop_all | limmask) >> l0;
                    u32 kind, e = 0;
                    if (!rest) {
                        kind = W_LEAVE;
                        insacc = ~0u << l0;
                    } else {
                        e = l0 + (u32)__ffs((int)rest) - 1u;
                        const u32 ebit = 1u << e;
                        insacc = (ebit - 1u) & (~0u << l0);
                        if (!(vmask & ebit)) kind = W_FIN;
                        else if ((unsafe | limmask) & ebit) kind = W_NEXTSCAN;
                        else {
                            kind = W_COPY;
                            insacc |= ebit;
                        }
                    }
                    const u32 r = __shfl_sync(kFullMask, tm, e);
                    if (kind == W_COPY && ((r >> 8) & 31u) == 16u) kind = W_SLOW;
                    d = kind | (e << 3) | (r & 0xffff1f00u);
                }
                bool beyond = false;
                for (;;) {
                    if (d & (1u << 13)) scan_s = base + curl + 1u;  // :162 a new scan started behind lane curl
                    const u32 kind = d & 7u;
                    if (kind > W_SLOW) break;
                    const u32 e = (d >> 3) & 31u;
                    u32 me = (d >> 8) & 31u;
                    if (kind == W_SLOW) me = ch.extend(base + e, d >> 16, 16);  // >= 16 bytes: the warp extends it
                    ch.keepq(lit_from, base + e, d >> 16, me);  // :200,:217
                    lit_from = base + e + me;
                    if ((int)lit_from >= lim) {  // :222
                        d = W_FIN;
                        break;
                    }
                    if (lit_from >= base + 32u) {
                        beyond = true;
                        break;
                    }
                    curl = lit_from - base;
                    d = __shfl_sync(kFullMask, desc, curl);
                    insacc |= __shfl_sync(kFullMask, ins, curl);
                }
                // commit the inserts of the path; the highest position wins (:191)
                if (((insacc >> lane) & 1u) && (mp & insacc & ~((2u << lane) - 1u)) == 0u) ch.tput(H, q);
                __syncwarp();
                const u32 kind = d & 7u, ev = (d >> 3) & 31u;
                if (beyond) {
                    cur = lit_from;
                    scanning = 0;
                } else if (kind == W_FIN) {
                    fin = 1;
                } else if (kind == W_LEAVE) {
                    cur = base + 32u;
                    scanning = 1;
                } else if (kind == W_NEXTSCAN) {
                    cur = base + ev;
                    scanning = 1;
                    done = 1;
                } else {  // W_NEXTARR at an untrusted lane
                    cur = base + ev;
                    scanning = 0;
                    done = 1;
                }
                if (lane == 0) {
                    vc->op = ch.op;
                    vc->nrec = ch.nrec;
                }
            }
            if (lane == 0) {
                if (k + 1 < kW) {
                    vc->cur = cur;
                    vc->scanning = scanning;
                    vc->done = done;
                    vc->fin = fin;
                    vc->scan_s = scan_s;
                    vc->lit_from = lit_from;
                    mbar_arrive(smem_u32(&ctl->tok[k + 1]));
                } else {  // last window: the next round starts where the chain stands
                    nx->a = cur;
                    nx->arrival = scanning ? 0u : 1u;
                    nx->scan_s = scan_s;
                    nx->lit_from = lit_from;
                    nx->fin = fin;
                }
            }
            tok_parity ^= 1u;
            chain_barrier(bar_id, bar_n);
        }
        // :242-248 pending records and the remainder literal
        if (k == 0) {
            ch.op = vc->op;
            ch.nrec = vc->nrec;
            if (ch.nrec) {
                __syncwarp();
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ch.r_lit), "=r"(ch.r_cpy) : "r"(ch.Qs + 8u * lane) : "memory");
            }
            ch.finish(lit_from);
            if (lane == 0) frag_sizes[frag] = ch.op;
        }
        chain_barrier(bar_id, bar_n);
    }
}

}  // namespace sb200
