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

    // src/internal.jl:252-287 -- tag, then the bytes.  lanes copy the literal cooperatively.
    __device__ __forceinline__ void literal(const u8* F, u32 from, u32 len) {
        u32 n = len - 1;
        u32 hdr;
        if (len < 60) {  // :271 (a 60-byte literal takes the 2-byte header, like the reference)
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
__device__ __forceinline__ void compress_fragment_serial(const u8* F, u16* T, const u32 n,
                                                         const u32 shift, FragEmitter& em) {
    const u32 lane = em.lane;
    const int lim = (int)n - 16;  // ip_limit, 0-based (src/internal.jl:131)
    u32 ip = 0, next_emit = 0;

    if (n >= kInputMargin) {
        for (;;) {
            // ---- scan for a 4-byte match, src/internal.jl:162-194
            u32 skip = 32;
            ip += 1;
            u32 next_ip = ip;
            u32 next_hash = (lds32u(F, ip) * kHashMul) >> shift;
            u32 cand = 0;
            bool bail = false;
            for (;;) {
                ip = next_ip;
                u32 h = next_hash;
                u32 between = skip >> 5;
                skip += between;
                next_ip = ip + between;
                if ((int)next_ip > lim) { bail = true; break; }  // :175
                next_hash = (lds32u(F, next_ip) * kHashMul) >> shift;
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
                u32 hp = (lds32u(F, ip - 1) * kHashMul) >> shift;
                u32 hc = (w * kHashMul) >> shift;
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

// K1b: batched pages -- one CTA (one warp) per independent stream (src/Snappy.jl:20-36 per page:
// own varint header, table sized from the page length).  Fragments of a page are compressed one
// after the other straight into the page's output slot, so no compaction pass is needed.
// Shared memory: frag_cap + kFragPad bytes of fragment, then table_cap u16 entries, then mbarrier.
__global__ void __launch_bounds__(32)
k_compress_pages(const u8* __restrict__ g_in, const u64* __restrict__ in_off,
                 const u32* __restrict__ in_size, u8* __restrict__ g_out,
                 const u64* __restrict__ out_off, u32* __restrict__ out_size, u32 frag_cap,
                 u32 table_cap) {
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
        reset_table(T, shift, lane);
        compress_fragment_serial(F, T, n, shift, em);
    }
    if (lane == 0) out_size[pg] = em.op;
}

// K2: exclusive scan of the fragment sizes (single CTA; nfrag <= 65536), plus the side index.
//   offsets[f]   : byte offset of fragment f behind `base` (base = varint header length, or 0)
//   offsets[nfrag] = base + total
__global__ void __launch_bounds__(1024)
k_scan_sizes(const u32* __restrict__ sizes, u32 nfrag, u64 base, u64* __restrict__ offsets) {
    __shared__ u64 warp_excl[32];
    __shared__ u64 carry_s;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = base;
    __syncthreads();
    for (u32 blk = 0; blk < nfrag; blk += 1024 * 4) {
        const u32 i0 = blk + tid * 4;
        u64 v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
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
        for (int k = 0; k < 4; k++) {
            if (i0 + k < nfrag) offsets[i0 + k] = ex;
            ex += v[k];
        }
        __syncthreads();
        if (tid == 1023) carry_s = ex;  // thread 1023 ends at carry + this block's total
        __syncthreads();
    }
    if (tid == 0) offsets[nfrag] = carry_s;
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
