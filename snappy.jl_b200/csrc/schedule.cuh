// schedule.cuh -- fragment ORDER for the persistent compress warps (longest processing time first).
//
// The reference compresses the fragments of a stream one after the other (src/Snappy.jl:29-33); they are
// independent, so the order in which the GPU takes them changes nothing in the bytes.  It does change the time:
// the persistent warps of k_compress_window pull fragments from a counter, a text-like fragment costs ~20x an
// incompressible one, and whatever is pulled LAST decides how long the launch runs after the queue is empty
// (measured on the 1 GiB mix: ~10.2 ms of steady state + ~3 ms during which warps idle while the last expensive
// fragments finish).  So the fragments are handed out by decreasing estimated cost: the tail then consists of the
// cheap ones.
//
//   k_estimate_cost : one warp per fragment looks at its first kSampleBytes: every position's 4 bytes go through a
//                     small hash table (last position per bucket, races between lanes are harmless: it is an
//                     estimate); a position whose bucket holds an earlier position with the same 4 bytes is a
//                     "hit".  cost = number of hit RUNS (maximal runs of consecutive hit positions) ~ number of
//                     copies the compressor will emit ~ rounds of the window kernel.  Incompressible data: no
//                     hits; long runs / short-period records: few, long hit runs; text and dictionary data: many.
//   k_order_by_cost : counting sort of the fragment numbers by cost, descending (single CTA; <= 256 cost classes).
#pragma once
#include "compress_chain.cuh"

namespace sb200 {

constexpr u32 kSampleBytes = 4096;
constexpr u32 kCostBuckets = 1024;     // u16 entries per warp
constexpr u32 kCostWarps = 8;          // warps per CTA of k_estimate_cost
constexpr u32 kCostClasses = 256;

__global__ void __launch_bounds__(kCostWarps * 32)
k_estimate_cost(const u8* __restrict__ g_in, u64 shard_len, u32 nfrag, const ShardDesc* __restrict__ descs,
                u32 ndesc, u8* __restrict__ cost) {
    __shared__ u16 tabs[kCostWarps][kCostBuckets];
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 frag = blockIdx.x * kCostWarps + warp;
    if (frag >= nfrag) return;
    const u8* sbase = g_in;
    u64 slen = shard_len;
    u32 local = frag;
    if (descs) {
        u32 k = 0;
        while (k + 1 < ndesc && descs[k + 1].frag_begin <= frag) k++;
        sbase = descs[k].ptr;
        slen = descs[k].len;
        local = frag - descs[k].frag_begin;
    }
    const u64 start = (u64)local * kBlockSize;
    const u32 n = (u32)((slen - start < kBlockSize) ? (slen - start) : kBlockSize);
    const u8* F = sbase + start;
    u16* T = tabs[warp];
    for (u32 i = lane; i < kCostBuckets; i += 32) T[i] = 0;
    __syncwarp();
    // positions whose 8 bytes lie inside the fragment (the unaligned load touches the aligned words around it)
    const u32 span = n < 8 ? 0u : (n - 8 < kSampleBytes ? n - 8 : kSampleBytes);
    u32 runs = 0, carry = 0;
    for (u32 base = 0; base < span; base += 32) {
        const u32 q = base + lane;
        const bool in = q < span;
        u32 w = 0, h = 0, old = 0;
        if (in) {
            w = ldg32u(F + q);
            h = (w * kHashMul) >> 22;
            old = T[h];
        }
        __syncwarp();
        if (in) T[h] = (u16)(q + 1);
        const bool hit = in && old != 0 && ldg32u(F + old - 1) == w;
        const u32 hm = __ballot_sync(kFullMask, hit);
        runs += (u32)__popc(hm & ~((hm << 1) | carry));
        carry = hm >> 31;
        __syncwarp();
    }
    // scale short samples up to the full sample length, so that a ragged last fragment is ranked by its density
    if (span && span < kSampleBytes) runs = runs * kSampleBytes / span;
    if (lane == 0) cost[frag] = (u8)(runs > kCostClasses - 1 ? kCostClasses - 1 : runs);
}

// order[0 .. nfrag): fragment numbers by decreasing cost class (stable inside a class up to atomics order)
__global__ void __launch_bounds__(1024)
k_order_by_cost(const u8* __restrict__ cost, u32 nfrag, u32* __restrict__ order) {
    __shared__ u32 hist[kCostClasses];
    __shared__ u32 base[kCostClasses];
    const u32 tid = threadIdx.x;
    if (tid < kCostClasses) hist[tid] = 0;
    __syncthreads();
    for (u32 i = tid; i < nfrag; i += 1024) atomicAdd(&hist[cost[i]], 1u);
    __syncthreads();
    if (tid == 0) {
        u32 acc = 0;
        for (int c = (int)kCostClasses - 1; c >= 0; c--) {
            base[c] = acc;
            acc += hist[c];
        }
    }
    __syncthreads();
    for (u32 i = tid; i < nfrag; i += 1024) order[atomicAdd(&base[cost[i]], 1u)] = i;
}

}  // namespace sb200
