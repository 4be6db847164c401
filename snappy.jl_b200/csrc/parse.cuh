// parse.cuh -- segmented speculative parse of an arbitrary Snappy stream (no side index).
//
// The tag chain of a stream is serial (each element's position depends on every earlier one).
// The stream is cut into fixed chunks of compressed bytes, one THREAD per chunk.  Round 0 guesses
// each chunk's entry (first element start at or after the chunk start) by walking from a
// look-back position -- tag chains started at different bytes merge quickly -- and parses the chunk
// from the guess.  k_link_chunks then compares every entry with the predecessor's exit; chunks
// whose entry was wrong are re-parsed, until nothing changes (at the fixpoint entry[0] is exact and
// entry[k+1] == exit[k], i.e. the chain is the true one no matter how bad the guesses were).
// An exclusive scan of the per-chunk output sizes gives each chunk's output offset, and
// k_build_index records the compressed position of every 64 KiB output boundary.  If the stream is
// "fragment-clean" (no element straddles, no copy reaches across a 64 KiB output boundary -- true
// for everything Snappy.jl and libsnappy emit) the result is the same side index the compressor
// produces, and the indexed decoder runs.  Anything else is left to the exact serial decoder.
#pragma once
#include "common.cuh"
#include "decompress.cuh"

namespace sb200 {

constexpr u32 kParseChunk = 4096;     // compressed bytes per chunk
constexpr u32 kParseLookback = 1024;  // guess walk starts this far before the chunk
constexpr u32 kParseThreads = 128;

enum : u32 { PF_ANOMALY = 1u, PF_NOT_CLEAN = 2u, PF_BAD_END = 4u, PF_BAD_TOTAL = 8u };

struct ParseArrays {
    u64* entry;   // [nchunk] first element start >= chunk start (under the current chain)
    u64* exit;    // [nchunk] first element start >= chunk end (== entry of the next chunk)
    u32* outb;    // [nchunk] output bytes produced by the chunk's elements
    u32* flags;   // [nchunk] PF_* seen while parsing from `entry`
    u32* dirty;   // [nchunk] needs a re-parse
    u32* counters;  // [0] changes made by k_link_chunks, [1] OR of flags, [2] not-clean flag
};

// one element at ip: returns false on an anomaly (header or literal past the end, empty element)
__device__ __forceinline__ bool walk_step(const u8* __restrict__ in, u64 L, u64& ip, Element& e) {
    u32 c, tag4;
    load_tag(in, ip, L, c, tag4);
    e = decode_tag(c, tag4);
    u64 nx = ip + 1 + e.extra;
    if (nx > L || e.len == 0) return false;
    if (!e.is_copy) {
        if ((u64)e.len > L - nx) return false;
        nx += e.len;
    }
    ip = nx;
    return true;
}

__global__ void __launch_bounds__(kParseThreads)
k_parse_chunks(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa, int first_round) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    const u64 start = hdr + (u64)k * kParseChunk;
    const u64 end = (start + kParseChunk < L) ? (start + kParseChunk) : L;
    u64 ip;
    Element e;
    if (first_round) {
        ip = hdr;
        if (k > 0) {
            ip = (start - hdr > kParseLookback) ? (start - kParseLookback) : hdr;
            while (ip < start && ip + 1 < L) {
                if (!walk_step(in, L, ip, e)) { ip = start; break; }  // garbage: any guess will do
            }
            if (ip < start) ip = start;
        }
        pa.entry[k] = ip;
    } else {
        if (!pa.dirty[k]) return;
        pa.dirty[k] = 0;
        ip = pa.entry[k];
    }
    u64 produced = 0;
    u32 flags = 0;
    while (ip < end && ip + 1 < L) {  // `ip + 1 < L`: src/internal.jl:416
        if (!walk_step(in, L, ip, e)) {
            flags |= PF_ANOMALY;
            ip = L;
            break;
        }
        produced += e.len;
    }
    if (produced > 0xffffffffull) {
        flags |= PF_ANOMALY;
        produced = 0xffffffffull;
    }
    pa.exit[k] = ip;
    pa.outb[k] = (u32)produced;
    pa.flags[k] = flags;
}

// entry[k] must equal exit[k-1]; chunks that the predecessor's last element jumps over entirely
// (long literals) are resolved on the spot.  counters[0] counts the changes.
__global__ void __launch_bounds__(256)
k_link_chunks(u64 L, u64 hdr, u32 nchunk, ParseArrays pa) {
    const u32 k = blockIdx.x * 256 + threadIdx.x + 1;
    if (k >= nchunk) return;
    const u64 v = pa.exit[k - 1];
    if (pa.entry[k] == v) return;
    pa.entry[k] = v;
    const u64 start = hdr + (u64)k * kParseChunk;
    const u64 end = (start + kParseChunk < L) ? (start + kParseChunk) : L;
    if (v >= end) {
        pa.exit[k] = v;
        pa.outb[k] = 0;
        pa.flags[k] = 0;
        pa.dirty[k] = 0;
    } else {
        pa.dirty[k] = 1;
    }
    atomicAdd(&pa.counters[0], 1u);
}

// OR of the per-chunk flags, and the chain must end exactly at L (a trailing ignored byte or a
// truncated header is the exact decoder's business)
__global__ void __launch_bounds__(256)
k_parse_check(u64 L, u32 nchunk, ParseArrays pa) {
    u32 f = 0;
    for (u32 k = blockIdx.x * 256 + threadIdx.x; k < nchunk; k += gridDim.x * 256) f |= pa.flags[k];
    if (blockIdx.x == 0 && threadIdx.x == 0 && pa.exit[nchunk - 1] != L) f |= PF_BAD_END;
    f = __reduce_or_sync(kFullMask, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(&pa.counters[1], f);
}

// Walk each chunk again with its output offset known; record the compressed position of every
// element that starts exactly on a 64 KiB output boundary; flag anything not fragment-clean.
__global__ void __launch_bounds__(kParseThreads)
k_build_index(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa,
              const u64* __restrict__ out_off, u64* __restrict__ index, u32 nfrag) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    const u64 start = hdr + (u64)k * kParseChunk;
    const u64 end = (start + kParseChunk < L) ? (start + kParseChunk) : L;
    u64 ip = pa.entry[k];
    u64 op = out_off[k];
    bool clean = true;
    Element e;
    while (ip < end && ip + 1 < L) {
        const u64 at = ip;
        if (!walk_step(in, L, ip, e)) { clean = false; break; }
        const u32 in_frag = (u32)(op & (kBlockSize - 1));
        if (in_frag == 0 && (op >> 16) < nfrag) index[op >> 16] = at;
        if (in_frag + (u64)e.len > kBlockSize) clean = false;
        if (e.is_copy && e.offset > in_frag) clean = false;
        op += e.len;
    }
    if (!clean) atomicOr(&pa.counters[2], 1u);
    if (k == 0) index[nfrag] = L;
}

__global__ void __launch_bounds__(256)
k_max_u32(const u32* __restrict__ v, u32 count, u32* __restrict__ result) {
    u32 m = 0;
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < count; i += gridDim.x * 256) m = max(m, v[i]);
    m = __reduce_max_sync(kFullMask, m);
    if ((threadIdx.x & 31) == 0) atomicMax(result, m);
}

}  // namespace sb200
