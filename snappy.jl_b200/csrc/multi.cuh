// multi.cuh -- device side of the multi-GPU path: assembly of a stream from the
// runs every rank compressed, and its inverse.
//
// The reference writes the fragments of a stream one behind the other into one output vector
// (src/Snappy.jl:25-35).  Here rank r compresses a contiguous run of whole fragments of every stream; the ONLY
// global fact it lacks is where its bytes go: header + the byte counts of the ranks before it.  Those counts are
// exchanged once (ncclAllGather of one u64 per rank and stream, on the compute stream, no host round trip);
// k_assemble then stores every fragment straight into the OWNER's buffer through a peer-mapped pointer (NVLink
// stores; the owner of stream s is rank s mod world) and the fragment's offset into the owner's side index.  No
// staging copy, no all-to-all.  Uncompress is the mirror image: k_pull loads a rank's compressed range (and its slice
// of the side index) from the owner's buffer over NVLink into local memory, where the indexed decoder runs.
#pragma once
#include "common.cuh"

namespace sb200 {

// first fragment and fragment count of rank r's run of a stream of nfrag fragments (whole-fragment sharding:
// the first nfrag % world ranks hold one fragment more)
__host__ __device__ inline void shard_frags(u32 nfrag, u32 world, u32 r, u32& lo, u32& cnt) {
    const u32 base = nfrag / world, extra = nfrag % world;
    cnt = base + (r < extra ? 1u : 0u);
    lo = r * base + (r < extra ? r : extra);
}

// ---- compress side ------------------------------------------------------------------------------------
struct AsmDesc {
    u8* dst_stream;    // stream region in the owner's arena (peer-mapped or local)
    u64* dst_index;    // side-index region in the owner's arena
    u32 frag_begin;    // first fragment of this shard in the launch's numbering (scratch slots, sizes, rel)
    u32 nfrag;         // fragments of the shard (> 0)
    u32 frag_lo;       // the shard's first fragment inside its stream
    u32 stream_nfrag;  // fragments of the whole stream
    u32 rank;          // global rank that compressed the shard
    u32 stream;
    u32 hdr_len;       // bytes of the stream's varint header
    u32 last;          // 1: the shard ends the stream (writes index[stream_nfrag] = stream length)
};

// one CTA per shard: exclusive scan of its fragment sizes (offsets inside the shard's segment) and the segment's
// byte count into this rank's row of the size matrix M[world][nstreams]
__global__ void __launch_bounds__(1024)
k_scan_shards(const u32* __restrict__ sizes, const AsmDesc* __restrict__ descs, u64* __restrict__ rel,
              u64* __restrict__ M, u32 nstreams) {
    __shared__ u64 warp_tot[32];
    __shared__ u64 carry_s;
    const AsmDesc d = descs[blockIdx.x];
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (u32 blk = 0; blk < d.nfrag; blk += 1024) {
        const u32 i = blk + tid;
        const u64 v = i < d.nfrag ? sizes[d.frag_begin + i] : 0;
        u64 incl = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const u64 t = __shfl_up_sync(kFullMask, incl, k);
            if (lane >= (u32)k) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        u64 before = 0, total = 0;
#pragma unroll
        for (u32 w = 0; w < 32; w++) {
            if (w < wid) before += warp_tot[w];
            total += warp_tot[w];
        }
        if (i < d.nfrag) rel[d.frag_begin + i] = carry_s + before + incl - v;
        __syncthreads();
        if (tid == 0) carry_s += total;
        __syncthreads();
    }
    if (tid == 0) M[(size_t)d.rank * nstreams + d.stream] = carry_s;
}

// one CTA per fragment: the fragment's bytes from its scratch slot to header + (bytes of the ranks before) + (offset
// inside the segment) in the owner's stream region; destination-aligned 16-byte stores (as k_compact), which cross
// NVLink when the owner is another GPU.  Thread 0 also stores the fragment's stream offset into the owner's index.
// `rot`: CTAs are handed out in blockIdx order, i.e. stream after stream; every rank starts with a DIFFERENT stream
// (the one behind its own), so that at any moment the N ranks store into N different owners.  In stream order all of
// them wrote to owner 0 first, then owner 1, ...: one GPU's NVLink ingress at a time (measured at N = 8: 2.2-2.4 ms
// for 0.42 GB per rank, profiles/r02v_trace_n8.txt).
__global__ void __launch_bounds__(256)
k_assemble(const u8* __restrict__ scratch, const u32* __restrict__ sizes, const u64* __restrict__ rel,
           const u64* __restrict__ M, u32 nstreams, const AsmDesc* __restrict__ descs, u32 ndesc, u32 rot) {
    u32 frag = blockIdx.x + rot;
    if (frag >= gridDim.x) frag -= gridDim.x;
    u32 k = 0;
    while (k + 1 < ndesc && descs[k + 1].frag_begin <= frag) k++;
    const AsmDesc d = descs[k];
    u64 before = d.hdr_len;
    for (u32 r = 0; r < d.rank; r++) before += M[(size_t)r * nstreams + d.stream];
    const u32 c = sizes[frag];
    const u64 at = before + rel[frag];
    if (threadIdx.x == 0) {
        d.dst_index[d.frag_lo + (frag - d.frag_begin)] = at;
        if (d.last && frag == d.frag_begin + d.nfrag - 1) d.dst_index[d.stream_nfrag] = at + c;
    }
    const u8* src = scratch + (u64)frag * kSlotStride;
    u8* dst = d.dst_stream + at;
    u32 head = (u32)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    if (head > c) head = c;
    for (u32 i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    const u32 nvec = (c - head) >> 4;
    const u32 sh = (head & 3) * 8;
    const u32* sw = reinterpret_cast<const u32*>(src + (head & ~3u));
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (u32 j = threadIdx.x; j < nvec; j += blockDim.x) {
        const u32* p = sw + 4 * j;
        const u32 w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = p[4];
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh);
        v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh);
        v.w = __funnelshift_r(w3, w4, sh);
        d4[j] = v;
    }
    for (u32 i = head + (nvec << 4) + threadIdx.x; i < c; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();  // peer stores are out before the kernel counts as finished
}

// ---- uncompress side ----------------------------------------------------------------------------------
struct StreamMeta {   // written by the owner, MAX-reduced over the ranks (everybody else contributes zeros)
    u64 stream_len;   // bytes of the stream in the owner's arena
    u64 flags;        // != 0: the owner could not provide stream + index (not fragment-clean, too large ...)
};

struct PullDesc {
    const u8* src_stream;   // stream region in the owner's arena (peer-mapped or local)
    const u64* src_index;   // side-index region in the owner's arena
    u8* staging;            // local: the rank's compressed range lands here, 16-byte phase kept
    u64* rel;               // local: nfrag + 1 offsets into `staging`
    u64 staging_cap;
    u32 frag_lo, nfrag;     // the rank's run of the stream (nfrag > 0)
    u32 stream;
    u32 pad_;
};

// grid (tiles, shards): copy the compressed bytes of fragments [frag_lo, frag_lo + nfrag) of a stream, and the slice
// of its side index, from the owner's arena (peer loads over NVLink when it is another GPU).  The byte range is known
// on the device only (two index entries), so every CTA derives its tile from them.  Anything inconsistent leaves an
// all-zero index slice: the decoder then rejects the shard and raises its flag.
__global__ void __launch_bounds__(256)
k_pull(const PullDesc* __restrict__ descs, const StreamMeta* __restrict__ meta) {
    const PullDesc d = descs[blockIdx.y];
    const StreamMeta m = meta[d.stream];
    const u64 a0 = d.src_index[d.frag_lo], a1 = d.src_index[d.frag_lo + d.nfrag];
    const u64 base = a0 & ~(u64)15;  // source and destination keep the same 16-byte phase
    const bool ok = m.flags == 0 && a0 <= a1 && a1 <= m.stream_len && a1 - base + 16 <= d.staging_cap;
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (u32 i = tid; i <= d.nfrag; i += nthr) d.rel[i] = ok ? d.src_index[d.frag_lo + i] - base : 0;
    if (!ok) return;
    const u64 nvec = (a1 - base + 15) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(d.src_stream + base);
    uint4* d4 = reinterpret_cast<uint4*>(d.staging);
    for (u64 j = tid; j < nvec; j += nthr) d4[j] = s4[j];
}

}  // namespace sb200
