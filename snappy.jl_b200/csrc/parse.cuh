// parse.cuh -- segmented speculative parse of an arbitrary Snappy stream (no side index).
//
// The tag chain of a stream is serial: each element's position depends on every earlier one
// (loop header of decompress_all_tags!, src/internal.jl:416-439).  It is cut into fixed chunks of
// compressed bytes, one THREAD per chunk, and recovered in a fixed number of parallel passes:
//
//  A  k_parse_guess   every chunk guesses its entry by walking from a look-back position (tag chains
//                     started at different bytes merge within a few elements) and parses itself:
//                     first[k] = first visited position >= chunk start, exit[k] = first visited
//                     position >= chunk end.  Chunks inside long literals produce garbage; that
//                     is harmless because nothing below trusts a chunk before it is REACHED.
//  B  k_parse_bridge  for every chunk k: where does the chain go if exit[k] is a real element
//                     start?  Normally exit[k] == first[j] of the chunk j it lands in (join).
//                     Otherwise (a long literal ended inside j and j's guess started in its bytes)
//                     the thread walks on from exit[k] ("bridge", at most kBridgeBudget elements;
//                     a 64 KiB literal is one element) until its first element start in some
//                     chunk equals that chunk's first[].  next[k] = joined chunk.
//  C  k_parse_reach   chunk 0 starts exactly behind the varint, so it is real; pointer doubling
//                     over next[] marks every chunk the real chain reaches (log2(nchunk) rounds).
//  D  k_parse_entries real entry of a reached chunk = first[k]; reached chunks re-walk their bridge
//                     and give every chunk it enters its real entry; all other chunks hold no
//                     element start.
//  E  k_parse_final   every real chunk is parsed once more from its real entry: output bytes,
//                     anomaly flags.  A scan gives output offsets, k_build_index records the
//                     compressed position of every 64 KiB output boundary.
//
// Segments (streamed host-buffer uncompress): the same passes run over a SEGMENT [hdr, E) of the
// stream whose first byte is a known element start (hdr) and whose elements may run past E (L stays
// the true end of the stream): a chain that leaves the segment ends at the first element start
// >= E, which is reported (counters64[1]) and becomes the next segment's hdr.
//
// If the stream is "fragment-clean" (no element straddles, no copy reaches across a 64 KiB output
// boundary -- true for everything Snappy.jl emits) the result is the side index the compressor
// would have produced and the indexed decoder runs; it re-validates every fragment, so a wrong
// index can only cost time, never change the result.  Anything else goes to the exact decoder.
#pragma once
#include "common.cuh"
#include "decompress.cuh"

namespace sb200 {

constexpr u32 kParseChunkLog2 = 10;   // default: 1 KiB of compressed bytes per chunk (one thread each)
#ifndef SB200_PARSE_LOOKBACK
#define SB200_PARSE_LOOKBACK 256
#endif
// guess walk starts this far before the chunk (1024 / 512 / 256: 8.22 / 7.85 / 7.67 ms per step of config 3,
// profiles/r02u_parse_lookback.txt; a guess that has not joined the real chain by the chunk start costs a bridge walk,
// never a wrong result)
constexpr u32 kParseLookback = SB200_PARSE_LOOKBACK;
constexpr u32 kParseThreads = 128;
constexpr u32 kBridgeBudget = 20000;  // elements a bridge may walk before it gives up
constexpr u64 kDeadPos = ~0ull;
constexpr u32 kNextDead = 0xffffffffu;  // chain cannot be followed from here
constexpr u32 kOutbUnknown = 0xffffffffu;  // outb[] after the guess pass: the walk failed

enum : u32 { PF_ANOMALY = 1u, PF_NOT_CLEAN = 2u, PF_BROKEN = 4u, PF_STRADDLE = 8u };

struct ParseArrays {
    u64* first;         // [nchunk]
    u64* exit;          // [nchunk]
    u64* entry;         // [nchunk] real entry (kDeadPos: chunk holds no element start)
    u32* next_a;        // [nchunk] successor chunk (nchunk = end of stream, kNextDead = dead)
    u32* next_b;        // [nchunk] double buffer for pointer doubling
    u32* reach;         // [nchunk]
    u32* outb;          // [nchunk] output bytes of the chunk's real elements
    u32* counters;      // [0] OR of PF_* flags
    static size_t bytes(size_t nchunk) {
        return nchunk * (8 * 3 + 4 * 4) + 64;
    }
    void carve(void* base, size_t nchunk) {
        u64* p = (u64*)base;
        first = p; p += nchunk;
        exit = p; p += nchunk;
        entry = p; p += nchunk;
        u32* q = (u32*)p;
        next_a = q; q += nchunk;
        next_b = q; q += nchunk;
        reach = q; q += nchunk;
        outb = q; q += nchunk;
        counters = q;
    }
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

// walk from ip until the chunk end (or the end of the stream, src/internal.jl:416: `ip + 1 < L`);
// returns false on an anomaly.  produced accumulates element lengths.
__device__ __forceinline__ bool walk_chunk(const u8* __restrict__ in, u64 L, u64 end, u64& ip, u64& produced) {
    Element e;
    while (ip < end && ip + 1 < L) {
        if (!walk_step(in, L, ip, e)) return false;
        produced += e.len;
    }
    return true;
}

// chunk k of the segment [hdr, E)
__device__ __forceinline__ void chunk_range(u64 hdr, u64 E, u32 k, u32 pshift, u64& start, u64& end) {
    start = hdr + ((u64)k << pshift);
    end = (start + (1ull << pshift) < E) ? (start + (1ull << pshift)) : E;
}

__global__ void __launch_bounds__(kParseThreads)
k_parse_guess(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa, u64 E, u32 pshift) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    u64 start, end;
    chunk_range(hdr, E, k, pshift, start, end);
    u64 ip = hdr;
    Element e;
    if (k > 0) {
        ip = (start - hdr > kParseLookback) ? (start - kParseLookback) : hdr;
        while (ip < start && ip + 1 < L) {
            if (!walk_step(in, L, ip, e)) { ip = start; break; }  // garbage: any guess will do
        }
        if (ip < start) ip = start;
    }
    pa.first[k] = ip;
    u64 produced = 0;
    const bool ok = walk_chunk(in, L, end, ip, produced);
    pa.exit[k] = ok ? ip : kDeadPos;
    pa.reach[k] = (k == 0) ? 1u : 0u;
    // output bytes of the chunk as walked from the guess: k_parse_final keeps them when the guess turns out to be the
    // chunk's real entry (nearly always), instead of walking the chunk once more
    pa.outb[k] = (ok && produced <= 0xfffffffeull) ? (u32)produced : kOutbUnknown;
}

__global__ void __launch_bounds__(kParseThreads)
k_parse_bridge(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa, u64 E, u32 pshift) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    u64 x = pa.exit[k];
    u32 nx = kNextDead, prev = kNextDead;
    Element e;
    pa.entry[k] = kDeadPos;  // until k_parse_entries: where a chain that leaves the segment through k ends
    for (u32 steps = 0; x != kDeadPos && steps < kBridgeBudget; steps++) {
        if (x + 1 >= L || x >= E) {  // the chain ends here (a lone trailing byte is ignored) or leaves the segment
            nx = nchunk;
            pa.entry[k] = x;
            break;
        }
        const u32 j = (u32)((x - hdr) >> pshift);
        if (j != prev) {  // x is this chain's first element start in chunk j
            if (pa.first[j] == x) { nx = j; break; }
            prev = j;
        }
        if (!walk_step(in, L, x, e)) break;
    }
    pa.next_a[k] = nx;
}

// one round of pointer doubling: reach spreads over `nx`, then nx_out = nx o nx
__global__ void __launch_bounds__(256)
k_parse_reach(u32 nchunk, const u32* __restrict__ nx, u32* __restrict__ nx_out, u32* reach) {
    const u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k >= nchunk) return;
    const u32 j = nx[k];
    if (j < nchunk) {
        if (reach[k]) reach[j] = 1u;
        nx_out[k] = nx[j];
    } else {
        nx_out[k] = j;
    }
}

// entry[k] of the chunks the real chain touches.  phase 0: reached chunks start at first[k].
// phase 1: every reached chunk re-walks its bridge (exit[k] .. the join) and gives each chunk the
// bridge enters its real entry.
__global__ void __launch_bounds__(kParseThreads)
k_parse_entries(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa,
                const u32* __restrict__ next_orig, int phase, u64 E, u32 pshift) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    if (phase == 0) {
        // the reached chunk whose chain leaves the segment reports where (exactly one such chunk)
        if (pa.reach[k] && next_orig[k] == nchunk) reinterpret_cast<u64*>(pa.counters)[1] = pa.entry[k];
        pa.entry[k] = pa.reach[k] ? pa.first[k] : kDeadPos;
        if (pa.reach[k] && next_orig[k] == kNextDead) atomicOr(&pa.counters[0], PF_BROKEN);
        return;
    }
    if (!pa.reach[k] || next_orig[k] == kNextDead) return;
    u64 x = pa.exit[k];
    u32 prev = kNextDead;
    Element e;
    for (u32 steps = 0; steps < kBridgeBudget; steps++) {
        if (x + 1 >= L || x >= E) break;
        const u32 j = (u32)((x - hdr) >> pshift);
        if (j != prev) {
            if (pa.first[j] == x) break;  // joined: phase 0 already set entry[j]
            pa.entry[j] = x;
            prev = j;
        }
        if (!walk_step(in, L, x, e)) break;
    }
}

__global__ void __launch_bounds__(kParseThreads)
k_parse_final(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa, u64 E, u32 pshift) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    u64 ip = pa.entry[k];
    u64 produced = 0;
    if (ip != kDeadPos && ip == pa.first[k] && pa.outb[k] != kOutbUnknown) return;  // the guess pass walked exactly this
    if (ip != kDeadPos) {
        u64 start, end;
        chunk_range(hdr, E, k, pshift, start, end);
        if (!walk_chunk(in, L, end, ip, produced) || produced > 0xffffffffull) {
            atomicOr(&pa.counters[0], PF_ANOMALY);
            produced = 0;
        }
    }
    pa.outb[k] = (u32)produced;
}

// Walk each real chunk again with its output offset known and record where every TILE of the output begins: tile k
// starts at the element that covers output byte k * 65536 (index[k] = its position in the stream, out_start[k] = its
// output offset).  Everything Snappy.jl, libsnappy and Google snappy emit starts a new element exactly on every
// 64 KiB boundary (they compress 64 KiB blocks independently), so out_start[k] == k * 65536 there.  `relaxed`: a
// LITERAL may straddle a boundary (streams of encoders that merge the literals of neighbouring blocks, e.g.
// tests/data/alice29.snappy); the tile then begins a few bytes early, at that literal.  A copy that straddles, or one
// that reaches back over the boundary of its tile, makes the tile depend on earlier output: flagged PF_NOT_CLEAN (the
// indexed decoder rejects such a tile anyway and the bounded serial walk decides).
__global__ void __launch_bounds__(kParseThreads)
k_build_index(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa,
              const u64* __restrict__ out_off, u64* __restrict__ index, u32 nfrag, u64 E, u64 out_base, u32 pshift,
              u64* __restrict__ out_start = nullptr, u32 relaxed = 0, u64 out_total = 0) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    if (k == 0 && E >= L) index[nfrag] = L;
    u64 ip = pa.entry[k];
    if (ip == kDeadPos) return;
    u64 start, end;
    chunk_range(hdr, E, k, pshift, start, end);
    u64 op = out_base + out_off[k];  // out_base: output bytes of the segments before this one
    bool clean = true, straddle = false;
    Element e;
    while (ip < end && ip + 1 < L) {
        const u64 at = ip;
        if (!walk_step(in, L, ip, e)) { clean = false; break; }
        const u32 in_frag = (u32)(op & (kBlockSize - 1));
        if (in_frag == 0) {
            if ((op >> 16) < nfrag) {
                index[op >> 16] = at;
                if (out_start) out_start[op >> 16] = op;
            }
            if ((u64)e.len > kBlockSize) {  // a literal longer than a tile: the tiles it swallows are empty
                if (relaxed && out_start && !e.is_copy) {
                    straddle = true;
                    for (u64 t = (op >> 16) + 1; (t << 16) < op + e.len && t < nfrag; t++) {
                        index[t] = at;
                        out_start[t] = op;
                    }
                } else {
                    clean = false;
                }
            }
        } else if (in_frag + (u64)e.len > kBlockSize) {  // the element straddles (at least) one boundary
            if (relaxed && out_start && !e.is_copy) {
                straddle = true;
                for (u64 t = (op >> 16) + 1; (t << 16) < op + e.len && t < nfrag; t++) {
                    index[t] = at;
                    out_start[t] = op;
                }
            } else {
                clean = false;
            }
        }
        if (e.is_copy && e.offset > in_frag) clean = false;
        if (relaxed) {
            // The reference's per-element checks depend on the element header and the output position only, never on
            // decoded bytes: offset == 0 or beyond what is produced (src/internal.jl:499), more bytes than the claimed
            // length leaves room for (:505, :518).  So the status of a corrupt stream falls out of the parse: the
            // failing element with the smallest stream position wins (the reference stops at the first one).
            u32 err = 0;
            const u64 room = out_total >= op ? out_total - op : 0;
            if (e.is_copy) {
                if (e.offset == 0 || (u64)e.offset > op) err = ST_CORRUPT_COPY_OFFSET;
                else if (room < e.len) err = ST_CORRUPT_COPY_LENGTH;
            } else if (room < (u64)e.len) {
                err = ST_CORRUPT_LITERAL;
            }
            if (err) {
                atomicMin(reinterpret_cast<unsigned long long*>(pa.counters) + 2, (unsigned long long)((at << 3) | err));
                break;  // nothing behind the first failing element of this chunk matters
            }
        }
        op += e.len;
    }
    if (!clean) atomicOr(&pa.counters[0], PF_NOT_CLEAN);
    if (straddle) atomicOr(&pa.counters[0], PF_STRADDLE);
}

// ---- clean cuts: tiles for streams whose copies reach across the 64 KiB output boundaries ------------------------
// An element start at output position p is a CLEAN CUT iff no element at or behind p reads output below p: then
// everything from p on can be decoded without what lies before it.  With reach(e) = op(e) - offset(e) for a copy e
// and R(p) = min reach over the copies at positions >= p (a suffix minimum, so R only falls as p falls):
// p is clean iff R(p) >= p, and if it is not, every cut in (R(p), p] is unclean as well (its suffix contains the
// same low copy).  So the greatest clean cut at or below a boundary B is found by jumping back: p = greatest element
// start <= B; while R(p) < p: p = greatest element start <= R(p).  It ends at p = 0 at the latest.
// Tile f of the re-tiled index starts at the greatest clean cut <= f * 65536 (monotone in f; a tile may be empty,
// its neighbour then spans more than 64 KiB), so a stream made of blocks that do not tile 64 KiB -- or one with a few
// long-range copies -- still decodes tile-parallel; only a stream whose copies chain across every boundary ends up
// as one long tile on one warp.
constexpr u32 kCutIters = 256;
constexpr u64 kCutNone = ~0ull - 1ull;

// low[k] = min reach of the copies that start in chunk k (~0: none)
__global__ void __launch_bounds__(kParseThreads)
k_cut_low(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa, const u64* __restrict__ out_off,
          u64 E, u32 pshift, u64* __restrict__ low) {
    const u32 k = blockIdx.x * kParseThreads + threadIdx.x;
    if (k >= nchunk) return;
    u64 m = ~0ull;
    u64 ip = pa.entry[k];
    if (ip != kDeadPos) {
        u64 start, end;
        chunk_range(hdr, E, k, pshift, start, end);
        u64 op = out_off[k];
        Element e;
        while (ip < end && ip + 1 < L) {
            if (!walk_step(in, L, ip, e)) break;
            if (e.is_copy) {
                const u64 r = (u64)e.offset <= op ? op - e.offset : 0;
                m = r < m ? r : m;
            }
            op += e.len;
        }
    }
    low[k] = m;
}

// sfx[k] = min(low[k + 1 ..]) (one CTA, from the end)
__global__ void __launch_bounds__(1024)
k_suffix_min(const u64* __restrict__ low, u32 n, u64* __restrict__ sfx) {
    __shared__ u64 warp_min[32];
    __shared__ u64 carry_s;
    const u32 tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = ~0ull;
    __syncthreads();
    for (u32 done = 0; done < n; done += 1024) {
        // thread t takes element i = n - 1 - (done + t): an inclusive scan over t is a suffix minimum over i
        const u32 t = done + tid;
        const bool in_range = t < n;
        const u32 i = in_range ? n - 1 - t : 0;
        const u64 v = in_range ? low[i] : ~0ull;
        u64 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 o = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= (u32)d) incl = o < incl ? o : incl;
        }
        if (lane == 31) warp_min[wid] = incl;
        __syncthreads();
        u64 before = carry_s, total = carry_s;
#pragma unroll
        for (u32 w = 0; w < 32; w++) {
            if (w < wid) before = warp_min[w] < before ? warp_min[w] : before;
            total = warp_min[w] < total ? warp_min[w] : total;
        }
        // exclusive: everything behind i, i.e. the threads before t and the blocks before this one
        const u64 up = __shfl_up_sync(kFullMask, incl, 1);
        u64 excl = before;
        if (lane > 0) excl = up < excl ? up : excl;
        if (in_range) sfx[i] = excl;
        __syncthreads();
        if (tid == 0) carry_s = total;
        __syncthreads();
    }
}

// one thread per tile f in [1, nfrag): the greatest clean cut <= f * 65536 (kCutNone: gave up, k_cut_fill decides)
__global__ void __launch_bounds__(kParseThreads)
k_cut_tiles(const u8* __restrict__ in, u64 L, u64 hdr, u32 nchunk, ParseArrays pa, const u64* __restrict__ out_off,
            const u64* __restrict__ sfx, u64 E, u32 pshift, u32 nfrag, u64* __restrict__ index,
            u64* __restrict__ out_start) {
    const u32 f = 1 + blockIdx.x * kParseThreads + threadIdx.x;
    if (f >= nfrag) return;
    u64 x = (u64)f << 16;
    for (u32 iter = 0; iter < kCutIters; iter++) {
        // the chunk whose elements cover output byte x: out_off[k] <= x < out_off[k + 1]
        u32 lo = 0, hi = nchunk;  // invariant: out_off[lo] <= x, out_off[hi] > x (out_off[nchunk] = total > x)
        while (hi - lo > 1) {
            const u32 mid = lo + ((hi - lo) >> 1);
            if (out_off[mid] <= x) lo = mid;
            else hi = mid;
        }
        const u32 k = lo;
        u64 ip = pa.entry[k];
        if (ip == kDeadPos) break;  // cannot happen for a complete parse (the covering chunk has output)
        u64 start, end;
        chunk_range(hdr, E, k, pshift, start, end);
        u64 op = out_off[k], p_op = op, p_ip = ip, m = ~0ull;
        Element e;
        while (ip < end && ip + 1 < L) {
            const u64 at = ip;
            if (!walk_step(in, L, ip, e)) break;
            if (op <= x) {  // a later element start at or below x: the suffix starts again here
                p_op = op;
                p_ip = at;
                m = ~0ull;
            }
            if (e.is_copy) {
                const u64 r = (u64)e.offset <= op ? op - e.offset : 0;
                m = r < m ? r : m;
            }
            op += e.len;
        }
        const u64 s = sfx[k];
        const u64 R = s < m ? s : m;
        if (R >= p_op) {
            index[f] = p_ip;
            out_start[f] = p_op;
            return;
        }
        x = R;
    }
    out_start[f] = kCutNone;
}

// tiles that gave up take their predecessor's start: the predecessor becomes empty, this tile spans both
__global__ void k_cut_fill(u32 nfrag, u64* __restrict__ index, u64* __restrict__ out_start) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (u32 f = 1; f < nfrag; f++)
        if (out_start[f] == kCutNone) {
            out_start[f] = out_start[f - 1];
            index[f] = index[f - 1];
        }
}

// first tile whose index entry was never written (the chain broke before it): everything below is trustworthy
__global__ void __launch_bounds__(256)
k_first_missing(const u64* __restrict__ index, u32 nfrag, u64 L, u32* __restrict__ first) {
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i <= nfrag && index[i] > L) atomicMin(first, i);
}

// flags, segment exit and output bytes straight into pinned host memory: a device-to-host copy would
// queue behind the output copies of the streamed path on the copy engine
__global__ void k_parse_report(const u32* __restrict__ counters, const u64* __restrict__ total, u64* host3) {
    host3[0] = counters[0];
    host3[1] = reinterpret_cast<const u64*>(counters)[1];
    host3[2] = *total;
    host3[3] = reinterpret_cast<const u64*>(counters)[2];  // (stream position << 3) | status of the first failing element
    __threadfence_system();
}

__global__ void __launch_bounds__(256)
k_max_u32(const u32* __restrict__ v, u32 count, u32* __restrict__ result) {
    u32 m = 0;
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < count; i += gridDim.x * 256) m = max(m, v[i]);
    m = __reduce_max_sync(kFullMask, m);
    if ((threadIdx.x & 31) == 0) atomicMax(result, m);
}

}  // namespace sb200
