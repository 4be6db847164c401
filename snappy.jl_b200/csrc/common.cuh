// common.cuh -- shared device helpers for libsnappy_b200 (sm_100a only).
#pragma once
#ifdef SB200_CPU_EMU  // tools/cpu_warp: one warp of the compress kernels on the CPU (test infrastructure)
#include "cuda_shim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace sb200 {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 kBlockSize = 65536;        // K_BLOCK_SIZE, src/internal.jl:31
constexpr u32 kInputMargin = 15;         // K_INPUT_MARGIN_BYTES, src/internal.jl:32
constexpr u32 kMaxTableEntries = 16384;  // K_MAX_HASH_TABLE_SIZE, src/internal.jl:33
constexpr u32 kHashMul = 0x1e35a7bdu;    // hashdword, src/internal.jl:94
// per-fragment scratch slot: maxlength_compressed(65536) = 76490 rounded up to 128 B
constexpr u32 kSlotStride = 76544;
constexpr u32 kFullMask = 0xffffffffu;

// status codes mirrored from include/snappy_b200.h
enum : int {
    ST_OK = 0, ST_INPUT_TOO_LARGE = 1, ST_INVALID_INPUT = 2, ST_CORRUPT_COPY_OFFSET = 3,
    ST_CORRUPT_COPY_LENGTH = 4, ST_CORRUPT_LITERAL = 5, ST_BAD_VARINT = 6
};

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }

#ifdef SB200_CPU_EMU
extern u8 smem[];  // the CTA's dynamic shared memory; a "shared-space address" is an offset into it
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)(reinterpret_cast<const u8*>(p) - smem); }
#else
__device__ __forceinline__ u32 smem_u32(const void* p) {
    return (u32)__cvta_generic_to_shared(p);
}
#endif

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------------------------
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; src/dst 16 B aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, u32 bytes, u64* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// unaligned little-endian 32-bit load from shared memory (fastmemory.jl:4 load32u).
// Reads the two aligned words around q: the buffer needs >= 7 readable bytes past q.
__device__ __forceinline__ u32 lds32u(const u8* base, u32 q) {
    const u32* w = reinterpret_cast<const u32*>(base + (q & ~3u));
    return __funnelshift_r(w[0], w[1], (q & 3u) * 8u);
}

// dst[0 .. len) = src[0 .. len), non-overlapping (or src at least 16 bytes below dst): 16-byte aligned
// stores; the source is read as the aligned words that hold each 16-byte group and funnel-shifted, so
// the two misalignments need not agree.  kReadOnly: src is input (read-only path), else output bytes
// this kernel wrote earlier (coherent loads).
template <bool kReadOnly>
__device__ __forceinline__ void warp_copy_forward(u8* __restrict__ dst, const u8* src, u64 len, u32 lane) {
    if (len < 64) {
        for (u64 i = lane; i < len; i += 32) dst[i] = src[i];
        return;
    }
    const u64 head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    for (u64 i = lane; i < head; i += 32) dst[i] = src[i];
    const u64 nvec = (len - head) >> 4;
    const uintptr_t sa = reinterpret_cast<uintptr_t>(src + head);
    const u32* sw = reinterpret_cast<const u32*>(sa & ~(uintptr_t)3);
    const u32 sh = (u32)sa << 3;
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (u64 j = lane; j < nvec; j += 32) {
        const u32* p = sw + 4 * j;
        u32 w0, w1, w2, w3, w4;
        if (kReadOnly) {
            w0 = __ldg(p); w1 = __ldg(p + 1); w2 = __ldg(p + 2); w3 = __ldg(p + 3);
            w4 = (sh & 31u) ? __ldg(p + 4) : 0u;  // the fifth word only when the source is misaligned
        } else {
            w0 = p[0]; w1 = p[1]; w2 = p[2]; w3 = p[3];
            w4 = (sh & 31u) ? p[4] : 0u;
        }
        uint4 v;
        v.x = __funnelshift_r(w0, w1, sh);
        v.y = __funnelshift_r(w1, w2, sh);
        v.z = __funnelshift_r(w2, w3, sh);
        v.w = __funnelshift_r(w3, w4, sh);
        d4[j] = v;
    }
    for (u64 i = head + (nvec << 4) + lane; i < len; i += 32) dst[i] = src[i];
}

}  // namespace sb200
