// run_decode_kernel.cpp -- runs the REAL decoder source (snappy.jl_b200/csrc/decompress.cuh) as one warp on the CPU
// (cuda_shim.h).  TEST INFRASTRUCTURE ONLY (tests/test_kernel_on_cpu_warp.py).
//   mode "indexed": every file is compressed with the oracle (fragment sizes = the side index), then each fragment
//                   goes through k_decode_fragments (the fused window-parallel decoder) and must give the file back;
//   mode "exact":   every file is taken as a STREAM (e.g. the reference's baddata*.snappy) and decoded by
//                   k_decode_serial; status, produced bytes and the bytes themselves must equal the oracle's
//                   decoder (sjo_uncompress_ex), error cases included.
//   g++ -O1 -std=c++17 -DSB200_CPU_EMU -Itools/cpu_warp tools/cpu_warp/run_decode_kernel.cpp oracle.o
//   ./a.out indexed|exact file...
#include "../../snappy.jl_b200/csrc/decompress.cuh"

extern "C" {
#include "../../oracle/snappy_oracle.h"
}

namespace sb200 {
u8 smem[1024] __attribute__((aligned(128)));
}
using namespace sb200;

struct IndexedArgs {
    const u8* in;
    const u64* off;
    u32 nfrag;
    u64 in_begin, in_end;
    u8* out;
    u64 out_len;
    DecodeResult* res;
};
static void entry_indexed(void* p) {
    const IndexedArgs& a = *(const IndexedArgs*)p;
    k_decode_fragments<12>(a.in, a.off, a.nfrag, 0u, a.nfrag, a.in_begin, a.in_end, a.out, a.out_len, a.res, nullptr, 0u);
}
struct ExactArgs {
    const u8* in;
    u64 L, ip0;
    u8* out;
    u64 n;
    DecodeResult* res;
};
static void entry_exact(void* p) {
    const ExactArgs& a = *(const ExactArgs*)p;
    k_decode_serial(a.in, a.L, a.ip0, a.out, a.n, a.res);
}

static u8* slurp(const char* path, size_t* sz, size_t pad) {
    FILE* fp = fopen(path, "rb");
    if (!fp) {
        perror(path);
        exit(2);
    }
    fseek(fp, 0, SEEK_END);
    *sz = (size_t)ftell(fp);
    fseek(fp, 0, SEEK_SET);
    u8* b = (u8*)aligned_alloc(256, (*sz + pad + 255) & ~(size_t)255);
    memset(b, 0, *sz + pad);
    if (fread(b, 1, *sz, fp) != *sz) exit(2);
    fclose(fp);
    return b;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const bool exact = !strcmp(argv[1], "exact");
    int failed = 0;
    for (int ai = 2; ai < argc; ai++) {
        size_t sz;
        u8* raw = slurp(argv[ai], &sz, 256);
        if (!exact) {
            const u32 nfrag = (u32)((sz + kBlockSize - 1) / kBlockSize);
            if (!nfrag) continue;
            u8* comp = (u8*)aligned_alloc(256, ((size_t)nfrag * kSlotStride + 512 + 255) & ~(size_t)255);
            memset(comp, 0, (size_t)nfrag * kSlotStride + 512);
            u32* sizes = (u32*)calloc(nfrag, 4);
            const size_t clen = sjo_compress_fragments(raw, sz, 0, nfrag, comp, sizes);
            u64* off = (u64*)calloc(nfrag + 1, 8);
            for (u32 f = 0; f < nfrag; f++) off[f + 1] = off[f] + sizes[f];
            u8* out = (u8*)aligned_alloc(256, ((size_t)nfrag * kBlockSize + 512 + 255) & ~(size_t)255);
            memset(out, 0xEE, (size_t)nfrag * kBlockSize + 512);
            DecodeResult res{};
            IndexedArgs a{comp, off, nfrag, 0, clen, out, (u64)sz, &res};
            for (u32 f = 0; f < nfrag; f++) {  // warp (f % warps-per-CTA) of CTA f / warps-per-CTA
                cpu_warp::W().block = f / kDecodeWarpsPerCta;
                cpu_warp::W().tid_base = (f % kDecodeWarpsPerCta) * 32;
                cpu_warp::W().block_dim = kDecodeWarpsPerCta * 32;
                cpu_warp::run_warp(entry_indexed, &a);
            }
            const bool ok = res.fallback == 0 && memcmp(out, raw, sz) == 0 && out[sz] == 0xEE;
            printf("%s: %u fragments, indexed decoder %s\n", argv[ai], nfrag, ok ? "0 mismatches" : "MISMATCH");
            failed += !ok;
            free(comp); free(sizes); free(off); free(out);
        } else {
            size_t claimed = 0;
            u32 v = 0;
            size_t hdr = 0;
            const int hrc = sjo_parse32(raw, sz, 0, &v, &hdr);
            if (hrc != SJO_OK) {
                printf("%s: header rejected by the oracle (%d), nothing to decode\n", argv[ai], hrc);
                continue;
            }
            claimed = v;
            u8* want = (u8*)malloc(claimed + 64);
            size_t wlen = claimed, werr = 0;
            const int wrc = sjo_uncompress_ex(raw, sz, want, &wlen, &werr);
            u8* out = (u8*)aligned_alloc(256, (claimed + 512 + 255) & ~(size_t)255);
            memset(out, 0xEE, claimed + 512);
            DecodeResult res{};
            ExactArgs a{raw, (u64)sz, (u64)hdr, out, (u64)claimed, &res};
            cpu_warp::W().block = 0;
            cpu_warp::W().tid_base = 0;
            cpu_warp::W().block_dim = 32;
            cpu_warp::run_warp(entry_exact, &a);
            bool ok = res.status == wrc;
            if (ok && wrc == SJO_OK) ok = res.produced == claimed && memcmp(out, want, claimed) == 0;
            if (ok && wrc != SJO_OK) ok = res.err_op == werr;
            printf("%s: exact decoder status %d (oracle %d), produced %llu, %s\n", argv[ai], res.status, wrc,
                   (unsigned long long)res.produced, ok ? "0 mismatches" : "MISMATCH");
            failed += !ok;
            free(want); free(out);
        }
        free(raw);
    }
    return failed ? 1 : 0;
}
