// run_pages_kernel.cpp -- runs the REAL batched-page kernels (k_compress_pages in snappy.jl_b200/csrc/compress.cuh,
// k_decode_pages in decompress.cuh; one warp per independent stream) on the CPU (cuda_shim.h).  TEST INFRASTRUCTURE
// ONLY (tests/test_kernel_on_cpu_warp.py).  Every file is cut into pages (PAGE bytes, default 4096; every 7th page
// ragged; the file's first 150 000 bytes also go in as ONE multi-fragment page), each page is compressed by the
// kernel and compared with the oracle's stream for that page (RULES=0|1|2), then all pages are decoded by
// k_decode_pages and compared with the input.  The fragment staging (TMA bulk copy + mbarrier on the GPU) is a plain
// copy under SB200_CPU_EMU.
#include "../../snappy.jl_b200/csrc/compress.cuh"
#include "../../snappy.jl_b200/csrc/compress_window.cuh"
#include "../../snappy.jl_b200/csrc/decompress.cuh"

#include <vector>

extern "C" {
#include "../../oracle/snappy_oracle.h"
}

namespace sb200 {
u8 smem[256 * 1024] __attribute__((aligned(128)));
}
using namespace sb200;

struct CArgs {
    const u8* in;
    const u64* in_off;
    const u32* in_size;
    u8* out;
    const u64* out_off;
    u32* out_size;
    u32 frag_cap, table_cap, rules;
};
struct WArgs {  // k_compress_pages_window (KERNEL=window): persistent warp, pages pulled from *counter
    const u8* in;
    const u64* in_off;
    const u32* in_size;
    u32 count;
    u8* out;
    const u64* out_off;
    u32* out_size;
    u32 ring, table_cap;
    u32* counter;
    u32 rules;
};
static void entry_window(void* p) {
    const WArgs& a = *(const WArgs*)p;
    if (a.rules) k_compress_pages_window<true>(a.in, a.in_off, a.in_size, a.count, a.out, a.out_off, a.out_size, a.ring, a.table_cap, a.counter, a.rules);
    else k_compress_pages_window<false>(a.in, a.in_off, a.in_size, a.count, a.out, a.out_off, a.out_size, a.ring, a.table_cap, a.counter, 0u);
}
static void entry_compress(void* p) {
    const CArgs& a = *(const CArgs*)p;
    if (a.rules) k_compress_pages<true>(a.in, a.in_off, a.in_size, a.out, a.out_off, a.out_size, a.frag_cap, a.table_cap, a.rules);
    else k_compress_pages<false>(a.in, a.in_off, a.in_size, a.out, a.out_off, a.out_size, a.frag_cap, a.table_cap, 0u);
}
struct DArgs {
    const u8* in;
    const u64* in_off;
    const u32* in_size;
    u32 count;
    u8* out;
    const u64* out_off;
    const u32* out_cap;
    u32* out_size;
    int* statuses;
};
static void entry_decode(void* p) {
    const DArgs& a = *(const DArgs*)p;
    k_decode_pages(a.in, a.in_off, a.in_size, a.count, a.out, a.out_off, a.out_cap, a.out_size, a.statuses);
}

static void entry_init(void*) { k_init_probe_offsets(); }

int main(int argc, char** argv) {
    cpu_warp::W().block = cpu_warp::W().tid_base = 0;
    cpu_warp::W().block_dim = 32;
    cpu_warp::run_warp(entry_init, nullptr);  // the skip-heuristic probe offsets (a __device__ table on the GPU)
    const u32 rules = getenv("RULES") ? (u32)atoi(getenv("RULES")) : 0u;
    const u32 page = getenv("PAGE") ? (u32)atoi(getenv("PAGE")) : 4096u;
    const bool window = getenv("KERNEL") && !strcmp(getenv("KERNEL"), "window");  // pages <= 8 KiB only
    int failed = 0;
    for (int ai = 1; ai < argc; ai++) {
        FILE* fp = fopen(argv[ai], "rb");
        if (!fp) { perror(argv[ai]); return 2; }
        fseek(fp, 0, SEEK_END);
        const size_t sz = (size_t)ftell(fp);
        fseek(fp, 0, SEEK_SET);
        std::vector<u8> raw(sz + 256, 0);
        if (fread(raw.data(), 1, sz, fp) != sz) return 2;
        fclose(fp);
        // the pages: offsets 16-byte aligned inside one flat buffer, like the Python wrapper lays them out
        std::vector<u64> in_off, out_off;
        std::vector<u32> in_size;
        std::vector<u8> flat;
        auto add = [&](const u8* p, size_t n) {
            in_off.push_back(flat.size());
            in_size.push_back((u32)n);
            flat.insert(flat.end(), p, p + n);
            flat.resize((flat.size() + 15) & ~(size_t)15, 0);
        };
        u32 k = 0;
        for (size_t o = 0; o < sz; o += page, k++) {
            size_t n = sz - o < page ? sz - o : page;
            if (k % 7 == 3) n = (n * 37 / 100);  // ragged (and an empty page now and then)
            if (k % 29 == 11) n = 0;
            add(raw.data() + o, n);
        }
        if (!window) add(raw.data(), sz < 150000 ? sz : 150000);  // one multi-fragment page
        flat.resize(flat.size() + 256, 0);
        const u32 count = (u32)in_size.size();
        u32 max_size = 0;
        u64 cap_total = 0;
        for (u32 i = 0; i < count; i++) {
            max_size = in_size[i] > max_size ? in_size[i] : max_size;
            out_off.push_back(cap_total);
            cap_total += (sjo_maxlength_compressed(in_size[i]) + 15) & ~(size_t)15;
        }
        std::vector<u8> out(cap_total + 256, 0xEE);
        std::vector<u32> out_size(count, 0);
        u32 frag_cap = max_size < kBlockSize ? ((max_size + 15) & ~15u) : kBlockSize;
        if (frag_cap < 16) frag_cap = 16;
        u32 entries = 256;
        while (entries < (rules == 2 ? 2u : 1u) * kMaxTableEntries && entries < max_size) entries <<= 1;
        CArgs ca{flat.data(), in_off.data(), in_size.data(), out.data(), out_off.data(), out_size.data(), frag_cap, entries, rules};
        long bad = 0;
        std::vector<u8> want(sjo_maxlength_compressed(max_size) + 64);
        if (window) {  // one persistent warp compresses every page (the CTA's warps never talk to each other)
            u32 ring = 1024;
            while (ring < max_size) ring <<= 1;
            if (getenv("RINGX")) ring *= (u32)atoi(getenv("RINGX"));
            u32 counter = 0;
            WArgs wa{flat.data(), in_off.data(), in_size.data(), count, out.data(), out_off.data(), out_size.data(),
                     ring, entries, &counter, rules};
            cpu_warp::W().block = 0;
            cpu_warp::W().tid_base = 32;   // warp 1 of a 2-warp CTA: table and ring are not at the start of smem
            cpu_warp::W().block_dim = 64;
            cpu_warp::run_warp(entry_window, &wa);
        }
        for (u32 i = 0; i < count; i++) {
            if (!window) {
                cpu_warp::W().block = i;
                cpu_warp::W().tid_base = 0;
                cpu_warp::W().block_dim = 32;
                cpu_warp::run_warp(entry_compress, &ca);
            }
            size_t wl = want.size();
            if (sjo_compress_rules(flat.data() + in_off[i], in_size[i], want.data(), &wl, (int)rules) != SJO_OK) return 2;
            if (wl != out_size[i] || memcmp(want.data(), out.data() + out_off[i], wl)) {
                if (bad++ < 3) {
                    fprintf(stderr, "%s: page %u (%u bytes) differs (%u vs %zu)\n", argv[ai], i, in_size[i], out_size[i], wl);
                    if (getenv("DUMP")) {
                        for (u32 k = 0; k < 48; k++) fprintf(stderr, "%02x ", out[out_off[i] + k]);
                        fprintf(stderr, "\n");
                        for (u32 k = 0; k < 48; k++) fprintf(stderr, "%02x ", want[k]);
                        fprintf(stderr, "\n");
                    }
                }
            }
        }
        // decode every page from the kernel's own output
        std::vector<u64> back_off;
        u64 back_total = 0;
        for (u32 i = 0; i < count; i++) {
            back_off.push_back(back_total);
            back_total += (in_size[i] + 15) & ~15u;
        }
        std::vector<u8> back(back_total + 256, 0xEE);
        std::vector<u32> back_size(count, 0);
        std::vector<int> statuses(count, -1);
        DArgs da{out.data(), out_off.data(), out_size.data(), count, back.data(), back_off.data(), in_size.data(),
                 back_size.data(), statuses.data()};
        long dbad = 0;
        for (u32 i = 0; i < count; i++) {
            cpu_warp::W().block = i / kDecodeWarpsPerCta;
            cpu_warp::W().tid_base = (i % kDecodeWarpsPerCta) * 32;
            cpu_warp::W().block_dim = kDecodeWarpsPerCta * 32;
            cpu_warp::run_warp(entry_decode, &da);
            if (statuses[i] != 0 || back_size[i] != in_size[i] || memcmp(back.data() + back_off[i], flat.data() + in_off[i], in_size[i]))
                if (dbad++ < 3) fprintf(stderr, "%s: page %u decodes wrong (status %d)\n", argv[ai], i, statuses[i]);
        }
        printf("%s: %u pages (rules %u), compress %ld, decode %ld: %s\n", argv[ai], count, rules, bad, dbad,
               (bad || dbad) ? "MISMATCH" : "0 mismatches");
        failed += (bad || dbad) != 0;
    }
    return failed ? 1 : 0;
}
