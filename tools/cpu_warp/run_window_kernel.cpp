// run_window_kernel.cpp -- runs the REAL k_compress_window kernel source (snappy.jl_b200/csrc/compress_window.cuh,
// compress_chain.cuh) as one warp on the CPU (cuda_shim.h) over the fragments of the given files and compares every
// fragment's bytes with the oracle.  TEST INFRASTRUCTURE ONLY: where no GPU exists this is the check that the kernel
// source itself -- not a model of it -- makes the reference's decisions (tests/test_kernel_on_cpu_warp.py).
//
//   g++ -O1 -std=c++17 -DSB200_CPU_EMU -Itools/cpu_warp tools/cpu_warp/run_window_kernel.cpp oracle/snappy_oracle.c
//   TABLE=smem|global RULES=0|1|2 RING=2048 KERNEL=window|chain SLOWCONT=0|1 ./a.out file...
#include "../../snappy.jl_b200/csrc/compress_pipe.cuh"

extern "C" {
#include "../../oracle/snappy_oracle.h"
}

namespace sb200 {
u8 smem[256 * 1024] __attribute__((aligned(128)));
}
using namespace sb200;

struct Args {
    bool smem_table;
    const u8* in;
    u64 len;
    u32 nfrag, shift;
    const u8* tail;
    u8* scratch;
    u32* sizes;
    u32* counter;
    u16* gtables;
    u32 ring, rules;
    bool chain;  // KERNEL=chain: the step-wise kernel (option window=0), rules 0 only
    bool slowcont;  // SLOWCONT=1: the experimental window variant (option slowcont), rules 0 only
    bool mixed;     // MIXED=1: k_compress_window_mixed (one CTA holds both table placements)
    bool two;       // TWO=1 (with MIXED=1, rules 0): the two-window round
    bool pipe;      // PIPE=1 (with MIXED=1): the pipelined round (compress_pipe.cuh)
};

static WindowArgs wargs(const Args& a) {
    WindowArgs w;
    memset(&w, 0, sizeof w);
    w.g_in = a.in;
    w.shard_len = a.len;
    w.nfrag = a.nfrag;
    w.shift = a.shift;
    w.tail_copy = a.tail;
    w.scratch = a.scratch;
    w.frag_sizes = a.sizes;
    w.counter = a.counter;
    w.gtables = a.gtables;
    w.done_div = 1;
    w.lib_rules = a.rules;
    return w;
}
template <bool kSmem, bool kLib>
static void launch(const Args& a) {
    if (a.mixed) {  // the default launch form: both placements in one CTA; this warp plays the role TABLE names
        // CTA of wb = 1 global-table warp and wa = 1 shared-table warp: warp 0 / warp 1 (the harness sets tid_base)
        if (getenv("UNI") && atoi(getenv("UNI")) && !kLib) {
            if (a.pipe) k_compress_window_mixed<false, false, false, 3, true>(wargs(a), 1u, 1u, a.ring, a.ring, 0u);
            else k_compress_window_mixed<false, false, false, 0, true>(wargs(a), 1u, 1u, a.ring, a.ring, 0u);
        } else if (a.pipe) k_compress_window_mixed<kLib, false, false, 3>(wargs(a), 1u, 1u, a.ring, a.ring, 0u);
        else if (a.two && !kLib) k_compress_window_mixed<false, true, true>(wargs(a), 1u, 1u, a.ring, a.ring, 0u);
        else k_compress_window_mixed<kLib>(wargs(a), 1u, 1u, a.ring, a.ring, 0u);
        return;
    }
    k_compress_window<kSmem, kLib>(wargs(a), a.ring);
}
static void entry(void* p) {
    const Args& a = *(const Args*)p;
    if (a.chain) {
        if (a.smem_table) k_compress_chain<true>(a.in, a.len, a.nfrag, a.shift, a.tail, a.scratch, a.sizes, a.counter, a.gtables, 32u, 0u, nullptr, 0u);
        else k_compress_chain<false>(a.in, a.len, a.nfrag, a.shift, a.tail, a.scratch, a.sizes, a.counter, a.gtables, 16u, 0u, nullptr, 0u);
        return;
    }
    if (a.slowcont) {
        if (a.smem_table)
            k_compress_window<true, false, true>(wargs(a), a.ring);
        else
            k_compress_window<false, false, true>(wargs(a), a.ring);
        return;
    }
    if (a.smem_table) {
        if (a.rules) launch<true, true>(a);
        else launch<true, false>(a);
    } else {
        if (a.rules) launch<false, true>(a);
        else launch<false, false>(a);
    }
}

int main(int argc, char** argv) {
    const char* tab = getenv("TABLE");
    const bool smem_table = !(tab && !strcmp(tab, "global"));
    const u32 rules = getenv("RULES") ? (u32)atoi(getenv("RULES")) : 0u;
    const u32 ring = getenv("RING") ? (u32)atoi(getenv("RING")) : 2048u;
    const bool chain = getenv("KERNEL") && !strcmp(getenv("KERNEL"), "chain");
    const bool slowcont = getenv("SLOWCONT") && atoi(getenv("SLOWCONT")) != 0;
    const bool mixed = getenv("MIXED") && atoi(getenv("MIXED")) != 0;
    const bool two = getenv("TWO") && atoi(getenv("TWO")) != 0;
    const bool pipe = getenv("PIPE") && atoi(getenv("PIPE")) != 0;
    if ((chain || slowcont) && rules) {
        fprintf(stderr, "KERNEL=chain has no rules instantiation\n");
        return 2;
    }
    k_init_probe_offsets();
    int failed = 0;
    for (int ai = 1; ai < argc; ai++) {
        FILE* fp = fopen(argv[ai], "rb");
        if (!fp) {
            perror(argv[ai]);
            return 2;
        }
        fseek(fp, 0, SEEK_END);
        const long sz = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        // 16-byte aligned like a device allocation; 0xA5 behind the data: nothing read there may matter
        u8* in = (u8*)aligned_alloc(256, ((size_t)sz + 1024 + 255) & ~(size_t)255);
        memset(in, 0xA5, (size_t)sz + 1024);
        if (fread(in, 1, (size_t)sz, fp) != (size_t)sz) return 2;
        fclose(fp);
        const u32 nfrag = (u32)((sz + kBlockSize - 1) / kBlockSize);
        if (nfrag == 0) continue;
        // padded copy of the last fragment (stage_tail in snappy_b200.cu)
        const u64 tail_start = (u64)(nfrag - 1) * kBlockSize;
        const size_t tail_len = (size_t)sz - tail_start;
        u8* tail = (u8*)aligned_alloc(256, kBlockSize + kTailPad + 512);
        memset(tail, 0xA5, kBlockSize + kTailPad + 512);
        memcpy(tail, in + tail_start, tail_len);
        memset(tail + tail_len, 0, kTailPad);
        u8* scratch = (u8*)malloc((size_t)nfrag * kSlotStride);
        u32* sizes = (u32*)calloc(nfrag, 4);
        u16* gtables = (u16*)aligned_alloc(256, 2 * kMaxTableEntries * 2);
        u32 counter = 0;
        u32 entries = sjo_hashtable_entries((u64)sz), shift = 32;
        for (u32 e = entries; e > 1; e >>= 1) shift--;
        Args a{smem_table, in, (u64)sz, nfrag, shift, tail, scratch, sizes, &counter, gtables, ring, rules, chain, slowcont, mixed, two, pipe};
        cpu_warp::W().collectives = 0;
        if (mixed) {  // CTA of two warps: warp 0 = global-table role, warp 1 = shared-table role
            cpu_warp::W().block = 0;
            cpu_warp::W().tid_base = smem_table ? 32 : 0;
            cpu_warp::W().block_dim = 64;
        }
        cpu_warp::run_warp(entry, &a);
        // the oracle, fragment by fragment
        u8* want = (u8*)malloc(kSlotStride);
        u16* table = (u16*)malloc(2 * kMaxTableEntries * 2);
        long bad = 0;
        for (u32 f = 0; f < nfrag; f++) {
            const size_t n = (f == nfrag - 1) ? tail_len : kBlockSize;
            size_t c;
            if (rules) {
                u32 e = 256;
                while (e < (rules == 2 ? 2 * kMaxTableEntries : kMaxTableEntries) && e < n) e <<= 1;
                memset(table, 0, e * 2);
                c = sjo_compress_fragment_rules(in + (size_t)f * kBlockSize, n, want, table, e, (int)rules);
            } else {
                memset(table, 0xff, entries * 2);
                c = sjo_compress_fragment(in + (size_t)f * kBlockSize, n, want, table, entries);
            }
            if (c != sizes[f] || memcmp(want, scratch + (size_t)f * kSlotStride, c)) {
                if (bad++ < 3) fprintf(stderr, "%s: fragment %u differs (%u vs %zu bytes)\n", argv[ai], f, sizes[f], c);
            }
        }
        if (getenv("ROUNDS")) {
            printf("rounds %lu, second windows entered %lu\n", sb200::g_emu_rounds, sb200::g_emu_second);
            if (pipe)
                printf("pipelined rounds %lu: %lu cold, %lu with re-evaluated lanes, mean window %.1f\n", sb200::g_emu_pipe_rounds,
                       sb200::g_emu_pipe_cold, sb200::g_emu_pipe_fix, (double)sb200::g_emu_pipe_w / (double)(sb200::g_emu_pipe_rounds ? sb200::g_emu_pipe_rounds : 1));
        }
#ifdef SB200_CPU_EMU_STATS
        {
            unsigned long all = 0, hits = 0, cum = 0;
            for (int b = 0; b < 17; b++) all += sb200::g_emu_dist[b], hits += sb200::g_emu_dist_hit[b];
            printf("valid lanes %lu, hits %lu (%.1f %%); cumulative share of candidates closer than 2^k:", all, hits, 100.0 * hits / (all ? all : 1));
            for (int b = 0; b < 17; b++) {
                cum += sb200::g_emu_dist[b];
                if (b >= 8) printf(" %d:%.0f%%", b + 1, 100.0 * cum / (all ? all : 1));
                sb200::g_emu_dist[b] = sb200::g_emu_dist_hit[b] = 0;
            }
            printf("\n");
        }
#endif
        sb200::g_emu_rounds = sb200::g_emu_second = 0;
        sb200::g_emu_pipe_rounds = sb200::g_emu_pipe_cold = sb200::g_emu_pipe_fix = sb200::g_emu_pipe_w = 0;
        printf("%s: %u fragments, %ld mismatches (%s kernel, %s table, rules %u, ring %u, %llu collectives)\n", argv[ai], nfrag, bad,
               chain ? "chain" : (slowcont ? "window+slowcont" : "window"), smem_table ? "shared" : "global", rules, ring, (unsigned long long)cpu_warp::W().collectives);
        failed += bad != 0;
        free(in); free(tail); free(scratch); free(sizes); free(gtables); free(want); free(table);
    }
    return failed ? 1 : 0;
}
