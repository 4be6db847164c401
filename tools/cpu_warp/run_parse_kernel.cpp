// run_parse_kernel.cpp -- runs the REAL index-free parse (snappy.jl_b200/csrc/parse.cuh: one thread per 1 KiB chunk of
// compressed bytes, passes A-E) on the CPU and checks the side index it builds against a plain sequential walk of
// the stream.  TEST INFRASTRUCTURE ONLY (tests/test_kernel_on_cpu_warp.py).  The kernel launches follow
// build_index_segment in snappy_b200.cu; the size scan between k_parse_final and k_build_index (k_scan_sizes, a
// multi-warp kernel) is a plain loop here.  Every file is a STREAM.  Checked per stream:
//   whole stream as one segment, and cut into two segments at the middle (the streamed host path):
//   flags == 0 and total == claimed  =>  the index equals the true element position at every 64 KiB output boundary
//   (copy offsets are not the parse's business: the indexed decoder validates every fragment it is given);
//   a stream the walk finds fragment-clean must come out with flags == 0.
// CUTS=1: the clean-cut re-tiling (k_cut_low, k_cut_tiles, k_cut_fill; the suffix minimum between them, a multi-warp
//   kernel, is a plain loop here) against a brute-force computation: tile f must start at the greatest element start
//   <= f * 65536 that no later copy reaches across.
#include "../../snappy.jl_b200/csrc/parse.cuh"

#include <string>
#include <vector>

extern "C" {
#include "../../oracle/snappy_oracle.h"
}

namespace sb200 {
u8 smem[1024] __attribute__((aligned(128)));
}
using namespace sb200;
using cpu_warp::launch_independent_threads;

struct Truth {
    bool valid = true, clean = true;  // valid: the tag chain is well formed and produces `claimed` bytes
    bool offsets_ok = true;           // copy offsets are the DECODER's business (it validates every fragment)
    std::vector<u64> index;  // stream offset of the element that starts each 64 KiB output boundary
    u64 produced = 0;
};

// Appendix A of SURVEY.md, sequentially
static Truth walk(const u8* in, u64 L, u64 hdr, u64 claimed) {
    Truth t;
    const u64 nfrag = (claimed + kBlockSize - 1) / kBlockSize;
    t.index.assign(nfrag + 1, ~0ull);
    u64 ip = hdr, op = 0;
    while (ip < L) {
        const u64 at = ip;
        const u32 c = in[ip++];
        u64 len, off = 0;
        bool copy = true;
        if ((c & 3) == 0) {
            copy = false;
            len = (c >> 2) + 1;
            if (len > 60) {
                const u32 nb = (u32)len - 60;
                if (ip + nb > L) { t.valid = false; break; }
                u32 v = 0;
                for (u32 i = 0; i < nb; i++) v |= (u32)in[ip + i] << (8 * i);
                ip += nb;
                len = (u64)v + 1;
            }
            if (ip + len > L) { t.valid = false; break; }
            ip += len;
        } else if ((c & 3) == 1) {
            if (ip + 1 > L) { t.valid = false; break; }
            len = 4 + ((c >> 2) & 7);
            off = ((u64)(c >> 5) << 8) | in[ip];
            ip += 1;
        } else {
            const u32 nb = (c & 3) == 2 ? 2 : 4;
            if (ip + nb > L) { t.valid = false; break; }
            len = (c >> 2) + 1;
            for (u32 i = 0; i < nb; i++) off |= (u64)in[ip + i] << (8 * i);
            ip += nb;
        }
        if (copy && (off == 0 || off > op)) t.offsets_ok = false;
        const u64 in_frag = op & (kBlockSize - 1);
        if (in_frag == 0 && (op >> 16) < nfrag) t.index[op >> 16] = at;
        if (in_frag + len > kBlockSize) t.clean = false;
        if (copy && off > in_frag) t.clean = false;
        op += len;
    }
    t.produced = op;
    if (op != claimed) t.valid = false;
    t.index[nfrag] = L;
    return t;
}

struct SegResult {
    u64 flags, exit, total;
};

// build_index_segment (snappy_b200.cu), kernel for kernel
static SegResult parse_segment(const u8* in, u64 n, u64 hdr, u64 E, u64 out_base, u32 nfrag, u64* index,
                               std::vector<u8>* keep_arena = nullptr, std::vector<u64>* keep_off = nullptr,
                               u64* out_start = nullptr, u64 out_total = 0) {
    const u32 pshift = kParseChunkLog2;
    const u64 pchunk = 1ull << pshift, body = E - hdr;
    const u32 nchunk = (u32)((body + pchunk - 1) / pchunk);
    std::vector<u8> own_arena;
    std::vector<u8>& arena = keep_arena ? *keep_arena : own_arena;
    arena.assign(ParseArrays::bytes(nchunk) + 64, 0);
    ParseArrays pa;
    pa.carve(arena.data(), nchunk);
    memset(pa.counters, 0, 64);
    const u32 pgrid = (nchunk + kParseThreads - 1) / kParseThreads, lgrid = (nchunk + 255) / 256;
    launch_independent_threads(pgrid, kParseThreads, k_parse_guess, in, n, hdr, nchunk, pa, E, pshift);
    launch_independent_threads(pgrid, kParseThreads, k_parse_bridge, in, n, hdr, nchunk, pa, E, pshift);
    std::vector<u32> next_orig(pa.next_a, pa.next_a + nchunk);
    u32 *nx = pa.next_a, *nx2 = pa.next_b;
    for (u32 span = 1; span < nchunk; span <<= 1) {
        launch_independent_threads(lgrid, 256u, k_parse_reach, nchunk, (const u32*)nx, nx2, pa.reach);
        u32* t = nx;
        nx = nx2;
        nx2 = t;
    }
    launch_independent_threads(lgrid, 256u, k_parse_reach, nchunk, (const u32*)nx, nx2, pa.reach);
    launch_independent_threads(pgrid, kParseThreads, k_parse_entries, in, n, hdr, nchunk, pa, (const u32*)next_orig.data(), 0, E, pshift);
    launch_independent_threads(pgrid, kParseThreads, k_parse_entries, in, n, hdr, nchunk, pa, (const u32*)next_orig.data(), 1, E, pshift);
    launch_independent_threads(pgrid, kParseThreads, k_parse_final, in, n, hdr, nchunk, pa, E, pshift);
    std::vector<u64> out_off(nchunk + 1, 0);  // k_scan_sizes: exclusive scan, total behind the end
    for (u32 k = 0; k < nchunk; k++) out_off[k + 1] = out_off[k] + pa.outb[k];
    if (out_start)  // relaxed form (decode_parsed_locked): tiles may start at a straddling literal
        launch_independent_threads(pgrid, kParseThreads, k_build_index, in, n, hdr, nchunk, pa, (const u64*)out_off.data(), index, nfrag,
                                   E, out_base, pshift, out_start, 1u, out_total);
    else
        launch_independent_threads(pgrid, kParseThreads, k_build_index, in, n, hdr, nchunk, pa, (const u64*)out_off.data(), index, nfrag,
                                   E, out_base, pshift, (u64*)nullptr, 0u, (u64)0);  // strict form: aligned tiles only
    u64 host3[4];
    k_parse_report(pa.counters, out_off.data() + nchunk, host3);
    if (keep_off) *keep_off = out_off;
    return SegResult{host3[0], host3[1], host3[2]};
}

// the re-tiling of decode_parsed_locked (snappy_b200.cu) against brute force; returns the number of wrong tiles
static long check_cuts(const u8* in, u64 n, u64 hdr, u32 claimed, u32 nfrag, std::string& note) {
    std::vector<u8> arena;
    std::vector<u64> out_off, index(nfrag + 1, ~0ull), ost(nfrag + 1, ~0ull);
    index[0] = hdr;
    ost[0] = 0;
    const SegResult r = parse_segment(in, n, hdr, n, 0, nfrag, index.data(), &arena, &out_off, ost.data(), claimed);
    if ((r.flags & (PF_ANOMALY | PF_BROKEN)) || r.total != claimed) {
        note += " [cuts: parse incomplete, skipped]";
        return 0;
    }
    ost[nfrag] = claimed;
    const u32 pshift = kParseChunkLog2;
    const u32 nchunk = (u32)(((n - hdr) + (1ull << pshift) - 1) >> pshift);
    ParseArrays pa;
    pa.carve(arena.data(), nchunk);
    std::vector<u64> low(nchunk, 0), sfx(nchunk, ~0ull);
    const u32 pgrid = (nchunk + kParseThreads - 1) / kParseThreads;
    launch_independent_threads(pgrid, kParseThreads, k_cut_low, in, n, hdr, nchunk, pa, (const u64*)out_off.data(), n, pshift, low.data());
    u64 run = ~0ull;  // k_suffix_min
    for (u32 k = nchunk; k-- > 0;) {
        sfx[k] = run;
        run = low[k] < run ? low[k] : run;
    }
    if (nfrag > 1)
        launch_independent_threads((nfrag - 1 + kParseThreads - 1) / kParseThreads, kParseThreads, k_cut_tiles, in, n, hdr, nchunk, pa,
                                   (const u64*)out_off.data(), (const u64*)sfx.data(), n, pshift, nfrag, index.data(), ost.data());
    k_cut_fill(nfrag, index.data(), ost.data());
    // brute force: every element's (output position, stream position, reach), suffix minimum of the reaches
    struct El { u64 op, ip, reach; };
    std::vector<El> els;
    {
        u64 ip = hdr, op = 0;
        Element e;
        while (ip + 1 < n) {
            const u64 at = ip;
            if (!walk_step(in, n, ip, e)) break;
            els.push_back(El{op, at, e.is_copy ? ((u64)e.offset <= op ? op - e.offset : 0) : ~0ull});
            op += e.len;
        }
    }
    std::vector<u64> R(els.size() + 1, ~0ull);
    for (size_t i = els.size(); i-- > 0;) R[i] = els[i].reach < R[i + 1] ? els[i].reach : R[i + 1];
    long bad = 0;
    size_t j = 0, best = 0;  // best: greatest clean element index with op <= boundary (element 0 is always clean)
    for (u32 f = 1; f < nfrag; f++) {
        const u64 B = (u64)f << 16;
        while (j < els.size() && els[j].op <= B) {
            if (R[j] >= els[j].op) best = j;
            j++;
        }
        if (ost[f] != els[best].op || index[f] != els[best].ip) {
            if (bad++ < 3)
                fprintf(stderr, "tile %u: cut at output %llu (stream %llu), brute force says %llu (%llu)\n", f, (unsigned long long)ost[f],
                        (unsigned long long)index[f], (unsigned long long)els[best].op, (unsigned long long)els[best].ip);
        }
    }
    return bad;
}

int main(int argc, char** argv) {
    int failed = 0;
    const bool cuts = getenv("CUTS") && atoi(getenv("CUTS")) != 0;
    for (int ai = 1; ai < argc; ai++) {
        FILE* fp = fopen(argv[ai], "rb");
        if (!fp) { perror(argv[ai]); return 2; }
        fseek(fp, 0, SEEK_END);
        const size_t sz = (size_t)ftell(fp);
        fseek(fp, 0, SEEK_SET);
        std::vector<u8> buf(sz + 64, 0);  // the device buffers carry 16 zero bytes of slack behind the stream
        if (fread(buf.data(), 1, sz, fp) != sz) return 2;
        fclose(fp);
        u32 claimed = 0;
        size_t hdr = 0;
        if (sjo_parse32(buf.data(), sz, 0, &claimed, &hdr) != SJO_OK || sz <= hdr || claimed == 0) {
            printf("%s: no body to parse\n", argv[ai]);
            continue;
        }
        const u32 nfrag = (claimed + kBlockSize - 1) / kBlockSize;
        const Truth t = walk(buf.data(), sz, hdr, claimed);
        bool ok = true;
        std::string note;
        // ---- one segment
        std::vector<u64> index(nfrag + 1, ~0ull);
        const SegResult r = parse_segment(buf.data(), sz, hdr, sz, 0, nfrag, index.data());
        const bool accepted = r.flags == 0 && r.total == claimed;
        if (accepted && (!t.valid || !t.clean || index != t.index)) { ok = false; note += " [accepted a wrong index]"; }
        if (t.valid && t.clean && !accepted) { ok = false; note += " [rejected a clean stream]"; }
        // ---- two segments (the second starts at the first one's reported exit)
        bool accepted2 = false;
        const bool two = sz - hdr > 4096;
        if (two) {
            std::vector<u64> index2(nfrag + 1, ~0ull);
            const u64 E1 = hdr + (sz - hdr) / 2;
            const SegResult a = parse_segment(buf.data(), sz, hdr, E1, 0, nfrag, index2.data());
            if (a.flags == 0 && a.exit >= E1 && a.exit <= sz) {
                SegResult b{0, sz, 0};
                if (a.exit < sz) b = parse_segment(buf.data(), sz, a.exit, sz, a.total, nfrag, index2.data());
                else index2[nfrag] = sz;
                accepted2 = b.flags == 0 && a.total + b.total == claimed;
                if (accepted2 && (!t.valid || !t.clean || index2 != t.index)) { ok = false; note += " [two segments: wrong index]"; }
            }
            if (t.valid && t.clean && !accepted2) { ok = false; note += " [two segments: rejected a clean stream]"; }
        }
        if (cuts && t.valid) {
            const long badc = check_cuts(buf.data(), sz, hdr, claimed, nfrag, note);
            if (badc) { ok = false; note += " [clean cuts: " + std::to_string(badc) + " wrong tiles]"; }
            else note += " [clean cuts ok]";
        }
        printf("%s: %u fragments, stream %s%s/%s, parse %s (flags %llu, total %llu of %u), two segments %s: %s%s\n", argv[ai], nfrag,
               t.valid ? "valid" : "INVALID", t.offsets_ok ? "" : " (bad copy offset)", t.clean ? "clean" : "not clean", accepted ? "accepted" : "declined",
               (unsigned long long)r.flags, (unsigned long long)r.total, claimed, !two ? "n/a" : (accepted2 ? "accepted" : "declined"),
               ok ? "0 mismatches" : "MISMATCH", note.c_str());
        failed += !ok;
    }
    return failed ? 1 : 0;
}
