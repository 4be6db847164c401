/* emulate_window.c -- CPU model of the window-parallel fragment compressor (csrc/compress_window.cuh).
 *
 * Development aid, not part of the product: it evaluates the *round* algorithm the CUDA kernel runs
 * (32 positions per round: every lane looks its position up in the table as of the round start,
 * the real chain is then followed through the window, and a lane whose hash equals an earlier
 * lane's is never trusted) with plain loops, so that its exactness against the oracle
 * (oracle/snappy_oracle.c, the restatement of src/internal.jl:127-250) can be checked on a CPU
 * before the kernel ever runs.   gcc -O2 -o emulate_window tools/emulate_window.c oracle/snappy_oracle.c
 *
 * Usage: emulate_window FILE...   -> per file: fragments compared, mismatches, rounds per fragment
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../oracle/snappy_oracle.h"

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* RULES (env): 0 = Snappy.jl, 1 = libsnappy <= 1.1.7, 2 = Google snappy >= 1.1.9 (the kernels' kLib instantiations,
 * DESIGN.md 4c): ip_limit margin, the 60-byte literal, the bucket function and the table size per fragment */
static int RULES = 0;
typedef struct { const uint8_t *F; long n, lim; uint32_t shift, hmask; uint16_t T[32768]; uint8_t *out, *op;
                 long rounds, hops, slow, generic; } Frag;

static uint32_t hashw(const Frag *f, uint32_t w) {
    return RULES == 2 ? (((w * 0x1e35a7bdu) >> 17) & f->hmask) : ((w * 0x1e35a7bdu) >> f->shift);
}

static void emit_literal(Frag *f, long from, long to) {
    long len = to - from; if (len <= 0) return;
    uint32_t n = (uint32_t)(len - 1); uint8_t *op = f->op;
    if (len < (RULES ? 61 : 60)) *op++ = (uint8_t)(n << 2);
    else { uint8_t *base = op; int count = 0; while (n > 0) { *++op = (uint8_t)n; n >>= 8; count++; } *base = (uint8_t)((59 + count) << 2); op++; }
    memcpy(op, f->F + from, (size_t)len); f->op = op + len;
}
static uint8_t *copy64(uint8_t *op, uint32_t off, uint32_t len) {
    if (len < 12 && off < 2048) { *op++ = (uint8_t)(1 + ((len - 4) << 2) + ((off >> 3) & 0xe0)); *op++ = (uint8_t)off; }
    else { uint32_t u = 2 + ((len - 1) << 2) + (off << 8); *op++ = (uint8_t)u; *op++ = (uint8_t)(u >> 8); *op++ = (uint8_t)(u >> 16); }
    return op;
}
static void emit_copy(Frag *f, uint32_t off, uint32_t len) {
    uint8_t *op = f->op;
    if (len >= 12) { while (len >= 68) { op = copy64(op, off, 64); len -= 64; } if (len > 64) { op = copy64(op, off, 60); len -= 60; } }
    f->op = copy64(op, off, len);
}
static void record(Frag *f, long lit_from, long ip, long cand, long M) { emit_literal(f, lit_from, ip); emit_copy(f, (uint32_t)(ip - cand), (uint32_t)M); }

/* probe offsets of the skip heuristic, :162-172 */
static uint32_t PO[400];
static void init_po(void) { uint32_t skip = 32, off = 0; PO[0] = 0; for (int i = 1; i < 400; i++) { uint32_t b = skip >> 5; skip += b; off += b; if (off > 0x100000u) off = 0x100000u; PO[i] = off; } }

enum { ARR = 0, SCAN = 1 };

/* bytes compared per lane (the kernels use 16; CAP=32 models "16 more for lanes that match all 16", which saved
 * rounds but cost more instructions than it saved on the GPU: 28.2 vs 27.0 ms); CAP equal bytes go to the whole-warp extension */
static uint32_t CAP = 16;
/* STATS=1: why rounds end and how far each kind advances the window start (guides what to optimise) */
enum { E_LEAVE, E_SCAN_DUP, E_SCAN_LIMIT, E_ARR_DUP, E_ARR_BEYOND, E_SLOW, E_FIN, E_KINDS };
static const char *const E_NAME[E_KINDS] = {"scan covers window", "scan stops at untrusted lane", "scan reaches probe 32",
                                            "arrival at untrusted lane", "copy lands beyond window", "copy >= 16 bytes", "fragment end"};
static long e_count[E_KINDS], e_adv[E_KINDS];
/* variants of the round that are NOT in the kernels (yet); all exact, measured here in rounds per fragment:
 *   SLOWCONT=1  a copy of >= 16 bytes is extended inside the hop loop and the chain goes on in the same window when
 *               it lands there (W=32: dictionary 1951 -> 1835, html 976 -> 900; W=64: 1272 -> 1084, 834 -> 654)
 *   PRECISE=1   a lane with an equal-hash lower lane is distrusted only if that lane was inserted on the path
 *               (W=32: -0.3 .. -2 %; W=64: text 1337 -> 1255) */
static int SLOWCONT = 0, PRECISE = 0;
/* W: positions evaluated per round (compile with -DW=64 to model two positions per lane; the kernels use 32).
 * Measured with this model (rounds per fragment, W = 32 -> 48 -> 64): dictionary class 1951 -> 1502 -> 1272, text
 * class 2014 -> 1543 -> 1337, alice29 1532 -> 1110 -> 911, html 976 -> 864 -> 834, records class 604 -> 603 (long copies). */
#ifndef W
#define W 32
#endif
static size_t compress_fragment_window(Frag *f) {
    const uint8_t *F = f->F; const long n = f->n, lim = n - (RULES ? 15 : 16); f->lim = lim; f->op = f->out;
    memset(f->T, 0, sizeof f->T);
    long lit_from = 0;
    if (n >= 15) {
        int mode = SCAN; long a = 1, scan_s = 1;
        for (;;) {
            /* ---- generic scan rounds (probe index >= 32, stride > 1): the old path */
            if (mode == SCAN && a - scan_s >= 32) {
                f->generic++;
                long k = a - scan_s; /* == 32 */
                long ip = -1, cand = 0; int fin = 0;
                for (;; k++) {
                    long p = scan_s + PO[k], pn = scan_s + PO[k + 1];
                    if (pn > lim) { fin = 1; break; }
                    uint32_t h = hashw(f, ld32(F + p)); cand = f->T[h]; f->T[h] = (uint16_t)p;
                    if (ld32(F + cand) == ld32(F + p)) { ip = p; break; }
                }
                if (fin) break;
                long M = 4; while (ip + M < n && F[cand + M] == F[ip + M]) M++;
                record(f, lit_from, ip, cand, M); ip += M; lit_from = ip;
                if (ip >= lim) break;
                mode = ARR; a = ip; continue;
            }
            f->rounds++;
            /* ---- lane evaluation against the table as of the round start */
            uint32_t H[W], t[W], m[W]; int V[W], dup[W];
            if (mode == ARR) f->T[hashw(f, ld32(F + a - 1))] = (uint16_t)(a - 1); /* :233 */
            for (int l = 0; l < W; l++) {
                long q = a + l; V[l] = q < lim;
                H[l] = V[l] ? hashw(f, ld32(F + q)) : (0x80000000u | (uint32_t)l);
                t[l] = V[l] ? f->T[H[l]] : 0; dup[l] = 0;
                for (int j = 0; j < l; j++) if (H[j] == H[l]) dup[l] = 1;
                uint32_t k = 0; if (V[l]) while (k < CAP && q + k < n && F[t[l] + k] == F[q + k]) k++;
                m[l] = k;
            }
            uint64_t ins = 0; int fin = 0, next_mode = -1; long next_a = 0;
            long slow_ip = -1, slow_cand = 0;
            int l = 0; int scanning = (mode == SCAN); int ek = E_FIN;
            for (;;) {
                f->hops++;
                if (!scanning) { /* arrival at lane l: :228-238 */
                    if (l > 0 && dup[l]) {
                        int stale = 1;
                        if (PRECISE) { stale = 0; for (int j = 0; j < l; j++) if (H[j] == H[l] && ((ins >> j & 1) || j == l - 1)) stale = 1; }
                        if (stale) { next_mode = ARR; next_a = a + l; ek = E_ARR_DUP; break; }
                    }
                    if (l > 0) ins |= 1ull << (l - 1);
                    ins |= 1ull << l;
                    if (m[l] >= 4) {
                        if (m[l] == CAP) {
                            if (!SLOWCONT) { slow_ip = a + l; slow_cand = t[l]; break; }
                            long M = CAP; while (a + l + M < n && F[t[l] + M] == F[a + l + M]) M++;
                            f->slow++;
                            record(f, lit_from, a + l, t[l], M); lit_from = a + l + M;
                            long tgt = l + M;
                            if (a + tgt >= lim) { fin = 1; break; }
                            if (tgt >= W) { next_mode = ARR; next_a = a + tgt; ek = E_SLOW; break; }
                            l = (int)tgt; continue;
                        }
                        record(f, lit_from, a + l, t[l], m[l]); lit_from = a + l + m[l];
                        long tgt = l + m[l];
                        if (a + tgt >= lim) { fin = 1; break; }
                        if (tgt >= W) { next_mode = ARR; next_a = a + tgt; ek = E_ARR_BEYOND; break; }
                        l = (int)tgt; continue;
                    }
                    scanning = 1; scan_s = a + l + 1; l = l + 1; continue;
                }
                /* scanning from lane l: :167-194 */
                int e = l;
                for (; e < W; e++) {
                    if (!V[e]) break;
                    if (a + e - scan_s >= 32) break;
                    if (dup[e] && e > 0) {
                        int stale = 1;
                        if (PRECISE) { stale = 0; for (int j = 0; j < e; j++) if (H[j] == H[e] && (ins >> j & 1)) stale = 1; }
                        if (stale) break;
                    }
                    if (m[e] >= 4) break;
                    ins |= 1ull << e;
                }
                if (e >= W) { next_mode = SCAN; next_a = a + W; ek = E_LEAVE; break; }
                if (!V[e]) { fin = 1; break; }
                if (a + e - scan_s >= 32) { next_mode = SCAN; next_a = a + e; ek = E_SCAN_LIMIT; break; }
                {   int stale = dup[e] && e > 0;
                    if (stale && PRECISE) { stale = 0; for (int j = 0; j < e; j++) if (H[j] == H[e] && (ins >> j & 1)) stale = 1; }
                    if (stale) { next_mode = SCAN; next_a = a + e; ek = E_SCAN_DUP; break; }
                }
                ins |= 1ull << e; /* hit */
                if (m[e] == CAP && SLOWCONT) {
                    long M = CAP; while (a + e + M < n && F[t[e] + M] == F[a + e + M]) M++;
                    f->slow++;
                    record(f, lit_from, a + e, t[e], M); lit_from = a + e + M;
                    long tgt = e + M;
                    if (a + tgt >= lim) { fin = 1; break; }
                    if (tgt >= W) { next_mode = ARR; next_a = a + tgt; ek = E_SLOW; break; }
                    scanning = 0; l = (int)tgt; continue;
                }
                if (m[e] == CAP) { slow_ip = a + e; slow_cand = t[e]; break; }
                record(f, lit_from, a + e, t[e], m[e]); lit_from = a + e + m[e];
                long tgt = e + m[e];
                if (a + tgt >= lim) { fin = 1; break; }
                if (tgt >= W) { next_mode = ARR; next_a = a + tgt; ek = E_ARR_BEYOND; break; }
                scanning = 0; l = (int)tgt;
            }
            /* ---- commit the inserts of the path; on equal hashes the later position wins (:191) */
            for (int k = 0; k < W; k++) if (ins >> k & 1) f->T[hashw(f, ld32(F + a + k))] = (uint16_t)(a + k);
            if (slow_ip >= 0) { /* long copy: full-length compare */
                f->slow++;
                long M = CAP; while (slow_ip + M < n && F[slow_cand + M] == F[slow_ip + M]) M++;
                record(f, lit_from, slow_ip, slow_cand, M); lit_from = slow_ip + M;
                e_count[E_SLOW]++; e_adv[E_SLOW] += lit_from - a;
                if (lit_from >= lim) break;
                mode = ARR; a = lit_from; continue;
            }
            if (fin) { e_count[E_FIN]++; break; }
            e_count[ek]++; e_adv[ek] += next_a - a;
            mode = next_mode; a = next_a;
        }
    }
    if (lit_from < n) emit_literal(f, lit_from, n);
    return (size_t)(f->op - f->out);
}


/* ---- W windows of 32 positions per round (compress_window.cuh, multi-warp form): all windows are
 * evaluated against the table as of the round start; the chain then passes through them in order,
 * the inserts of window k being committed before window k+1 is entered; a lane of window k+1 is
 * trusted iff the table still holds the value it looked up and no lower lane of ITS window has its hash. */
static int WW = 4;
static long st_windows, st_stale_lanes, st_entered;
static size_t compress_fragment_multi(Frag *f) {
    const uint8_t *F = f->F; const long n = f->n, lim = n - (RULES ? 15 : 16); f->lim = lim; f->op = f->out;
    memset(f->T, 0, sizeof f->T);
    long lit_from = 0;
    if (n >= 15) {
        int mode = SCAN; long a = 1, scan_s = 1;
        for (;;) {
            if (mode == SCAN && a - scan_s >= 32) {
                f->generic++;
                long k = a - scan_s; long ip = -1, cand = 0; int fin = 0;
                for (;; k++) {
                    long p = scan_s + PO[k], pn = scan_s + PO[k + 1];
                    if (pn > lim) { fin = 1; break; }
                    uint32_t h = hashw(f, ld32(F + p)); cand = f->T[h]; f->T[h] = (uint16_t)p;
                    if (ld32(F + cand) == ld32(F + p)) { ip = p; break; }
                }
                if (fin) break;
                long M = 4; while (ip + M < n && F[cand + M] == F[ip + M]) M++;
                record(f, lit_from, ip, cand, M); ip += M; lit_from = ip;
                if (ip >= lim) break;
                mode = ARR; a = ip; continue;
            }
            f->rounds++;
            if (mode == ARR) f->T[hashw(f, ld32(F + a - 1))] = (uint16_t)(a - 1);
            /* parallel evaluation of W windows against the table as of now */
            static uint32_t H[8][32], t[8][32], m[8][32]; static int V[8][32], dup[8][32];
            for (int w = 0; w < WW; w++) for (int l = 0; l < 32; l++) {
                long q = a + 32 * w + l; V[w][l] = q < lim;
                H[w][l] = V[w][l] ? hashw(f, ld32(F + q)) : (0x80000000u | (uint32_t)l);
                t[w][l] = V[w][l] ? f->T[H[w][l]] : 0; dup[w][l] = 0;
                for (int j = 0; j < l; j++) if (H[w][j] == H[w][l]) dup[w][l] = 1;
                uint32_t k = 0; if (V[w][l]) while (k < 16 && F[t[w][l] + k] == F[q + k]) k++;
                m[w][l] = k;
            }
            int fin = 0, done = 0; long slow_ip = -1, slow_cand = 0;
            long cur = a;            /* absolute position of the pending arrival / next scan probe */
            int scanning = (mode == SCAN);
            for (int w = 0; w < WW && !done; w++) {
                long base = a + 32 * w;
                if (cur >= base + 32) continue;          /* the chain jumped over this window */
                st_entered++;
                uint32_t ins = 0; int l = (int)(cur - base);
                /* :233 an arrival at lane 0 of a later window: the position before it lies in the previous window */
                if (!scanning && l == 0 && w > 0) f->T[hashw(f, ld32(F + base - 1))] = (uint16_t)(base - 1);
                /* validation: the table must still hold what the lane looked up */
                int stale[32];
                for (int l = 0; l < 32; l++) { stale[l] = V[w][l] && w > 0 && f->T[H[w][l]] != t[w][l]; st_stale_lanes += stale[l]; }
                st_windows++;
                for (;;) {
                    f->hops++;
                    if (!scanning) {
                        int untrusted = (l > 0 && dup[w][l]) || stale[l];
                        if (untrusted && !(w == 0 && l == 0)) { mode = ARR; cur = base + l; done = 1; break; }
                        /* :233 position before the arrival */
                        if (l > 0) ins |= 1u << (l - 1);
                        ins |= 1u << l;
                        if (m[w][l] >= 4) {
                            if (m[w][l] == 16) { slow_ip = base + l; slow_cand = t[w][l]; done = 1; break; }
                            record(f, lit_from, base + l, t[w][l], m[w][l]); lit_from = base + l + m[w][l];
                            long tgt = l + m[w][l];
                            if (base + tgt >= lim) { fin = 1; done = 1; break; }
                            if (tgt >= 32) { mode = ARR; scanning = 0; cur = base + tgt; break; }
                            l = (int)tgt; continue;
                        }
                        scanning = 1; scan_s = base + l + 1; l = l + 1; continue;
                    }
                    int e = l;
                    for (; e < 32; e++) {
                        if (!V[w][e]) break;
                        if (base + e - scan_s >= 32) break;
                        if ((dup[w][e] && e > 0) || stale[e]) break;
                        if (m[w][e] >= 4) break;
                        ins |= 1u << e;
                    }
                    if (e >= 32) { mode = SCAN; cur = base + 32; break; }
                    if (!V[w][e]) { fin = 1; done = 1; break; }
                    if (base + e - scan_s >= 32) { mode = SCAN; cur = base + e; done = 1; break; }
                    if ((dup[w][e] && e > 0) || stale[e]) { mode = SCAN; cur = base + e; done = 1; break; }
                    ins |= 1u << e;
                    if (m[w][e] == 16) { slow_ip = base + e; slow_cand = t[w][e]; done = 1; break; }
                    record(f, lit_from, base + e, t[w][e], m[w][e]); lit_from = base + e + m[w][e];
                    long tgt = e + m[w][e];
                    if (base + tgt >= lim) { fin = 1; done = 1; break; }
                    if (tgt >= 32) { mode = ARR; scanning = 0; cur = base + tgt; break; }
                    scanning = 0; l = (int)tgt;
                }
                for (int k = 0; k < 32; k++) if (ins >> k & 1) f->T[H[w][k]] = (uint16_t)(base + k);
            }
            if (slow_ip >= 0) {
                f->slow++;
                long M = 16; while (slow_ip + M < n && F[slow_cand + M] == F[slow_ip + M]) M++;
                record(f, lit_from, slow_ip, slow_cand, M); lit_from = slow_ip + M;
                if (lit_from >= lim) break;
                mode = ARR; a = lit_from; continue;
            }
            if (fin) break;
            a = cur;   /* mode set above */
        }
    }
    if (lit_from < n) emit_literal(f, lit_from, n);
    return (size_t)(f->op - f->out);
}

int main(int argc, char **argv) {
    init_po();
    if (getenv("WW")) WW = atoi(getenv("WW"));
    if (getenv("CAP")) CAP = (uint32_t)atoi(getenv("CAP"));
    if (getenv("RULES")) RULES = atoi(getenv("RULES"));
    if (getenv("SLOWCONT")) SLOWCONT = atoi(getenv("SLOWCONT"));
    if (getenv("PRECISE")) PRECISE = atoi(getenv("PRECISE"));
    for (int ai = 1; ai < argc; ai++) {
        FILE *fp = fopen(argv[ai], "rb"); if (!fp) { perror(argv[ai]); return 1; }
        fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
        uint8_t *buf = calloc((size_t)sz + 256, 1); if (fread(buf, 1, (size_t)sz, fp) != (size_t)sz) return 1; fclose(fp);
        uint32_t entries = sjo_hashtable_entries((uint64_t)sz);
        uint32_t shift = 32; for (uint32_t e = entries; e > 1; e >>= 1) shift--;
        long nfrag = (sz + 65535) / 65536, bad = 0; Frag *f = calloc(1, sizeof(Frag)); f->shift = shift;
        uint8_t *o1 = malloc(80000), *o2 = malloc(80000); uint16_t *tab = malloc(32768 * 2);
        for (long fr = 0; fr < nfrag; fr++) {
            long n = sz - fr * 65536 < 65536 ? sz - fr * 65536 : 65536;
            /* the kernel reads past the fragment end only into readable memory; values there must not matter */
            uint8_t *frag = calloc((size_t)n + 256, 1); memcpy(frag, buf + fr * 65536, (size_t)n); memset(frag + n, 0xA5, 200);
            f->F = frag; f->n = n; f->out = o1;
            if (RULES) {  /* GetHashTable: per fragment, up to 16384 (rules 1) or 32768 (rules 2) buckets */
                entries = 256; while (entries < (RULES == 2 ? 32768u : 16384u) && entries < (uint32_t)n) entries <<= 1;
                f->shift = 32; for (uint32_t e = entries; e > 1; e >>= 1) f->shift--;
            }
            f->hmask = entries - 1;
            size_t c1 = getenv("WW") ? compress_fragment_multi(f) : compress_fragment_window(f);
            size_t c2;
            if (RULES) { memset(tab, 0, entries * 2); c2 = sjo_compress_fragment_rules(frag, (size_t)n, o2, tab, entries, RULES); }
            else { memset(tab, 0xff, entries * 2); c2 = sjo_compress_fragment(frag, (size_t)n, o2, tab, entries); }
            if (c1 != c2 || memcmp(o1, o2, c1)) { bad++; if (bad < 4) fprintf(stderr, "%s: fragment %ld differs (%zu vs %zu)\n", argv[ai], fr, c1, c2); }
            free(frag);
        }
        if (getenv("STATS") && !getenv("WW")) {
            long tot = 0; for (int k = 0; k < E_KINDS; k++) tot += e_count[k];
            for (int k = 0; k < E_KINDS; k++) {
                printf("    %-30s %5.1f %% of rounds, advance %5.1f bytes\n", E_NAME[k], 100.0 * e_count[k] / (tot ? tot : 1), (double)e_adv[k] / (e_count[k] ? e_count[k] : 1));
                e_count[k] = e_adv[k] = 0;
            }
        }
        printf("%s: %ld fragments, %ld mismatches, rounds/frag %.0f hops/round %.2f slow/frag %.0f generic/frag %.0f windows entered/round %.2f stale lanes/window %.2f\n", argv[ai], nfrag, bad,
               (double)f->rounds / nfrag, (double)f->hops / (f->rounds ? f->rounds : 1), (double)f->slow / nfrag, (double)f->generic / nfrag, (double)st_entered / (f->rounds ? f->rounds : 1), (double)st_stale_lanes / (st_windows ? st_windows : 1)); st_entered = st_windows = st_stale_lanes = 0;
    }
    return 0;
}
