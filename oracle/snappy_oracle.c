/*
 * oracle/snappy_oracle.c -- CPU oracle: a 0-based C restatement of Snappy.jl's
 * compress / uncompress (reference: /root/reference/src/{Snappy,internal,varint}.jl).
 *
 * TEST INFRASTRUCTURE ONLY -- see snappy_oracle.h.  The product (libsnappy_b200.so) never
 * links or calls this file; it exists so the CUDA path can be checked byte for byte, and so
 * bench.py can time "the reference's algorithm on host cores" (Julia is not in this image).
 *
 * Index convention: the reference is 1-based with inclusive end indices; here every position
 * is 0-based and every end is exclusive.  Fragment-relative position q == julia_index - base_ip.
 * Compressed-byte parity is UNPINNED by the reference's tests (header of snappy_oracle.h).
 */
#include "snappy_oracle.h"

#include <string.h>

#define K_BLOCK_SIZE 65536u          /* internal.jl:31 */
#define K_INPUT_MARGIN_BYTES 15      /* internal.jl:32 */
#define K_MAX_HASH_TABLE_SIZE 16384u /* internal.jl:33 */

/* fastmemory.jl:4-21 -- unaligned little-endian loads (host is 64-bit LE, internal.jl:9-10) */
static inline uint32_t load32u(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t load64u(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

/* Snappy.jl:80-82 */
size_t sjo_maxlength_compressed(size_t n) { return 32 + n + n / 6; }

/* varint.jl:46-69 -- returns the number of bytes written (1..5) */
int sjo_encode32(uint8_t *buf, uint32_t v) {
    int k = 0;
    if (v < (1u << 7)) {
        buf[k++] = (uint8_t)v;
    } else if (v < (1u << 14)) {
        buf[k++] = (uint8_t)(v | 128);
        buf[k++] = (uint8_t)(v >> 7);
    } else if (v < (1u << 21)) {
        buf[k++] = (uint8_t)(v | 128);
        buf[k++] = (uint8_t)((v >> 7) | 128);
        buf[k++] = (uint8_t)(v >> 14);
    } else if (v < (1u << 28)) {
        buf[k++] = (uint8_t)(v | 128);
        buf[k++] = (uint8_t)((v >> 7) | 128);
        buf[k++] = (uint8_t)((v >> 14) | 128);
        buf[k++] = (uint8_t)(v >> 21);
    } else {
        buf[k++] = (uint8_t)(v | 128);
        buf[k++] = (uint8_t)((v >> 7) | 128);
        buf[k++] = (uint8_t)((v >> 14) | 128);
        buf[k++] = (uint8_t)((v >> 21) | 128);
        buf[k++] = (uint8_t)(v >> 28);
    }
    return k;
}

/* varint.jl:12-37 -- `offset` is the 0-based start; *next is the 0-based index past the varint.
 * Fails when the buffer ends inside the varint, or the 5th byte is >= 0x10. */
int sjo_parse32(const uint8_t *buf, size_t len, size_t offset, uint32_t *value, size_t *next) {
    uint32_t result = 0;
    for (int i = 0; i < 5; i++) {
        if (offset >= len) return SJO_BAD_VARINT;          /* varint.jl:13,18,22,26,30 */
        uint32_t b = buf[offset++];
        result |= (b & 0x7f) << (7 * i);                   /* 5th byte: bits above 31 fall off */
        if (i < 4 ? (b < 0x80) : (b < 0x10)) {             /* varint.jl:17..33 */
            *value = result;
            *next = offset;
            return SJO_OK;
        }
    }
    return SJO_BAD_VARINT;                                 /* varint.jl:35-36 */
}

/* internal.jl:344-387 (64-bit LE variant).  Longest common prefix of a[i1..] and a[i2..limit),
 * `limit` exclusive (== the reference's inclusive 1-based `limit`). */
size_t sjo_find_match_length(const uint8_t *a, size_t i1, size_t i2, size_t limit) {
    size_t matched = 0;
    if (i2 + 8 <= limit) {                                 /* :356 */
        uint64_t a1 = load64u(a + i1), a2 = load64u(a + i2);
        if (a1 != a2) return (size_t)(__builtin_ctzll(a1 ^ a2) >> 3); /* :360 */
        i2 += 8;
        matched = 8;
    }
    while (i2 + 8 <= limit) {                              /* :371 */
        uint64_t x = load64u(a + i2) ^ load64u(a + i1 + matched);
        if (x == 0) {
            i2 += 8;
            matched += 8;
        } else {
            return matched + (size_t)(__builtin_ctzll(x) >> 3); /* :376-379 */
        }
    }
    while (i2 < limit && a[i1 + matched] == a[i2]) {       /* :382-385 */
        i2++;
        matched++;
    }
    return matched;
}

/* internal.jl:107-113 -- table size from the TOTAL input length (Snappy.jl:27) */
uint32_t sjo_hashtable_entries(uint64_t total_len) {
    uint32_t htsize = 256;
    while (htsize < K_MAX_HASH_TABLE_SIZE && htsize < total_len) htsize <<= 1;
    return htsize;
}

/* internal.jl:252-287.  The fast path (:265-269) writes the same tag + bytes as the slow path
 * plus up to 15 garbage bytes that the next emit overwrites; only the exact bytes are written. */
static uint8_t *emit_literal(uint8_t *op, const uint8_t *lit, size_t len) {
    uint32_t n = (uint32_t)(len - 1);
    if (len < 60) {                                        /* :271  (so len==60 takes 2 bytes) */
        *op++ = (uint8_t)(n << 2);
    } else {
        uint8_t *base = op;
        int count = 0;
        while (n > 0) {                                    /* :279-282 */
            *++op = (uint8_t)n;
            n >>= 8;
            count++;
        }
        *base = (uint8_t)((59 + count) << 2);              /* :283 */
        op++;
    }
    memcpy(op, lit, len);                                  /* :285 */
    return op + len;
}

/* internal.jl:289-304 */
static uint8_t *emit_copy_upto_64(uint8_t *op, uint32_t offset, uint32_t len) {
    if (len < 12 && offset < 2048) {
        *op++ = (uint8_t)(1 + ((len - 4) << 2) + ((offset >> 3) & 0xe0));
        *op++ = (uint8_t)(offset & 0xff);
    } else {
        uint32_t u = 2 + ((len - 1) << 2) + (offset << 8); /* reference stores 4, advances 3 */
        *op++ = (uint8_t)u;
        *op++ = (uint8_t)(u >> 8);
        *op++ = (uint8_t)(u >> 16);
    }
    return op;
}

/* internal.jl:306-329 */
static uint8_t *emit_copy(uint8_t *op, uint32_t offset, uint32_t len) {
    if (len < 12) return emit_copy_upto_64(op, offset, len);
    while (len >= 68) {
        op = emit_copy_upto_64(op, offset, 64);
        len -= 64;
    }
    if (len > 64) {
        op = emit_copy_upto_64(op, offset, 60);
        len -= 60;
    }
    return emit_copy_upto_64(op, offset, len);             /* :326 may pick the 2-byte form */
}

/* internal.jl:127-250.  `table` holds position-1 in u16 with 0xffff == empty, so that
 * (uint16)(t + 1) is the candidate position and an empty slot yields position 0 (:177-191). */
size_t sjo_compress_fragment(const uint8_t *F, size_t n_, uint8_t *out, uint16_t *table,
                             uint32_t entries) {
    const long n = (long)n_;
    uint32_t shift = 32;
    for (uint32_t e = entries; e > 1; e >>= 1) shift--;    /* :128  32 - log2floor(entries) */
    uint8_t *op = out;
    long ip = 0, next_emit = 0, candidate = 0;
    const long ip_limit = n - 1 - K_INPUT_MARGIN_BYTES;    /* :131  (inclusive end n-1) - 15 */

#define HASH(q) ((uint32_t)(load32u(F + (q)) * 0x1e35a7bdu) >> shift) /* :94 */

    if (n >= K_INPUT_MARGIN_BYTES) {                       /* :133 */
        for (;;) {
            uint32_t skip = 32;                            /* :162 */
            ip += 1;
            uint32_t next_hash = HASH(ip);                 /* :163 */
            long next_ip = ip;
            for (;;) {                                     /* :167-194 */
                ip = next_ip;
                uint32_t cur_hash = next_hash;
                uint32_t between = skip >> 5;
                skip += between;
                next_ip = ip + between;
                if (next_ip > ip_limit) goto emit_remainder; /* :175 */
                next_hash = HASH(next_ip);
                candidate = (uint16_t)(table[cur_hash] + 1); /* :190 */
                table[cur_hash] = (uint16_t)(ip - 1);        /* :191 */
                if (load32u(F + candidate) == load32u(F + ip)) break; /* :193 */
            }
            op = emit_literal(op, F + next_emit, (size_t)(ip - next_emit)); /* :200 */
            for (;;) {                                     /* :211-239 */
                long matched = 4 + (long)sjo_find_match_length(F, (size_t)candidate + 4,
                                                               (size_t)ip + 4, (size_t)n); /* :216 */
                op = emit_copy(op, (uint32_t)(ip - candidate), (uint32_t)matched);        /* :217 */
                ip += matched;
                next_emit = ip;
                if (ip >= ip_limit) goto emit_remainder;   /* :222 */
                uint32_t prev_hash = HASH(ip - 1);         /* :228 */
                uint32_t input_bytes = load32u(F + ip);
                uint32_t cur_hash = (input_bytes * 0x1e35a7bdu) >> shift;
                table[prev_hash] = (uint16_t)(ip - 1 - 1); /* :233 */
                candidate = (uint16_t)(table[cur_hash] + 1); /* :234 */
                table[cur_hash] = (uint16_t)(ip - 1);      /* :235 */
                if (input_bytes != load32u(F + candidate)) break; /* :238 */
            }
        }
    }
emit_remainder:
    if (next_emit < n)                                     /* :244 */
        op = emit_literal(op, F + next_emit, (size_t)(n - next_emit));
#undef HASH
    return (size_t)(op - out);
}

/* Snappy.jl:29-33 restricted to fragments [first_frag, first_frag+nfrag); `in` points at the
 * START OF THE STREAM.  The loop `for i in 0:65536:n` also visits an empty trailing fragment
 * when n is a multiple of 65536 (it emits nothing), so fragment counts here use ceil(n/65536). */
size_t sjo_compress_fragments(const uint8_t *in, uint64_t total_len, size_t first_frag,
                              size_t nfrag, uint8_t *out, uint32_t *frag_sizes) {
    uint16_t table[K_MAX_HASH_TABLE_SIZE];
    uint32_t entries = sjo_hashtable_entries(total_len);   /* Snappy.jl:27 */
    uint8_t *op = out;
    for (size_t f = first_frag; f < first_frag + nfrag; f++) {
        uint64_t s = (uint64_t)f * K_BLOCK_SIZE;
        if (s >= total_len) break;
        size_t n = (size_t)((total_len - s < K_BLOCK_SIZE) ? (total_len - s) : K_BLOCK_SIZE);
        memset(table, 0xff, entries * sizeof(uint16_t));   /* Snappy.jl:30 */
        size_t c = sjo_compress_fragment(in + s, n, op, table, entries);
        if (frag_sizes) frag_sizes[f - first_frag] = (uint32_t)c;
        op += c;
    }
    return (size_t)(op - out);
}

/* Snappy.jl:20-36.  *out_len: in = capacity of out, out = compressed length. */
int sjo_compress(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len) {
    if (n > 0xffffffffull) return SJO_INPUT_TOO_LARGE;     /* :21 */
    if (*out_len < sjo_maxlength_compressed(n)) return SJO_BUFFER_TOO_SMALL;
    size_t k = (size_t)sjo_encode32(out, (uint32_t)n);     /* :26 */
    size_t nfrag = (n + K_BLOCK_SIZE - 1) / K_BLOCK_SIZE;
    k += sjo_compress_fragments(in, n, 0, nfrag, out + k, NULL);
    *out_len = k;                                          /* :35 */
    return SJO_OK;
}

/* ---------------------------------------------------------------------------------------------
 * libsnappy rules (option `rules` of the product).
 * NOT a restatement of the reference: this is Google snappy's compressor (snappy.cc: Compress,
 * WorkingMemory::GetHashTable, CompressFragment, EmitLiteral, EmitCopy), which Snappy.jl ports with the
 * deviations of appendix B.4.  Differences from sjo_compress above, all of them:
 *   ip_limit = n - 15 (kInputMarginBytes; the reference's inclusive end makes it n - 16);
 *   a literal of exactly 60 bytes keeps the one-byte tag (n = len - 1 < 60);
 *   the table is sized from each fragment's length (GetHashTable(num_to_read));
 *   rules = 1 (libsnappy <= 1.1.7): bucket = (bytes * 0x1e35a7bd) >> (32 - log2(table size)), <= 16384 buckets;
 *   rules = 2 (Google snappy >= 1.1.9): bucket = ((bytes * 0x1e35a7bd) >> 17) & (table size - 1), <= 32768.
 * PINNING: rules = 2 is byte-identical to pyarrow's bundled Google snappy on every fixture file and on
 * fuzzed inputs (tests/test_libsnappy_rules.py); rules = 1 differs from it only in the bucket function and
 * table cap (no libsnappy 1.1.7 binary exists in this image: that one line is unpinned).
 * The table holds positions directly, 0 = empty = fragment start. */
size_t sjo_compress_fragment_rules(const uint8_t *F, size_t n_, uint8_t *out, uint16_t *table,
                                   uint32_t entries, int rules) {
    const long n = (long)n_;
    uint32_t shift = 32;
    for (uint32_t e = entries; e > 1; e >>= 1) shift--;
    const uint32_t mask = entries - 1;
#define BUCKET(w) (rules == 2 ? ((((uint32_t)(w) * 0x1e35a7bdu) >> 17) & mask) \
                              : (((uint32_t)(w) * 0x1e35a7bdu) >> shift))
    uint8_t *op = out;
    long ip = 0, next_emit = 0, candidate = 0;
    if (n >= K_INPUT_MARGIN_BYTES) {
        const long ip_limit = n - K_INPUT_MARGIN_BYTES;
        for (;;) {
            uint32_t skip = 32;
            ip += 1;
            uint32_t next_hash = BUCKET(load32u(F + ip));
            long next_ip = ip;
            for (;;) {
                ip = next_ip;
                uint32_t h = next_hash;
                uint32_t between = skip >> 5;
                skip += between;
                next_ip = ip + between;
                if (next_ip > ip_limit) goto emit_remainder;
                next_hash = BUCKET(load32u(F + next_ip));
                candidate = table[h];
                table[h] = (uint16_t)ip;
                if (load32u(F + candidate) == load32u(F + ip)) break;
            }
            {   /* EmitLiteral: n = len - 1 < 60 takes the short tag */
                size_t len = (size_t)(ip - next_emit);
                if (len == 60) {
                    *op++ = (uint8_t)(59u << 2);
                    memcpy(op, F + next_emit, 60);
                    op += 60;
                } else {
                    op = emit_literal(op, F + next_emit, len);
                }
            }
            for (;;) {
                long matched = 4;
                while (ip + matched < n && F[candidate + matched] == F[ip + matched]) matched++;
                op = emit_copy(op, (uint32_t)(ip - candidate), (uint32_t)matched);
                ip += matched;
                next_emit = ip;
                if (ip >= ip_limit) goto emit_remainder;
                table[BUCKET(load32u(F + ip - 1))] = (uint16_t)(ip - 1);
                uint32_t cur = load32u(F + ip);
                uint32_t h = BUCKET(cur);
                candidate = table[h];
                table[h] = (uint16_t)ip;
                if (cur != load32u(F + candidate)) break;
            }
        }
    }
emit_remainder:
    if (next_emit < n) {
        size_t len = (size_t)(n - next_emit);
        if (len == 60) {
            *op++ = (uint8_t)(59u << 2);
            memcpy(op, F + next_emit, 60);
            op += 60;
        } else {
            op = emit_literal(op, F + next_emit, len);
        }
    }
#undef BUCKET
    return (size_t)(op - out);
}

/* snappy.cc Compress(): varint, then fragments of 65536 bytes, each with its own table size */
int sjo_compress_rules(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len, int rules) {
    if (rules == 0) return sjo_compress(in, n, out, out_len);
    if (n > 0xffffffffull) return SJO_INPUT_TOO_LARGE;
    if (*out_len < sjo_maxlength_compressed(n)) return SJO_BUFFER_TOO_SMALL;
    static _Thread_local uint16_t table[2 * K_MAX_HASH_TABLE_SIZE];
    const uint32_t cap = rules == 2 ? 2 * K_MAX_HASH_TABLE_SIZE : K_MAX_HASH_TABLE_SIZE;
    uint8_t *op = out + sjo_encode32(out, (uint32_t)n);
    for (size_t s = 0; s < n; s += K_BLOCK_SIZE) {
        size_t fn = (n - s < K_BLOCK_SIZE) ? (n - s) : K_BLOCK_SIZE;
        uint32_t entries = 256;
        while (entries < cap && entries < fn) entries <<= 1;
        memset(table, 0, entries * sizeof(uint16_t));
        op += sjo_compress_fragment_rules(in + s, fn, op, table, entries, rules);
    }
    *out_len = (size_t)(op - out);
    return SJO_OK;
}

/* Snappy.jl:90-92 */
int sjo_uncompressed_length(const uint8_t *in, size_t n, size_t *result) {
    uint32_t v;
    size_t next;
    int rc = sjo_parse32(in, n, 0, &v, &next);
    if (rc == SJO_OK) *result = v;
    return rc;
}

/* internal.jl:47-80 regenerated from the format definition instead of the literal table */
uint16_t sjo_char_table_entry(uint32_t c) {
    uint32_t kind = c & 3, hi = c >> 2;
    if (kind == 0) return (uint16_t)(hi < 60 ? (hi + 1) : (1 | ((hi - 59) << 11)));
    if (kind == 1) return (uint16_t)((4 + (hi & 7)) | ((c >> 5) << 8) | (1u << 11));
    return (uint16_t)((hi + 1) | ((kind == 2 ? 2u : 4u) << 11));
}

/* Snappy.jl:46-52 + internal.jl:411-527.  *out_len: in = capacity, out = claimed length. */
int sjo_uncompress_ex(const uint8_t *in, size_t L, uint8_t *out, size_t *out_len, size_t *err_op) {
    static const uint32_t wordmask[5] = {0u, 0xffu, 0xffffu, 0xffffffu, 0xffffffffu}; /* :83-85 */
    uint32_t claimed;
    size_t ip;
    if (err_op) *err_op = 0;
    int rc = sjo_parse32(in, L, 0, &claimed, &ip);         /* Snappy.jl:47 */
    if (rc != SJO_OK) return rc;
    const size_t n = claimed;
    if (*out_len < n) return SJO_BUFFER_TOO_SMALL;
    *out_len = n;
    size_t op = 0;
    while (ip + 1 < L) {                                   /* internal.jl:416  ip < ip_limit, strict */
        uint32_t c = in[ip++];
        uint32_t tag = 0;                                  /* :426-430 zero-padded 4-byte trailer */
        for (int k = 0; k < 4 && ip + (size_t)k < L; k++) tag |= (uint32_t)in[ip + k] << (8 * k);
        uint32_t entry = sjo_char_table_entry(c);          /* :435 */
        uint32_t len = entry & 0xff, taglen = entry >> 11;
        uint32_t trailer = tag & wordmask[taglen];         /* :438 */
        ip += taglen;                                      /* may run past L; see avail_in below */
        if ((c & 3) != 0) {                                /* :458 copy */
            uint32_t offset = (entry & 0x700) + trailer;
            size_t avail_out = n - op;
            if (offset == 0 || (size_t)offset > op) {      /* :499 */
                if (err_op) *err_op = op;
                return SJO_CORRUPT_COPY_OFFSET;
            }
            if (avail_out < len) {                         /* :500,:505 fast path needs avail>=16>=len */
                if (err_op) *err_op = op;
                return SJO_CORRUPT_COPY_LENGTH;
            }
            for (uint32_t i = 0; i < len; i++) out[op + i] = out[op - offset + i]; /* :477-481 */
            op += len;
        } else {                                           /* :462 literal */
            uint32_t llen = len + trailer;                 /* UInt32 wrap-around, as in Julia */
            size_t avail_out = n - op;
            long long avail_in = (long long)L - (long long)ip;
            if (avail_out < llen || avail_in < (long long)llen) { /* :518 */
                if (err_op) *err_op = op;
                return SJO_CORRUPT_LITERAL;
            }
            memcpy(out + op, in + ip, llen);
            op += llen;
            ip += llen;
        }
    }
    if (op != n) {                                         /* Snappy.jl:50 */
        if (err_op) *err_op = op;
        return SJO_INVALID_INPUT;
    }
    return SJO_OK;
}

int sjo_uncompress(const uint8_t *in, size_t L, uint8_t *out, size_t *out_len) {
    return sjo_uncompress_ex(in, L, out, out_len, NULL);
}
