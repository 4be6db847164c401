/*
 * oracle/snappy_oracle.h -- CPU oracle for the Snappy.jl compress/uncompress path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (snappy.jl_b200/, include/,
 * libsnappy_b200.so) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker / baseline.
 *
 * Every function is a 0-based C restatement of a Julia routine in /root/reference/src
 * (cited at each definition in snappy_oracle.c).
 *
 * PARITY PINNING (see DESIGN.md "Oracle"):
 *   pinned by the reference's own tests: find_match_length (45 KATs), varint32 encode/parse
 *   (31 round trips + 3 rejects), the 14 must-throw streams, round-trip identity on the 15
 *   fixture files + edge strings + dictionary fuzz.
 *   COMPRESSED BYTES: parity unpinned -- the reference holds no golden compressed vector
 *   produced by Snappy.jl and Julia cannot run in this image; byte-exactness rests on two
 *   independently written restatements (this file and oracle/py_restatement.py) agreeing
 *   with each other and with the SHA-256 table in SURVEY.md Appendix C.
 */
#ifndef SNAPPY_ORACLE_H
#define SNAPPY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes; one per error(...) site of the reference. */
enum {
    SJO_OK = 0,
    SJO_INPUT_TOO_LARGE = 1,     /* Snappy.jl:21  "Input too large." */
    SJO_INVALID_INPUT = 2,       /* Snappy.jl:50  "Invalid input." */
    SJO_CORRUPT_COPY_OFFSET = 3, /* internal.jl:499 */
    SJO_CORRUPT_COPY_LENGTH = 4, /* internal.jl:505 */
    SJO_CORRUPT_LITERAL = 5,     /* internal.jl:518 */
    SJO_BAD_VARINT = 6,          /* varint.jl:36 */
    SJO_BUFFER_TOO_SMALL = 7     /* not a reference error: caller's buffer is too small */
};

size_t sjo_maxlength_compressed(size_t n);
int sjo_encode32(uint8_t *buf, uint32_t value);
int sjo_parse32(const uint8_t *buf, size_t len, size_t offset, uint32_t *value, size_t *next);
size_t sjo_find_match_length(const uint8_t *a, size_t i1, size_t i2, size_t limit);
uint32_t sjo_hashtable_entries(uint64_t total_len);

/* compress one <=64 KiB fragment with a table of `entries` u16 slots (all 0xffff on entry) */
size_t sjo_compress_fragment(const uint8_t *frag, size_t n, uint8_t *out, uint16_t *table,
                             uint32_t entries);
/* compress fragments [first_frag, first_frag+nfrag) of a stream whose TOTAL length is total_len;
 * no varint header; frag_sizes (optional) receives each fragment's compressed size. */
size_t sjo_compress_fragments(const uint8_t *in, uint64_t total_len, size_t first_frag,
                              size_t nfrag, uint8_t *out, uint32_t *frag_sizes);
int sjo_compress(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len);
/* Google snappy's own rules (NOT the reference's; see the .c file): rules 0 = sjo_compress, 1 = libsnappy
 * <= 1.1.7, 2 = Google snappy >= 1.1.9 (pinned against pyarrow's bundled codec) */
size_t sjo_compress_fragment_rules(const uint8_t *frag, size_t n, uint8_t *out, uint16_t *table,
                                   uint32_t entries, int rules);
int sjo_compress_rules(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len, int rules);
/* CHAR_TABLE[c] (internal.jl:47-80) as this oracle regenerates it; tests hold it to the reference's literal table */
uint16_t sjo_char_table_entry(uint32_t c);
int sjo_uncompressed_length(const uint8_t *in, size_t n, size_t *result);
int sjo_uncompress(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len);
/* same as sjo_uncompress, additionally reports the output position at which the error fired */
int sjo_uncompress_ex(const uint8_t *in, size_t n, uint8_t *out, size_t *out_len,
                      size_t *err_op);

#ifdef __cplusplus
}
#endif
#endif
