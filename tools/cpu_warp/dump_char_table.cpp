// dump_char_table.cpp -- prints the 256 CHAR_TABLE entries (src/internal.jl:47-80) and the 5 WORDMASK values
// (:83-85) as the CUDA decoders evaluate them arithmetically: decode_tag (exact decoder, parse kernels) and
// decode_tag_fast (window parse of the indexed decoder), csrc/decompress.cuh.  TEST INFRASTRUCTURE: the real header
// compiled for the CPU (-DSB200_CPU_EMU); tests/test_oracle.py compares the output with tests/golden/char_table.json.
//   g++ -O1 -std=c++17 -DSB200_CPU_EMU -Itools/cpu_warp tools/cpu_warp/dump_char_table.cpp
#include "cuda_shim.h"
#include "../../snappy.jl_b200/csrc/decompress.cuh"

#include <cstdio>

namespace sb200 {
u8 smem[16];
}

int main() {
    using namespace sb200;
    for (int variant = 0; variant < 2; variant++) {
        for (u32 c = 0; c < 256; c++) {
            u32 len, offset, extra;
            if (variant == 0) {
                const Element e = decode_tag(c, 0u);  // trailer 0: the table's own length / offset-high fields
                len = e.len;
                offset = e.offset;
                extra = e.extra;
            } else {
                bool is_copy;
                decode_tag_fast(c, 0u, len, offset, extra, is_copy);
            }
            printf("%u ", len | offset | (extra << 11));
        }
        printf("\n");
    }
    // WORDMASK: a literal tag with `extra` length bytes keeps exactly the low 8 * extra bits of the trailer
    for (int variant = 0; variant < 2; variant++) {
        printf("0 ");
        for (u32 extra = 1; extra <= 4; extra++) {
            const u32 c = (59 + extra) << 2;
            u32 len;
            if (variant == 0) len = decode_tag(c, 0xffffffffu).len;
            else {
                u32 offset, ex;
                bool is_copy;
                decode_tag_fast(c, 0xffffffffu, len, offset, ex, is_copy);
            }
            printf("%u ", len - 1u);  // 1 + (trailer & mask), UInt32 arithmetic
        }
        printf("\n");
    }
    return 0;
}
