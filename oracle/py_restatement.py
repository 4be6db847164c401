"""Second, independently written restatement of Snappy.jl -- pure Python, 1-BASED like the Julia.

TEST INFRASTRUCTURE ONLY (same rules as snappy_oracle.h).  Purpose: the reference's tests pin no
compressed bytes and Julia cannot run here, so the C oracle (snappy_oracle.c, written 0-based from
SURVEY Appendix B) is cross-checked against this file, which was transcribed separately from the
Julia source keeping its 1-based inclusive indices.  Two transcriptions that agree byte for byte on
the fixtures and on fuzz, and with SURVEY Appendix C's SHA-256 table, are the best available pin.
Slow (pure Python loops): use on inputs up to a few hundred KB.
"""

M32 = 0xFFFFFFFF


class SnappyError(Exception):
    pass


class A1:
    """1-based view of a bytearray (a[1] is the first byte), like a Julia Vector{UInt8}."""

    def __init__(self, data):
        self.d = data

    def __getitem__(self, i):
        return self.d[i - 1]

    def __setitem__(self, i, v):
        self.d[i - 1] = v & 0xFF

    def endof(self):
        return len(self.d)


def load32u(a, i):  # fastmemory.jl:4
    d = a.d
    j = i - 1
    return d[j] | (d[j + 1] << 8) | (d[j + 2] << 16) | (d[j + 3] << 24)


def load64u(a, i):  # fastmemory.jl:7
    return int.from_bytes(a.d[i - 1:i + 7], "little")


def trailing_zeros(x):
    return (x & -x).bit_length() - 1


# ---- varint.jl ----
def parse32(buf, offset):  # varint.jl:12-37 ; returns (value, index past varint), 1-based
    n = len(buf.d)
    if offset > n:
        raise SnappyError("Could not decode varint32.")
    b = buf[offset]
    result = b & 0x7F
    if b < 0x80:
        return result, offset + 1
    for shift in (7, 14, 21):
        offset += 1
        if offset > n:
            raise SnappyError("Could not decode varint32.")
        b = buf[offset]
        result |= (b & 0x7F) << shift
        if b < 0x80:
            return result & M32, offset + 1
    offset += 1
    if offset > n:
        raise SnappyError("Could not decode varint32.")
    b = buf[offset]
    result = (result | ((b & 0x7F) << 28)) & M32
    if b < 0x10:
        return result, offset + 1
    raise SnappyError("Could not decode varint32.")


def encode32(buf, offset, value):  # varint.jl:46-69 ; returns index past the varint
    if value < (1 << 7):
        buf[offset] = value
    elif value < (1 << 14):
        buf[offset] = value | 128
        offset += 1
        buf[offset] = value >> 7
    elif value < (1 << 21):
        buf[offset] = value | 128
        offset += 1
        buf[offset] = (value >> 7) | 128
        offset += 1
        buf[offset] = value >> 14
    elif value < (1 << 28):
        buf[offset] = value | 128
        offset += 1
        buf[offset] = (value >> 7) | 128
        offset += 1
        buf[offset] = (value >> 14) | 128
        offset += 1
        buf[offset] = value >> 21
    else:
        buf[offset] = value | 128
        offset += 1
        buf[offset] = (value >> 7) | 128
        offset += 1
        buf[offset] = (value >> 14) | 128
        offset += 1
        buf[offset] = (value >> 21) | 128
        offset += 1
        buf[offset] = value >> 28
    return offset + 1


# ---- internal.jl ----
K_BLOCK_SIZE = 1 << 16
K_INPUT_MARGIN_BYTES = 15
K_MAX_HASH_TABLE_SIZE = 1 << 14


def hashdword(b, shift):  # internal.jl:94
    return ((b * 0x1E35A7BD) & M32) >> shift


def alloc_hashtable(n):  # internal.jl:107-113
    htsize = 256
    while htsize < K_MAX_HASH_TABLE_SIZE and htsize < n:
        htsize <<= 1
    return [0] * htsize


def find_match_length(a, i1, i2, limit):  # internal.jl:344-387 (1-based, limit inclusive)
    matched = 0
    if i2 <= limit - 7:
        a1 = load64u(a, i1)
        a2 = load64u(a, i2)
        if a1 != a2:
            return trailing_zeros(a1 ^ a2) >> 3
        i2 += 8
        matched = 8
    while i2 <= limit - 7:
        if load64u(a, i2) == load64u(a, i1 + matched):
            i2 += 8
            matched += 8
        else:
            x = load64u(a, i2) ^ load64u(a, i1 + matched)
            return matched + (trailing_zeros(x) >> 3)
    while i2 <= limit and a[i1 + matched] == a[i2]:
        i2 += 1
        matched += 1
    return matched


def emit_literal(output, outputindex, inp, inputindex, length):  # internal.jl:252-287
    n = (length - 1) & M32
    if length < 60:
        output[outputindex] = (n << 2) & 0xFF
    else:
        count = 0
        base = outputindex
        while n > 0:
            outputindex += 1
            output[outputindex] = n & 0xFF
            n >>= 8
            count += 1
        output[base] = ((59 + count) << 2) & 0xFF
    outputindex += 1
    output.d[outputindex - 1:outputindex - 1 + length] = inp.d[inputindex - 1:inputindex - 1 + length]
    return outputindex + length


def emit_copy_upto_64(output, outputindex, offset, length):  # internal.jl:289-304
    if length < 12 and offset < 2048:
        output[outputindex] = (1 + ((length - 4) << 2) + ((offset >> 3) & 0xE0)) & 0xFF
        output[outputindex + 1] = offset & 0xFF
        return outputindex + 2
    u = (2 + ((length - 1) << 2) + (offset << 8)) & M32
    output[outputindex] = u & 0xFF
    output[outputindex + 1] = (u >> 8) & 0xFF
    output[outputindex + 2] = (u >> 16) & 0xFF
    return outputindex + 3


def emit_copy(output, outputindex, offset, length):  # internal.jl:306-329
    if length < 12:
        return emit_copy_upto_64(output, outputindex, offset, length)
    while length >= 68:
        outputindex = emit_copy_upto_64(output, outputindex, offset, 64)
        length -= 64
    if length > 64:
        outputindex = emit_copy_upto_64(output, outputindex, offset, 60)
        length -= 60
    return emit_copy_upto_64(output, outputindex, offset, length)


def compress_fragment(output, outputindex, inp, ip, ip_end, table):  # internal.jl:127-250
    shift = 32 - (len(table).bit_length() - 1)
    candidate = next_emit = base_ip = ip
    input_size = ip_end - ip + 1
    ip_limit = ip_end - K_INPUT_MARGIN_BYTES
    remainder = False
    if input_size >= K_INPUT_MARGIN_BYTES:
        while not remainder:
            skip = 32
            ip += 1
            next_hash = hashdword(load32u(inp, ip), shift)
            next_ip = ip
            while True:
                ip = next_ip
                cur_hash = next_hash
                bytes_between = skip >> 5
                skip += bytes_between
                next_ip = ip + bytes_between
                if next_ip > ip_limit:
                    remainder = True
                    break
                next_hash = hashdword(load32u(inp, next_ip), shift)
                candidate = base_ip + ((table[cur_hash] + 1) & 0xFFFF)
                table[cur_hash] = ((ip - base_ip) - 1) & 0xFFFF
                if load32u(inp, candidate) == load32u(inp, ip):
                    break
            if remainder:
                break
            outputindex = emit_literal(output, outputindex, inp, next_emit, ip - next_emit)
            while True:
                matched = 4 + find_match_length(inp, candidate + 4, ip + 4, ip_end)
                outputindex = emit_copy(output, outputindex, ip - candidate, matched)
                ip += matched
                next_emit = ip
                if ip >= ip_limit:
                    remainder = True
                    break
                prev_hash = hashdword(load32u(inp, ip - 1), shift)
                input_bytes = load32u(inp, ip)
                cur_hash = hashdword(input_bytes, shift)
                table[prev_hash] = (ip - base_ip - 1 - 1) & 0xFFFF
                candidate = base_ip + ((table[cur_hash] + 1) & 0xFFFF)
                table[cur_hash] = (ip - base_ip - 1) & 0xFFFF
                if input_bytes != load32u(inp, candidate):
                    break
    if next_emit <= ip_end:
        outputindex = emit_literal(output, outputindex, inp, next_emit, ip_end - next_emit + 1)
    return outputindex


def compress(data):  # Snappy.jl:20-36
    data = bytearray(data)
    if len(data) > M32:
        raise SnappyError("Input too large.")
    sourcelen = len(data)
    inp = A1(data)
    output = A1(bytearray(32 + sourcelen + sourcelen // 6))
    outputindex = encode32(output, 1, sourcelen)
    table = alloc_hashtable(sourcelen)
    for i in range(0, sourcelen + 1, K_BLOCK_SIZE):
        for k in range(len(table)):
            table[k] = 0xFFFF
        outputindex = compress_fragment(output, outputindex, inp, i + 1,
                                        min(i + K_BLOCK_SIZE, sourcelen), table)
    return bytes(output.d[:outputindex - 1])


def _char_table():
    t = []
    for c in range(256):
        kind, hi = c & 3, c >> 2
        if kind == 0:
            t.append(hi + 1 if hi < 60 else (1 | ((hi - 59) << 11)))
        elif kind == 1:
            t.append((4 + (hi & 7)) | ((c >> 5) << 8) | (1 << 11))
        else:
            t.append((hi + 1) | ((2 if kind == 2 else 4) << 11))
    return t


CHAR_TABLE = _char_table()
WORDMASK = [0, 0xFF, 0xFFFF, 0xFFFFFF, 0xFFFFFFFF]


def uncompress(data):  # Snappy.jl:46-52, internal.jl:411-527
    inp = A1(bytearray(data))
    output_size, ip = parse32(inp, 1)
    out = A1(bytearray(output_size))
    ip_limit = inp.endof()
    op = 1
    while ip < ip_limit:
        c = inp[ip]
        ip += 1
        chunk = bytes(inp.d[ip - 1:ip_limit]) + b"\0\0\0\0"
        tag = int.from_bytes(chunk[:4], "little")
        entry = CHAR_TABLE[c]
        length = entry & 0xFF
        taglen = entry >> 11
        trailer = tag & WORDMASK[taglen]
        ip += taglen
        if (c & 3) != 0:
            offset = (entry & 0x700) + trailer
            avail_out = out.endof() - op + 1
            if (op - 1) <= ((offset - 1) & M32):
                raise SnappyError("Invalid input: corrupt copy offset")
            if not (length <= 16 and offset >= 8 and avail_out >= 16):
                if avail_out < length:
                    raise SnappyError("Invalid input: corrupt copy length")
            for i in range(length):
                out[op + i] = out[op - offset + i]
            op += length
        else:
            length = (length + trailer) & M32
            avail_out = out.endof() - op + 1
            avail_in = inp.endof() - ip + 1
            if avail_out < length or avail_in < length:
                raise SnappyError("Invalid input: corrupt literal")
            out.d[op - 1:op - 1 + length] = inp.d[ip - 1:ip - 1 + length]
            op += length
            ip += length
    if output_size != op - 1:
        raise SnappyError("Invalid input.")
    return bytes(out.d)
