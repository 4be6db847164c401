# varint.jl -- varint32 header codec kept on the host, as in the reference (src/varint.jl).
# Same names, arguments, return values and failure as the reference's parse32 / encode32!
# (1-based offsets; both return the index just past the varint), written as loops.

const K_MAX32 = 5   # the maximum length in bytes of a varint32

"""
    parse32(buf::Vector{UInt8}, offset::Integer)

Read a varint32 from `buf` at `offset`; returns `(value::UInt32, next_index)`.
Fails with "Could not decode varint32." when the buffer ends inside the varint or the fifth byte
is >= 0x10 (reference: src/varint.jl:12-37).
"""
function parse32(buf::Vector{UInt8}, offset::Integer)
    result = UInt32(0)
    for i in 0:K_MAX32-1
        offset > length(buf) && break
        b = UInt32(buf[offset])
        result |= (b & 0x7f) << (7 * i)      # UInt32 shift: bits above 31 of the 5th byte fall off
        offset += 1
        if (i < 4 && b < 0x80) || (i == 4 && b < 0x10)
            return (result, offset)
        end
        i == 4 && break
    end
    error("Could not decode varint32.")
end

"""
    encode32!(buf::Vector{UInt8}, offset::Integer, value::UInt32)

Write `value` as a varint32 into `buf` at `offset`; returns the index just past it
(reference: src/varint.jl:46-69).
"""
function encode32!(buf::Vector{UInt8}, offset::Integer, value::UInt32)
    while value >= 0x80
        buf[offset] = (value | 0x80) % UInt8
        value >>= 7
        offset += 1
    end
    buf[offset] = value % UInt8
    return offset + 1
end
