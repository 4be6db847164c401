"""ctypes loader for oracle/liboracle.so (the C restatement of Snappy.jl).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

OK, INPUT_TOO_LARGE, INVALID_INPUT, CORRUPT_COPY_OFFSET, CORRUPT_COPY_LENGTH, CORRUPT_LITERAL, \
    BAD_VARINT, BUFFER_TOO_SMALL = range(8)

MESSAGES = {
    INPUT_TOO_LARGE: "Input too large.",
    INVALID_INPUT: "Invalid input.",
    CORRUPT_COPY_OFFSET: "Invalid input: corrupt copy offset",
    CORRUPT_COPY_LENGTH: "Invalid input: corrupt copy length",
    CORRUPT_LITERAL: "Invalid input: corrupt literal",
    BAD_VARINT: "Could not decode varint32.",
    BUFFER_TOO_SMALL: "output buffer too small",
}


class OracleError(Exception):
    def __init__(self, code, err_op=None):
        super().__init__(MESSAGES.get(code, "status %d" % code))
        self.code = code
        self.err_op = err_op


def build(force=False):
    src = os.path.join(_HERE, "snappy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p, szp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)
        L.sjo_maxlength_compressed.restype = ctypes.c_size_t
        L.sjo_maxlength_compressed.argtypes = [ctypes.c_size_t]
        L.sjo_encode32.restype = ctypes.c_int
        L.sjo_encode32.argtypes = [u8p, ctypes.c_uint32]
        L.sjo_parse32.restype = ctypes.c_int
        L.sjo_parse32.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t,
                                  ctypes.POINTER(ctypes.c_uint32), szp]
        L.sjo_find_match_length.restype = ctypes.c_size_t
        L.sjo_find_match_length.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t]
        L.sjo_hashtable_entries.restype = ctypes.c_uint32
        L.sjo_hashtable_entries.argtypes = [ctypes.c_uint64]
        L.sjo_compress_fragments.restype = ctypes.c_size_t
        L.sjo_compress_fragments.argtypes = [u8p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_size_t,
                                             u8p, ctypes.c_void_p]
        L.sjo_compress_fragment.restype = ctypes.c_size_t
        L.sjo_compress_fragment.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.c_void_p, ctypes.c_uint32]
        L.sjo_compress.restype = ctypes.c_int
        L.sjo_compress.argtypes = [u8p, ctypes.c_size_t, u8p, szp]
        L.sjo_compress_rules.restype = ctypes.c_int
        L.sjo_compress_rules.argtypes = [u8p, ctypes.c_size_t, u8p, szp, ctypes.c_int]
        L.sjo_uncompressed_length.restype = ctypes.c_int
        L.sjo_uncompressed_length.argtypes = [u8p, ctypes.c_size_t, szp]
        L.sjo_char_table_entry.restype = ctypes.c_uint16
        L.sjo_char_table_entry.argtypes = [ctypes.c_uint32]
        L.sjo_uncompress_ex.restype = ctypes.c_int
        L.sjo_uncompress_ex.argtypes = [u8p, ctypes.c_size_t, u8p, szp, szp]
        _lib = L
    return _lib


def _as_u8(data):
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
    else:
        a = np.frombuffer(bytes(data), dtype=np.uint8)
    return a


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a.size else ctypes.c_void_p(0)


def maxlength_compressed(n):
    return lib().sjo_maxlength_compressed(n)


def compress_np(data):
    """np.uint8 array in, np.uint8 array (compressed stream) out."""
    a = _as_u8(data)
    out = np.empty(maxlength_compressed(a.size), dtype=np.uint8)
    n = ctypes.c_size_t(out.size)
    rc = lib().sjo_compress(_ptr(a), a.size, _ptr(out), ctypes.byref(n))
    if rc != OK:
        raise OracleError(rc)
    return out[: n.value]


def compress(data):
    return compress_np(data).tobytes()


def compress_rules(data, rules):
    """Google snappy's rules instead of the reference's (1 = libsnappy <= 1.1.7, 2 = snappy >= 1.1.9)."""
    a = _as_u8(data)
    out = np.empty(maxlength_compressed(a.size), dtype=np.uint8)
    n = ctypes.c_size_t(out.size)
    rc = lib().sjo_compress_rules(_ptr(a), a.size, _ptr(out), ctypes.byref(n), int(rules))
    if rc != OK:
        raise OracleError(rc)
    return out[: n.value].tobytes()


def compress_fragments(data, total_len, first_frag, nfrag):
    """Compress fragments [first_frag, first_frag+nfrag) of the stream `data` (whole stream)."""
    a = _as_u8(data)
    out = np.empty(nfrag * (65536 + 65536 // 6 + 32) + 32, dtype=np.uint8)
    sizes = np.zeros(max(nfrag, 1), dtype=np.uint32)
    n = lib().sjo_compress_fragments(_ptr(a), total_len, first_frag, nfrag, _ptr(out),
                                     ctypes.c_void_p(sizes.ctypes.data))
    return out[:n], sizes[:nfrag]


def compress_one_fragment(frag, total_len):
    """One <= 64 KiB fragment of a stream whose TOTAL length is total_len (the table is sized from the total,
    Snappy.jl:27, and reset per fragment, :30): the fragment's element bytes, no header."""
    a = _as_u8(frag)
    assert a.size <= 65536
    entries = lib().sjo_hashtable_entries(int(total_len))
    table = np.full(16384, 0xFFFF, dtype=np.uint16)
    out = np.empty(65536 + 65536 // 6 + 64, dtype=np.uint8)
    n = lib().sjo_compress_fragment(_ptr(a), a.size, _ptr(out), ctypes.c_void_p(table.ctypes.data), entries)
    return out[:n]


def uncompressed_length(data):
    a = _as_u8(data)
    r = ctypes.c_size_t(0)
    rc = lib().sjo_uncompressed_length(_ptr(a), a.size, ctypes.byref(r))
    if rc != OK:
        raise OracleError(rc)
    return r.value


def uncompress_np(data):
    a = _as_u8(data)
    n = uncompressed_length(a)
    out = np.zeros(n, dtype=np.uint8)
    cap = ctypes.c_size_t(n)
    err_op = ctypes.c_size_t(0)
    rc = lib().sjo_uncompress_ex(_ptr(a), a.size, _ptr(out), ctypes.byref(cap), ctypes.byref(err_op))
    if rc != OK:
        raise OracleError(rc, err_op.value)
    return out


def uncompress(data):
    return uncompress_np(data).tobytes()


def status_of_uncompress(data):
    """Return the oracle's status code for uncompress(data) without raising."""
    try:
        uncompress_np(data)
        return OK
    except OracleError as e:
        return e.code


def parse32(buf, offset=0):
    a = _as_u8(buf)
    v = ctypes.c_uint32(0)
    nxt = ctypes.c_size_t(0)
    rc = lib().sjo_parse32(_ptr(a), a.size, offset, ctypes.byref(v), ctypes.byref(nxt))
    if rc != OK:
        raise OracleError(rc)
    return v.value, nxt.value


def encode32(value):
    buf = np.zeros(5, dtype=np.uint8)
    k = lib().sjo_encode32(_ptr(buf), value)
    return buf[:k].tobytes()


def find_match_length(a, i1, i2, limit):
    """0-based, `limit` exclusive."""
    arr = _as_u8(a)
    return lib().sjo_find_match_length(_ptr(arr), i1, i2, limit)


def char_table():
    """CHAR_TABLE (internal.jl:47-80) as the C oracle regenerates it: 256 ints."""
    return [int(lib().sjo_char_table_entry(c)) for c in range(256)]
