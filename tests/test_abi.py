"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/snappy_b200.h declares,
its host-only helpers agree with the reference's tests, and -- with no GPU -- every compute entry
point refuses loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from test_oracle import FML_KATS


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "snappy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(snappy_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(snappy):
    syms = declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(snappy._abi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libsnappy_b200.so does not export %s" % s
    # and the Python binding table covers exactly the header
    assert sorted(snappy._abi.SIGNATURES) == syms


def test_library_does_not_link_oracle():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "snappy.jl_b200", "libsnappy_b200.so")],
                         capture_output=True, text=True).stdout
    assert "oracle" not in out
    nm = subprocess.run(["nm", "-D", os.path.join(ROOT, "snappy.jl_b200", "libsnappy_b200.so")],
                        capture_output=True, text=True).stdout
    assert "sjo_" not in nm


def test_maxlength_compressed(snappy):
    for n in (0, 1, 5, 6, 65536, 10 ** 6, 2 ** 30):
        assert snappy.maxlength_compressed(n) == 32 + n + n // 6  # src/Snappy.jl:80-82


def test_varint_host(snappy):
    for i in range(31):  # test/runtests.jl:157-163
        enc = snappy.encode32(1 << i)
        assert snappy.parse32(enc, 0) == (1 << i, len(enc))
    assert snappy.length_uncompressed(bytes([0xA0, 0x8D, 0x06, 0x00])) == (100000, 3)
    for bad in (bytes([0xF0]), bytes([0x80, 0x80, 0x80, 0x80, 0x80, 0x0A]),
                bytes([0xFB, 0xFF, 0xFF, 0xFF, 0x7F]), b""):
        with pytest.raises(snappy.SnappyError) as e:  # test/runtests.jl:101-111
            snappy.parse32(bad, 0)
        assert str(e.value) == "Could not decode varint32."


@pytest.mark.parametrize("a,b,limit,want", FML_KATS)
def test_find_match_length_host(snappy, a, b, limit, want):  # test/runtests.jl:176-267
    c = (a + b).encode("latin-1")
    assert snappy.find_match_length(c, 0, len(a), len(a) + limit) == want


def test_status_strings(snappy):
    msgs = {1: "Input too large.", 2: "Invalid input.", 3: "Invalid input: corrupt copy offset",
            4: "Invalid input: corrupt copy length", 5: "Invalid input: corrupt literal",
            6: "Could not decode varint32."}
    for code, msg in msgs.items():
        assert snappy._abi.status_string(code) == msg


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(snappy):
    with pytest.raises(snappy.SnappyError) as e:
        snappy.compress(b"no gpu here, so this must fail")
    assert e.value.status == snappy._abi.NO_DEVICE
    with pytest.raises(snappy.SnappyError) as e:
        snappy.uncompress(b"\x03\x08abc")
    assert e.value.status == snappy._abi.NO_DEVICE
    # the varint is still checked on the host first, as in src/Snappy.jl:47
    with pytest.raises(snappy.SnappyError) as e:
        snappy.uncompress(bytes([0xF0]))
    assert e.value.status == snappy._abi.BAD_VARINT


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "snappy.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "snappy_oracle" not in text and "liboracle" not in text, f


def test_index_sidecar_roundtrip_and_rejects(snappy):
    """side-index sidecar: host-only pack / unpack"""
    rng = np.random.default_rng(5)
    for nfrag, total in ((1, 1), (1, 65536), (3, 3 * 65536 - 17), (1000, 1000 * 65536)):
        sizes = rng.integers(1, 76490, size=nfrag)
        index = np.concatenate([[5], 5 + np.cumsum(sizes)]).astype(np.uint64)
        side = snappy.pack_index(index, total)
        back, ulen, slen = snappy.unpack_index(side)
        assert np.array_equal(back, index) and ulen == total and slen == int(index[-1])
        assert len(side) < 40 + 3 * (nfrag + 1)          # ~2-3 bytes per fragment
        for cut in (0, 7, len(side) - 1):
            with pytest.raises(snappy.SnappyError):
                snappy.unpack_index(side[:cut])
        bad = bytearray(side)
        bad[len(bad) // 2] ^= 0x40
        with pytest.raises(snappy.SnappyError):
            snappy.unpack_index(bytes(bad))
    with pytest.raises(snappy.SnappyError):                 # nfrag does not match the length
        snappy.unpack_index(snappy.pack_index(np.array([1, 9, 20], dtype=np.uint64), 65536))
    with pytest.raises(snappy.SnappyError):                 # decreasing offsets
        snappy.pack_index(np.array([5, 4], dtype=np.uint64), 10)


def test_index_sidecar_fuzz_with_valid_checksum(snappy):
    """the sidecar comes from outside: mutated blobs whose checksum has been fixed up (so the parser goes all the
    way in) either unpack to exactly what a re-pack reproduces or are rejected -- never a crash or a wrong length"""
    def fnv1a(b):
        h = 2166136261
        for x in b:
            h = ((h ^ x) * 16777619) & 0xFFFFFFFF
        return h

    rng = np.random.default_rng(17)
    sizes = rng.integers(1, 76490, size=40)
    index = np.concatenate([[5], 5 + np.cumsum(sizes)]).astype(np.uint64)
    good = snappy.pack_index(index, 40 * 65536 - 3)
    accepted = rejected = 0
    for _ in range(3000):
        b = bytearray(good[:-4])
        for _ in range(int(rng.integers(1, 4))):
            kind = int(rng.integers(0, 4))
            pos = int(rng.integers(8, len(b)))
            if kind == 0:
                b[pos] = int(rng.integers(0, 256))
            elif kind == 1:
                b[pos] ^= 1 << int(rng.integers(0, 8))
            elif kind == 2 and len(b) > 40:
                del b[pos]
            else:
                b.insert(pos, int(rng.integers(0, 256)))
        blob = bytes(b) + fnv1a(b).to_bytes(4, "little")
        try:
            back, ulen, slen = snappy.unpack_index(blob)
        except snappy.SnappyError:
            rejected += 1
            continue
        accepted += 1
        assert len(back) == (ulen + 65535) // 65536 + 1 and int(back[-1]) == slen
        assert np.all(np.diff(back.astype(np.int64)) >= 0) or int(back[-1]) < 2 ** 63
        assert snappy.unpack_index(snappy.pack_index(back, ulen))[1:] == (ulen, slen)
    assert rejected > 1000 and accepted + rejected == 3000


def test_input_too_large_at_2_pow_32(snappy):
    """src/Snappy.jl:21: `length(input) > 2^32 - 1` is "Input too large." -- decided from the length alone, before any
    byte is read or any device is touched (so the check needs neither 4 GiB of memory nor a GPU)."""
    import ctypes
    lib = snappy._abi.lib()
    buf = (ctypes.c_uint8 * 16)()
    out_len = ctypes.c_size_t(1 << 40)
    for n in (1 << 32, (1 << 32) + 1, 1 << 40):
        assert lib.snappy_b200_compress(buf, n, buf, ctypes.byref(out_len)) == snappy._abi.INPUT_TOO_LARGE
        assert lib.snappy_b200_compress_device(buf, n, buf, 1 << 41, ctypes.byref(out_len), None, None) == \
            snappy._abi.INPUT_TOO_LARGE
        assert lib.snappy_b200_compress_shard_device(buf, 65536, n, buf, 1 << 20, ctypes.byref(out_len), None,
                                                     None) == snappy._abi.INPUT_TOO_LARGE
    assert snappy._abi.status_string(snappy._abi.INPUT_TOO_LARGE) == "Input too large."
    # the largest legal length passes the length check (and then stops at the next one: the output capacity)
    small = ctypes.c_size_t(16)
    assert lib.snappy_b200_compress(buf, (1 << 32) - 1, buf, ctypes.byref(small)) == snappy._abi.BUFFER_TOO_SMALL


def test_get_option_and_product_build_has_no_experiments(snappy):
    lib = snappy._abi.lib()
    assert lib.snappy_b200_get_option(b"experiments") == 0  # the shipped library holds the default kernels only
    assert lib.snappy_b200_get_option(b"rules") == 0 and lib.snappy_b200_get_option(b"lpt") == 1
    assert lib.snappy_b200_get_option(b"no_such_option") == -1
    lib.snappy_b200_set_option(b"wide", 4)  # experiment selectors are inert in the product build
    assert lib.snappy_b200_get_option(b"wide") == 0
