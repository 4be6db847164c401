"""the `rules` switch -- the compressor reproduces Google snappy's bytes instead of Snappy.jl's.

CPU part: the oracle's rule sets.  rules = 2 is pinned against a REAL Google snappy (the codec bundled with pyarrow)
byte for byte; rules = 1 (libsnappy <= 1.1.7: another bucket function, half the table) has no binary to pin against
here and rests on the two independent restatements agreeing (snappy_oracle.c and oracle/py_rules.py), on the known
deltas against Snappy.jl (alice29: 88 034 vs 88 039 bytes) and on every decoder accepting it.
GPU part (-m gpu): the kernels under set_rules(1|2) against the oracle and against pyarrow."""
import os
import sys

import numpy as np
import pytest

from conftest import ALL_FILES, ROOT, dictionary_fuzz, edge_inputs, read_data

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pa = pytest.importorskip("pyarrow")
FILES = ALL_FILES + ["plrabn12.txt", "urls.10K", "random1.bin", "smallrandom1.bin"]


def google(raw):
    return pa.Codec("snappy").compress(bytes(raw), asbytes=True)


def small_inputs(seed, count):
    rng = np.random.default_rng(seed)
    for _ in range(count):
        n = int(rng.integers(0, 700))
        k = int(rng.choice([2, 3, 4, 8, 16, 256]))
        yield rng.integers(0, k, n, dtype=np.uint8).tobytes()


def boundary_inputs():
    rng = np.random.default_rng(77)
    words = [rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8).tobytes() for _ in range(200)]
    text = b" ".join(words[i] for i in rng.integers(0, 200, 60000))
    for n in (0, 1, 14, 15, 16, 17, 59, 60, 61, 62, 255, 256, 257, 65535, 65536, 65537, 65536 + 14, 65536 + 15,
              65536 + 16, 2 * 65536, 2 * 65536 + 300, 3 * 65536 - 1):
        yield text[:n]
    yield bytes(range(60))                       # the 60-byte literal (appendix B.4 (2))
    yield bytes(range(61))
    yield b"a" * 70 + bytes(range(100, 130))     # copy splitting (appendix B.4 (3))
    yield b"ab" * 40000 + bytes(range(256)) * 3


@pytest.mark.parametrize("name", FILES)
def test_oracle_rules2_is_google_snappy_on_files(oracle, name):
    raw = read_data(name)
    assert oracle.compress_rules(raw, 2) == google(raw)


def test_oracle_rules2_is_google_snappy_on_small_and_boundary_inputs(oracle):
    for raw in list(small_inputs(5, 1500)) + list(boundary_inputs()) + list(edge_inputs().values()):
        assert oracle.compress_rules(raw, 2) == google(raw), len(raw)
    for raw in dictionary_fuzz(3, 6):
        assert oracle.compress_rules(raw, 2) == google(raw)


def test_oracle_rules_agree_with_the_python_statement(oracle):
    import py_rules
    inputs = list(small_inputs(6, 400)) + [b for b in boundary_inputs() if len(b) <= 140000]
    inputs += [read_data("html"), read_data("sample-tweet.json")]
    for raw in inputs:
        for rules in (0, 1, 2):
            assert oracle.compress_rules(raw, rules) == py_rules.compress(raw, rules), (len(raw), rules)


def test_oracle_rules0_is_the_reference_restatement_and_rules1_deltas(oracle):
    for name in FILES:
        raw = read_data(name)
        assert oracle.compress_rules(raw, 0) == oracle.compress(raw)
        c1 = oracle.compress_rules(raw, 1)
        assert oracle.uncompress(c1) == raw                                   # the reference's decoder accepts it
        assert pa.Codec("snappy").decompress(c1, len(raw), asbytes=True) == raw  # and so does Google's
    # alice29 is 88 034 bytes with libsnappy 1.1.x, 88 039 with Snappy.jl
    assert len(oracle.compress_rules(read_data("alice29.txt"), 1)) == 88034
    assert len(oracle.compress(read_data("alice29.txt"))) == 88039
    # appendix B.4 (2): only Snappy.jl spends two header bytes on a 60-byte literal
    assert oracle.compress(bytes(range(60)))[:3] == bytes([0x3C, 0xF0, 0x3B])
    assert oracle.compress_rules(bytes(range(60)), 1)[:2] == bytes([0x3C, 0xEC])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture()
def rules_mode(snappy):
    yield snappy.set_rules
    snappy.set_rules(0)


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("rules", [1, 2])
def test_gpu_rules_files_and_edges(snappy, oracle, rules_mode, rules):
    rules_mode(rules)
    inputs = [read_data(n) for n in FILES] + list(boundary_inputs()) + list(edge_inputs().values())
    inputs += list(small_inputs(8, 60)) + list(dictionary_fuzz(4, 4))
    for raw in inputs:
        c = snappy.compress(raw)
        assert c == oracle.compress_rules(raw, rules), (len(raw), rules)
        if rules == 2:
            assert c == google(raw), len(raw)
        assert snappy.uncompress(c) == bytes(raw)
    rules_mode(0)
    raw = read_data("alice29.txt")
    assert snappy.compress(raw) == oracle.compress(raw)      # and back to the reference's bytes


@pytest.mark.gpu
@pytest.mark.parametrize("rules", [1, 2])
def test_gpu_rules_mix_device_shards_and_streamed(snappy, oracle, rules_mode, rules):
    from snappy_jl_b200 import device as dev, synth
    rules_mode(rules)
    raw = synth.mix(700, seed=12, tail=4321)                 # both kernels (shared and global tables) take part
    want = oracle.compress_rules(raw, rules)
    if rules == 2:
        assert want == google(raw)
    stream, index = dev.compress_device(to_dev(raw), want_index=True)
    assert stream.cpu().numpy().tobytes() == want
    assert np.array_equal(dev.uncompress_device(stream).cpu().numpy(), raw)
    assert np.array_equal(dev.uncompress_device(stream, index=index).cpu().numpy(), raw)   # the side index holds
    # shards: tables are per fragment, so shard outputs concatenate to the stream
    hdr = len(oracle.encode32(raw.size))
    cuts = [0, 100 * 65536, 333 * 65536, raw.size]
    parts = [dev.compress_shard_device(to_dev(raw[a:b]), raw.size)[0].cpu().numpy().tobytes()
             for a, b in zip(cuts[:-1], cuts[1:])]
    assert b"".join(parts) == want[hdr:]
    # streamed host-buffer path (> 128 MiB)
    big = synth.mix(2304, seed=13, tail=99)
    c = snappy.compress_np(big)
    assert c.tobytes() == oracle.compress_rules(big, rules)
    assert np.array_equal(snappy.uncompress_np(c), big)


@pytest.mark.gpu
@pytest.mark.parametrize("rules", [1, 2])
def test_gpu_rules_batched_pages(snappy, oracle, rules_mode, rules):
    from snappy_jl_b200 import device as dev, synth
    rules_mode(rules)
    count, page = 256, 4096
    pages = synth.pages(count, page, seed=6)
    rng = np.random.default_rng(2)
    sizes = np.full(count, page, dtype=np.int32)
    sizes[rng.integers(0, count, 40)] = rng.integers(0, page, 40)
    offs = np.arange(count, dtype=np.int64) * page
    out, oo, os_ = dev.compress_batched_device(to_dev(pages.reshape(-1)), to_dev(offs), to_dev(sizes))
    out_h, oo, os_ = out.cpu().numpy(), oo.cpu().numpy(), os_.cpu().numpy()
    for i in range(count):
        got = out_h[oo[i]: oo[i] + os_[i]].tobytes()
        assert got == oracle.compress_rules(pages[i, : sizes[i]], rules), i
        if rules == 2:
            assert got == google(pages[i, : sizes[i]].tobytes()), i
    # one multi-fragment page with a 64 KiB table under rules 2
    one = synth.mix(3, seed=14, tail=1000)
    out, oo, os_ = dev.compress_batched_device(to_dev(one), to_dev(np.zeros(1, dtype=np.int64)),
                                               to_dev(np.array([one.size], dtype=np.int32)))
    assert out.cpu().numpy()[: int(os_.cpu()[0])].tobytes() == oracle.compress_rules(one, rules)


def test_set_rules_validates_and_needs_no_device(snappy):
    """the switch is host state: it can be set before any device exists, and rejects unknown rule sets"""
    snappy.set_rules(2)
    snappy.set_rules(0)
    with pytest.raises(ValueError):
        snappy.set_rules(3)
