"""GPU suite: the parity tests proper.  Everything goes through the C ABI of libsnappy_b200.so
(host-buffer and device-resident entry points) and is compared with the CPU oracle on the same
inputs -- bit-exact, both directions."""
import hashlib
import os

import numpy as np
import pytest

from conftest import (ALL_FILES, ROOT, corrupt_streams, dictionary_fuzz, edge_inputs, read_data)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(snappy):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from snappy_jl_b200 import device
    return device


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", ALL_FILES)
def test_files_bit_exact(snappy, oracle, golden, name):  # test/runtests.jl:6-33 + byte parity
    raw = read_data(name)
    c = snappy.compress(raw)
    assert c == oracle.compress(raw)
    g = golden["files"][name]
    assert len(c) == g["comp_len"] and hashlib.sha256(c).hexdigest() == g["comp_sha256"]
    assert snappy.uncompress(c) == raw            # round trip on the GPU
    assert oracle.uncompress(c) == raw            # GPU stream decodes on the reference path
    assert c != raw


def test_edges_bit_exact(snappy, oracle, golden):  # test/runtests.jl:125-137 and quirk vectors
    for name, raw in edge_inputs().items():
        c = snappy.compress(raw)
        assert c == oracle.compress(raw), name
        assert hashlib.sha256(c).hexdigest() == golden["edges"][name]["comp_sha256"], name
        assert snappy.uncompress(c) == raw, name


def test_string_api(snappy, oracle):  # src/Snappy.jl:38
    s = "héllo wörld " * 50
    assert snappy.compress(s) == oracle.compress(s.encode("utf-8"))
    assert snappy.uncompress(snappy.compress(s)).decode("utf-8") == s


def test_dictionary_fuzz(snappy, oracle):  # test/runtests.jl:35-60
    for raw in dictionary_fuzz(21, 24):
        c = snappy.compress(raw)
        assert c == oracle.compress(raw)
        assert snappy.uncompress(c) == raw


def test_size_sweep(snappy, oracle):
    rng = np.random.default_rng(17)
    sizes = list(range(0, 40)) + [59, 60, 61, 62, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097,
                                  16383, 16384, 16385, 65520, 65521, 65535, 65536, 65537, 65550, 65551,
                                  65552, 131071, 131072, 131073, 200000]
    for size in sizes:
        for alpha in (2, 3, 256):
            raw = rng.integers(0, alpha, size, dtype=np.uint8).tobytes()
            c = snappy.compress(raw)
            assert c == oracle.compress(raw), (size, alpha)
            assert snappy.uncompress(c) == raw, (size, alpha)


def test_max_blowup(snappy, oracle):  # test/runtests.jl:148-154
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 2 ** 32, 20000, dtype=np.uint32).view(np.uint8)
    raw = np.concatenate([raw, raw[::-1]]).tobytes()
    c = snappy.compress(raw)
    assert c == oracle.compress(raw)
    assert snappy.uncompress(c) == raw


def test_must_throw_same_status_as_reference(snappy, oracle):  # test/runtests.jl:62-123
    for name, stream in corrupt_streams(oracle):
        want = oracle.status_of_uncompress(stream)
        assert want != oracle.OK
        with pytest.raises(snappy.SnappyError) as e:
            snappy.uncompress(stream)
        assert e.value.status == want, name
        assert str(e.value) == oracle.MESSAGES[want], name


def test_decoder_quirks(snappy):
    assert snappy.uncompress(bytes([0x00, 0x00])) == b""
    assert snappy.uncompress(bytes([0x01, 0x00, 0x41, 0x77])) == b"A"


def test_mutation_fuzz_status_parity(snappy, oracle):
    """flip bytes of valid streams: GPU decoder's status / output must equal the oracle's"""
    rng = np.random.default_rng(99)
    bases = [oracle.compress(read_data("sample-tweet.json")), oracle.compress(read_data("html")[:30000]),
             oracle.compress(b"abcd" * 3000 + bytes(range(256)) * 4)]
    for base in bases:
        for _ in range(60):
            s = bytearray(base)
            for _k in range(rng.integers(1, 4)):
                s[rng.integers(1, len(s))] = rng.integers(0, 256)
            if rng.integers(0, 4) == 0:
                s = s[: rng.integers(2, len(s))]
            s = bytes(s)
            want = oracle.status_of_uncompress(s)
            if want == oracle.OK:
                assert snappy.uncompress(s) == oracle.uncompress(s)
            else:
                with pytest.raises(snappy.SnappyError) as e:
                    snappy.uncompress(s)
                assert e.value.status == want


def test_foreign_streams(snappy, oracle):
    # 32 KiB-block encoder output shipped with the reference (unreferenced fixture)
    assert snappy.uncompress(read_data("alice29.snappy")) == read_data("alice29.txt")
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    for name in ("html_x_4", "urls.10K", "kppkn.gtb"):
        raw = read_data(name)
        theirs = codec.compress(raw).to_pybytes()
        assert snappy.uncompress(theirs) == raw
        ours = snappy.compress(raw)
        assert codec.decompress(ours, decompressed_size=len(raw)).to_pybytes() == raw


def test_hand_made_streams(snappy, oracle):
    # 4-byte-offset copy (never emitted by Snappy.jl, decoder accepts; Appendix A) and overlapping copies
    lit = bytes([0x0C]) + b"abcd"                       # literal len 4
    copy4 = bytes([(8 - 1) << 2 | 3]) + (4).to_bytes(4, "little")   # copy len 8 offset 4 (4-byte form)
    rle = bytes([(64 - 1) << 2 | 2]) + (1).to_bytes(2, "little")     # copy len 64 offset 1
    body = lit + copy4 + rle
    s = oracle.encode32(4 + 8 + 64) + body
    want = oracle.uncompress(s)
    assert want == b"abcd" * 3 + b"d" * 64
    assert snappy.uncompress(s) == want


def test_device_api_roundtrip_with_index(dev, oracle):
    import torch
    from snappy_jl_b200 import synth
    raw = synth.mix(64, seed=5, tail=12345)
    d = to_dev(raw)
    stream, index = dev.compress_device(d, want_index=True)
    want = oracle.compress_np(raw)
    assert np.array_equal(stream.cpu().numpy(), want)
    idx = index.cpu().numpy()
    nfrag = dev.nfragments(raw.size)
    assert idx.shape[0] == nfrag + 1 and idx[-1] == want.size and idx[0] == len(oracle.encode32(raw.size))
    _, sizes = oracle.compress_fragments(raw, raw.size, 0, nfrag)
    assert np.array_equal(np.diff(idx), sizes.astype(np.int64))
    back = dev.uncompress_device(stream, index=index, claimed=raw.size)
    assert torch.equal(back, d)
    back2 = dev.uncompress_device(stream)  # no side index: segmented speculative parse
    assert torch.equal(back2, d)
    # a wrong index must not change the result
    bad = index.clone()
    bad[3] += 1
    back3 = dev.uncompress_device(stream, index=bad, claimed=raw.size)
    assert torch.equal(back3, d)


def test_shard_api(dev, oracle):
    from snappy_jl_b200 import synth
    import torch
    raw = synth.mix(40, seed=9, tail=777)
    total = raw.size
    whole = oracle.compress_np(raw)
    hdr = oracle.encode32(total)
    cuts = [0, 7 * 65536, 19 * 65536, total]
    parts, all_sizes = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        out, sizes = dev.compress_shard_device(to_dev(raw[a:b]), total, want_sizes=True)
        parts.append(out.cpu().numpy())
        all_sizes.append(sizes.cpu().numpy())
    assert hdr + b"".join(p.tobytes() for p in parts) == whole.tobytes()
    # decode each shard from its own element bytes
    for (a, b), p, s in zip(zip(cuts[:-1], cuts[1:]), parts, all_sizes):
        offs = np.concatenate([[0], np.cumsum(s.astype(np.int64))])
        got = dev.uncompress_shard_device(to_dev(p), to_dev(offs), b - a)
        assert np.array_equal(got.cpu().numpy(), raw[a:b])


def test_small_total_uses_small_table(snappy, oracle):
    # table size comes from the TOTAL length (src/Snappy.jl:27): 256..16384 entries
    rng = np.random.default_rng(2)
    for n in (100, 300, 600, 1500, 3000, 6000, 12000, 20000):
        raw = rng.integers(0, 3, n, dtype=np.uint8).tobytes()
        assert snappy.compress(raw) == oracle.compress(raw), n


def test_batched_pages(dev, oracle):
    import torch
    from snappy_jl_b200 import synth
    count, page = 512, 4096
    pages = synth.pages(count, page, seed=4)
    # ragged sizes as well: shrink some pages
    rng = np.random.default_rng(1)
    sizes = np.full(count, page, dtype=np.int32)
    sizes[rng.integers(0, count, 64)] = rng.integers(0, page, 64)
    offs = (np.arange(count, dtype=np.int64) * page)
    d = to_dev(pages.reshape(-1))
    out, out_offs, out_sizes = dev.compress_batched_device(d, to_dev(offs), to_dev(sizes))
    out_h, oo, os_ = out.cpu().numpy(), out_offs.cpu().numpy(), out_sizes.cpu().numpy()
    for i in range(count):
        want = oracle.compress_np(pages[i, : sizes[i]])
        got = out_h[oo[i]: oo[i] + os_[i]]
        assert np.array_equal(got, want), i
    back = torch.zeros(count * page, dtype=torch.uint8, device="cuda")
    caps = to_dev(np.full(count, page, dtype=np.int32))
    got_sizes, statuses = dev.uncompress_batched_device(out, out_offs, out_sizes, back, to_dev(offs), caps)
    assert int(statuses.abs().sum().item()) == 0
    assert np.array_equal(got_sizes.cpu().numpy(), sizes)
    bh = back.cpu().numpy().reshape(count, page)
    for i in range(count):
        assert np.array_equal(bh[i, : sizes[i]], pages[i, : sizes[i]]), i


def test_mix_64mib_bit_exact(dev, oracle):
    import torch
    from snappy_jl_b200 import synth
    raw = synth.mix(1024, seed=2026)
    d = to_dev(raw)
    stream, index = dev.compress_device(d, want_index=True)
    want = oracle.compress_np(raw)
    assert stream.numel() == want.size
    assert np.array_equal(stream.cpu().numpy(), want)
    assert torch.equal(dev.uncompress_device(stream, index=index, claimed=raw.size), d)
    # config 3: the reference-produced stream, no side index
    assert torch.equal(dev.uncompress_device(to_dev(want)), d)


def test_full_size_1gib_properties(dev, oracle):
    """BASELINE config 2 at full size: round trip, and bit-exactness of sampled fragments through
    the side index (the whole-stream oracle compare runs at 64 MiB above)."""
    import torch
    from snappy_jl_b200 import synth
    nfrag = 16384
    raw = synth.mix(nfrag, seed=2026)
    d = to_dev(raw)
    stream, index = dev.compress_device(d, want_index=True)
    idx = index.cpu().numpy()
    assert idx[0] == 5 and idx[-1] == stream.numel()
    assert bytes(stream[:5].cpu().numpy()) == bytes([0x80, 0x80, 0x80, 0x80, 0x04])
    sh = stream.cpu().numpy()
    rng = np.random.default_rng(0)
    for f in np.concatenate([[0, nfrag - 1], rng.integers(0, nfrag, 200)]):
        want, _ = oracle.compress_fragments(raw, raw.size, int(f), 1)
        assert np.array_equal(sh[idx[f]: idx[f + 1]], want), f
    back = dev.uncompress_device(stream, index=index, claimed=raw.size)
    assert torch.equal(back, d)
    del back
    back = dev.uncompress_device(stream)  # arbitrary-stream path at full size
    assert torch.equal(back, d)


def test_host_api_pipelined_paths(snappy, oracle):
    """> 256 MiB through the host-buffer C ABI takes the chunk-pipelined H2D / kernel / D2H path
    (SURVEY.md 8(f)1); the bytes must still be the oracle's, both ways."""
    from snappy_jl_b200 import synth
    raw = synth.mix(4096 + 1000, seed=77, tail=4242)
    got = snappy.compress_np(raw)
    want = oracle.compress_np(raw)
    assert got.size == want.size and np.array_equal(got, want)
    assert np.array_equal(snappy.uncompress_np(want), raw)


def test_batched_shard_api(dev, oracle):
    """several shards of different streams (different totals, ragged tails, a tiny-table stream) in
    one kernel pass == the same shards one by one == the oracle"""
    from snappy_jl_b200 import synth
    import torch
    streams = [synth.mix(9, seed=31, tail=555), synth.mix(3, seed=32), np.frombuffer(read_data("urls.10K"), dtype=np.uint8),
               np.frombuffer(b"tiny stream with a small table " * 20, dtype=np.uint8)]
    cuts = [(0, 4 * 65536), (4 * 65536, streams[0].size), (0, streams[1].size), (65536, streams[2].size), (0, streams[3].size)]
    which = [0, 0, 1, 2, 3]
    shards = [to_dev(streams[w][a:b]) for w, (a, b) in zip(which, cuts)]
    totals = [streams[w].size for w in which]
    res = dev.compress_shards_device(shards, totals)
    datas, fos, lens = [], [], []
    for (seg, sizes), w, (a, b) in zip(res, which, cuts):
        f0 = a // 65536
        nf = (b - a + 65535) // 65536
        want, wsizes = oracle.compress_fragments(streams[w], streams[w].size, f0, nf)
        assert np.array_equal(seg.cpu().numpy(), want)
        assert np.array_equal(sizes.cpu().numpy().astype(np.uint32), wsizes)
        datas.append(seg)
        fos.append(to_dev(np.concatenate([[0], np.cumsum(wsizes.astype(np.int64))])))
        lens.append(b - a)
    outs = dev.uncompress_shards_device(datas, fos, lens)
    for o, w, (a, b) in zip(outs, which, cuts):
        assert np.array_equal(o.cpu().numpy(), streams[w][a:b])


VARIANT_DEFAULTS = {"window": 1, "wide": 0, "l2_chains": 14, "smem_chains": 6, "slowcont": 0, "lpt": 1, "mixed": 1}


def _variant_input():
    from snappy_jl_b200 import synth
    return np.concatenate([synth.mix(96, seed=5, tail=777),
                           np.frombuffer(read_data("alice29.txt") + read_data("html_x_4") + read_data("urls.10K"),
                                         dtype=np.uint8)])


@pytest.mark.parametrize("options", [
    {"l2_chains": 0},       # window kernel, shared-memory tables only
    {"smem_chains": 0},     # window kernel, global tables only
    {"lpt": 0},             # fragments in stream order instead of expensive-first (schedule.cuh)
    {"mixed": 0},           # the two table placements as two concurrent kernels (round 1) instead of one CTA
    {"mixed": 2},           # one CTA, shared-table warps on the low warp numbers
])
def test_compress_kernel_variants_bit_exact(dev, oracle, options):
    """every table placement / fragment order of the shipped compress kernel produces the oracle's bytes (the default
    is both table placements side by side, fragments handed out expensive-first)"""
    raw = _variant_input()
    want = oracle.compress_np(raw)
    try:
        for k, v in options.items():
            dev.set_option(k, v)
        stream, _ = dev.compress_device(to_dev(raw), want_index=False)
        got = stream.cpu().numpy()
    finally:
        for k, v in VARIANT_DEFAULTS.items():
            dev.set_option(k, v)
    assert got.size == want.size and np.array_equal(got, want)


def test_lpt_order_is_a_permutation_and_large_input_bit_exact(dev, oracle):
    """enough fragments for the expensive-first order to engage (>= 2 per SM): bytes and side index unchanged"""
    from snappy_jl_b200 import synth
    raw = synth.mix(700, seed=91, tail=12345)
    want = oracle.compress_np(raw)
    stream, index = dev.compress_device(to_dev(raw), want_index=True)
    assert np.array_equal(stream.cpu().numpy(), want)
    idx = index.cpu().numpy()
    assert idx[-1] == want.size and np.all(np.diff(idx) > 0)
    try:
        dev.set_option("lpt", 0)
        stream2, index2 = dev.compress_device(to_dev(raw), want_index=True)
    finally:
        dev.set_option("lpt", 1)
    assert np.array_equal(stream2.cpu().numpy(), want) and np.array_equal(index2.cpu().numpy(), idx)


EXPERIMENT_SCRIPT = r"""
import sys, json
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/oracle"); sys.path.insert(0, %(root)r + "/tests")
import numpy as np, torch, pyoracle
import snappy_jl_b200 as S
from snappy_jl_b200 import device as dev
from test_gpu_parity import _variant_input
assert S._abi.lib().snappy_b200_get_option(b"experiments") == 1
raw = _variant_input()
want = pyoracle.compress_np(raw)
d = torch.from_numpy(raw).cuda()
bad = []
for options in ({"window": 0}, {"wide": 4}, {"wide": 2}, {"slowcont": 1}, {"compress_variant": 1}, {"compress_variant": 2},
                {"two": 1}, {"two": 2}, {"two": 3}, {"mixed": 0}, {"mixed": 0, "l2_first": 1},
                {"pipe": 1}, {"pipe": 2}, {"pipe": 3}, {"unified": 1}, {"unified": 1, "pipe": 3}, {"pipe": 3, "mixed": 0}):
    for k, v in options.items():
        dev.set_option(k, v)
    got = dev.compress_device(d)[0].cpu().numpy()
    for k in options:
        dev.set_option(k, {"window": 1, "mixed": 1}.get(k, 0))
    if got.size != want.size or not np.array_equal(got, want):
        bad.append(options)
print("EXPERIMENTS_BAD", json.dumps(bad))
"""


def test_experimental_kernels_bit_exact():
    """The designs that were measured and lost (step-wise chain kernel, 2/4 warps per fragment, slowcont, the two
    TMA-staged first versions) live in libsnappy_b200_exp.so (make exp, -DSB200_EXPERIMENTS), not in the product
    library; they stay byte-identical to the oracle."""
    import subprocess
    import sys
    exp = os.path.join(ROOT, "snappy.jl_b200", "libsnappy_b200_exp.so")
    if not os.path.exists(exp):
        pytest.skip("libsnappy_b200_exp.so not built (make -C snappy.jl_b200/csrc exp)")
    p = subprocess.run([sys.executable, "-c", EXPERIMENT_SCRIPT % {"root": ROOT}], capture_output=True, text=True,
                       env=dict(os.environ, SNAPPY_B200_LIB=exp), timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "EXPERIMENTS_BAD []" in p.stdout, p.stdout


def test_streamed_uncompress_foreign_and_corrupt(snappy, oracle):
    """> 64 MiB through the host-buffer C ABI takes the streamed uncompress (the stream is parsed in
    segments while it uploads).  A foreign encoder's stream must decode, and a stream corrupted inside a
    late segment must end with the oracle's status (the whole-stream paths arbitrate)."""
    from snappy_jl_b200 import synth
    raw = synth.mix(1400, seed=123, tail=31337)          # ~92 MB, compresses to ~41 MB ... make it bigger
    raw = np.concatenate([raw, synth.source_like(96 << 20, seed=9)])   # > 64 MiB of compressed bytes
    ours = oracle.compress_np(raw)
    assert ours.size > (64 << 20)
    assert np.array_equal(snappy.uncompress_np(ours), raw)
    pa = pytest.importorskip("pyarrow")
    theirs = np.frombuffer(pa.Codec("snappy").compress(raw.tobytes()).to_pybytes(), dtype=np.uint8)
    if theirs.size > (64 << 20):
        assert np.array_equal(snappy.uncompress_np(theirs), raw)
    # corrupt one tag deep inside the stream (segment 6 of 8): same verdict as the reference
    bad = ours.copy()
    pos = int(bad.size * 0.7)
    bad[pos: pos + 3] = (0xFF, 0xFF, 0xFF)               # 4-byte-offset copy tag with a huge offset
    want = oracle.status_of_uncompress(bad.tobytes())
    if want == oracle.OK:
        assert np.array_equal(snappy.uncompress_np(bad), oracle.uncompress_np(bad))
    else:
        with pytest.raises(snappy.SnappyError) as e:
            snappy.uncompress_np(bad)
        assert e.value.status == want


@pytest.mark.parametrize("shift", [1, 3, 8, 13])
def test_unaligned_device_pointers(dev, oracle, shift):
    """device buffers that start at odd addresses (the ring staging and the far-candidate loads of
    the window kernel take their byte-wise forms) and odd-length tails"""
    import torch
    from snappy_jl_b200 import synth
    raw = synth.mix(24, seed=31, tail=12345)
    base = torch.zeros(raw.size + 64, dtype=torch.uint8, device="cuda")
    view = base[shift: shift + raw.size]
    view.copy_(to_dev(raw))
    want = oracle.compress_np(raw)
    stream, index = dev.compress_device(view, want_index=True)
    assert np.array_equal(stream.cpu().numpy(), want)
    out = torch.zeros(raw.size + 64, dtype=torch.uint8, device="cuda")
    sview = torch.zeros(want.size + 64, dtype=torch.uint8, device="cuda")[shift: shift + want.size]
    sview.copy_(stream)
    back = dev.uncompress_device(sview, out=out[shift: shift + raw.size], index=index, claimed=raw.size)
    assert np.array_equal(back.cpu().numpy(), raw)


def test_index_sidecar_drives_the_indexed_decoder(dev, snappy, oracle):
    """compress on the device, ship stream + sidecar, decode with the unpacked index; a sidecar of
    ANOTHER stream must not change the result (every fragment is validated)"""
    import torch
    from snappy_jl_b200 import synth
    raw = synth.mix(40, seed=8, tail=999)
    other = synth.mix(40, seed=9, tail=999)
    stream, index = dev.compress_device(to_dev(raw), want_index=True)
    _, index2 = dev.compress_device(to_dev(other), want_index=True)
    side = snappy.pack_index(index.cpu().numpy().astype(np.uint64), raw.size)
    assert len(side) < 200
    idx, ulen, slen = snappy.unpack_index(side)
    assert ulen == raw.size and slen == stream.numel()
    back = dev.uncompress_device(stream, index=torch.from_numpy(idx.astype(np.int64)).cuda(), claimed=ulen)
    assert np.array_equal(back.cpu().numpy(), raw)
    assert dev.last_launch_count(1) <= 2                       # indexed decoder only, no parse
    wrong, _, _ = snappy.unpack_index(snappy.pack_index(index2.cpu().numpy().astype(np.uint64), other.size))
    back = dev.uncompress_device(stream, index=torch.from_numpy(wrong.astype(np.int64)).cuda(), claimed=ulen)
    assert np.array_equal(back.cpu().numpy(), raw)


def _adversarial(rng, n):
    """byte patterns aimed at the window kernel: repeats at distances around the 32-lane window, the ring
    size and its staging chunk, equal 4-grams inside one window, runs, near-copies with single-byte edits"""
    kind = int(rng.integers(0, 7))
    if kind == 0:      # short period
        p = rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8)
        return np.resize(p, n)
    if kind == 1:      # period near the ring / chunk sizes
        per = int(rng.choice([511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097]))
        p = rng.integers(0, 256, per, dtype=np.uint8)
        return np.resize(p, n)
    if kind == 2:      # words from a tiny alphabet: many equal 4-grams inside 32 bytes
        words = [bytes(rng.integers(97, 100, int(rng.integers(1, 6)), dtype=np.uint8)) for _ in range(6)]
        out = bytearray()
        while len(out) < n:
            out += words[int(rng.integers(0, len(words)))]
        return np.frombuffer(bytes(out[:n]), dtype=np.uint8)
    if kind == 3:      # runs of random length
        out = np.empty(n, dtype=np.uint8)
        i = 0
        while i < n:
            l = int(rng.integers(1, 200))
            out[i:i + l] = rng.integers(0, 256)
            i += l
        return out
    if kind == 4:      # a block repeated with single-byte edits (copies of 16..31 bytes, long copies)
        blk = rng.integers(0, 256, int(rng.integers(40, 3000)), dtype=np.uint8)
        out = np.resize(blk, n).copy()
        idx = rng.integers(0, max(n, 1), max(n // int(rng.integers(17, 300)), 1))
        out[idx] ^= 0x55
        return out
    if kind == 5:      # random with planted repeats at window-ish distances
        out = rng.integers(0, 256, n, dtype=np.uint8)
        for _ in range(n // 64):
            d = int(rng.choice([4, 5, 8, 16, 31, 32, 33, 63, 64, 65, 96, 127, 128]))
            l = int(rng.integers(4, 40))
            s = int(rng.integers(0, max(n - d - l, 1)))
            out[s + d: s + d + l] = out[s: s + l][: max(0, min(l, n - s - d))]
        return out
    return rng.integers(0, 4, n, dtype=np.uint8)   # 2-bit noise: dense hash collisions


def test_adversarial_patterns_bit_exact(dev, oracle):
    """differential test against the oracle on inputs built to stress the window kernel's trust rules,
    its ring and the decoder's window parse; both table placements"""
    rng = np.random.default_rng(20261018)
    for it in range(120):
        n = int(rng.choice([0, 1, 14, 15, 16, 17, 100, 4095, 4096, 65535, 65536, 65537, 70000, 131072 + 13,
                            int(rng.integers(1, 300000))]))
        raw = _adversarial(rng, n) if n else np.empty(0, dtype=np.uint8)
        want = oracle.compress_np(raw)
        opts = [{"l2_chains": 0}, {"smem_chains": 0}][it % 2]
        try:
            for k, v in opts.items():
                dev.set_option(k, v)
            stream, index = dev.compress_device(to_dev(raw) if n else torch_empty(), want_index=True)
        finally:
            dev.set_option("l2_chains", 14)
            dev.set_option("smem_chains", 6)
        got = stream.cpu().numpy()
        assert got.size == want.size and np.array_equal(got, want), "compress differs (case %d, n=%d)" % (it, n)
        back = dev.uncompress_device(stream, index=index, claimed=n)
        assert np.array_equal(back.cpu().numpy(), raw)
        back = dev.uncompress_device(stream, claimed=n)     # index-free parse
        assert np.array_equal(back.cpu().numpy(), raw)


def torch_empty():
    import torch
    return torch.empty(0, dtype=torch.uint8, device="cuda")


@pytest.mark.parametrize("nfrag,tail", [(2049, 0), (3072, 1), (4096, 0)])
def test_streamed_host_paths_boundaries(snappy, oracle, nfrag, tail):
    """host-buffer API right above the streaming thresholds: chunk-aligned and ragged sizes"""
    from snappy_jl_b200 import synth
    raw = synth.mix(nfrag, seed=400 + nfrag, tail=tail)
    got = snappy.compress_np(raw)
    want = oracle.compress_np(raw)
    assert got.size == want.size and np.array_equal(got, want)
    assert np.array_equal(snappy.uncompress_np(want), raw)


def test_streamed_uncompress_incompressible(snappy, oracle):
    """a > 64 MiB stream made of 64 KiB literals only: every segment boundary falls inside a literal"""
    from snappy_jl_b200 import synth
    raw = synth.random_bytes(80 << 20, seed=77)
    want = oracle.compress_np(raw)
    assert want.size > raw.size
    assert np.array_equal(snappy.uncompress_np(want), raw)
    assert np.array_equal(snappy.compress_np(raw), want)


def test_host_api_pinned_and_pageable_buffers_agree(snappy, oracle):
    """the host-buffer C ABI takes pinned buffers straight and bounces pageable ones (a plain Vector{UInt8},
    src/Snappy.jl:25,48) through pinned slots: same bytes either way, both directions, at a streamed size"""
    import ctypes
    import torch
    from snappy_jl_b200 import synth
    raw = synth.mix(2100, seed=77, tail=333)
    want = oracle.compress_np(raw)
    lib = snappy._abi.lib()
    cap = snappy.maxlength_compressed(raw.size)
    pin_in = torch.from_numpy(raw.copy()).pin_memory()
    pin_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    pin_back = torch.empty(raw.size, dtype=torch.uint8).pin_memory()
    pg_in = raw.copy()
    pg_out = np.empty(cap, dtype=np.uint8)
    pg_back = np.empty(raw.size, dtype=np.uint8)
    for name, (i, o, b) in {"pinned": (pin_in.data_ptr(), pin_out.data_ptr(), pin_back.data_ptr()),
                            "pageable": (pg_in.ctypes.data, pg_out.ctypes.data, pg_back.ctypes.data),
                            "pinned in, pageable out": (pin_in.data_ptr(), pg_out.ctypes.data, pg_back.ctypes.data)}.items():
        ol = ctypes.c_size_t(cap)
        assert lib.snappy_b200_compress(i, raw.size, o, ctypes.byref(ol)) == 0, name
        got = (pin_out.numpy() if o == pin_out.data_ptr() else pg_out)[: ol.value]
        assert ol.value == want.size and np.array_equal(got, want), name
        bl = ctypes.c_size_t(raw.size)
        assert lib.snappy_b200_uncompress(o, ol.value, b, ctypes.byref(bl)) == 0, name
        back = pin_back.numpy() if b == pin_back.data_ptr() else pg_back
        assert bl.value == raw.size and np.array_equal(back, raw), name
        back[:] = 0
