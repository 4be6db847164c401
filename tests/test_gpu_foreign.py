"""GPU: streams this library did not write and corrupt streams at size.

The index-free path: the parse builds a tile index (a literal may straddle a 64 KiB output boundary: the tile then
starts at that literal), the indexed decoder runs over all tiles and marks the ones it rejects, and a bounded serial
walk decodes only those with the reference's checks.  `last_decode_path`: 0 side index, 1 parse + tiles, 2 parse + tiles
+ bounded serial walk, 3 whole-stream serial decoder."""
import time

import numpy as np
import pytest

from conftest import read_data

pytestmark = pytest.mark.gpu
FRAGMENT = 65536


def _path(snappy):
    return snappy._abi.lib().snappy_b200_get_option(b"last_decode_path")


def test_alice29_snappy_takes_the_parallel_path(snappy, oracle):
    """tests/data/alice29.snappy (a 32 KiB-block encoder's stream whose literals straddle the block boundaries;
    test/runtests.jl decodes it): tiles start at the straddling literals, no serial decoding at all"""
    import torch
    from snappy_jl_b200 import device
    data = read_data("alice29.snappy")
    want = read_data("alice29.txt")
    got = device.uncompress_device(torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda())
    assert got.cpu().numpy().tobytes() == want
    assert _path(snappy) == 1
    assert snappy.uncompress(data) == want   # host-buffer API, same path


def _blocky_stream(oracle, raw, block):
    """every `block` bytes compressed on their own (copies never leave their block), concatenated behind one header:
    a valid stream whose elements straddle the 64 KiB boundaries wherever the block size does not divide 65536"""
    parts = [oracle.encode32(raw.size)]
    for o in range(0, raw.size, block):
        s = oracle.compress(raw[o: o + block].tobytes())
        _, k = oracle.parse32(s, 0)
        parts.append(s[k:])
    return b"".join(parts)


@pytest.mark.parametrize("block", [32768, 50000, 65536, 7777])
def test_block_structured_foreign_streams(snappy, oracle, block):
    import torch
    from snappy_jl_b200 import device, synth
    raw = synth.mix(48, seed=17, tail=4321)
    stream = _blocky_stream(oracle, raw, block)
    assert oracle.uncompress_np(stream).tobytes() == raw.tobytes()
    d = torch.from_numpy(np.frombuffer(stream, dtype=np.uint8).copy()).cuda()
    got = device.uncompress_device(d)
    assert np.array_equal(got.cpu().numpy(), raw)
    # every block size decodes tile-parallel: where the blocks do not tile 64 KiB the tiles start at the nearest clean
    # cut (an element start no later copy reaches across, parse.cuh) instead of on the boundary
    assert _path(snappy) == 1
    # without the clean cuts the tiles a copy reaches out of go to the bounded serial walk -- never to the
    # whole-stream serial decoder -- and the bytes are the same
    device.set_option("clean_cuts", 0)
    try:
        got = device.uncompress_device(d)
        assert np.array_equal(got.cpu().numpy(), raw)
        assert _path(snappy) == (1 if 65536 % block == 0 else 2)
    finally:
        device.set_option("clean_cuts", 1)


def _lit(data):
    n = len(data)
    assert 60 < n <= 1 << 24
    return bytes([62 << 2]) + (n - 1).to_bytes(3, "little") + data     # literal, 3 length bytes


def _copy4(length, offset):
    assert 1 <= length <= 64
    return bytes([((length - 1) << 2) | 3]) + offset.to_bytes(4, "little")  # copy with a 4-byte offset


def test_long_range_copies_keep_the_rest_tile_parallel(snappy, oracle):
    """a hand-made stream whose copies use 4-byte offsets and reach hundreds of KiB back (no compressor in this repo
    emits them; the format allows them, src/internal.jl:28-30): the tiles around such a copy merge into one that
    begins at the clean cut below the copy's source, everything else stays 64 KiB-parallel; also a copy chain that
    leaves no clean cut at all (one tile = the whole stream), and an offset that reaches before the stream"""
    import torch
    from snappy_jl_b200 import device
    rng = np.random.default_rng(99)
    parts, raw = [], bytearray()

    def lit(n):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        parts.append(_lit(b))
        raw.extend(b)

    def cp(length, offset):
        parts.append(_copy4(length, offset))
        for _ in range(length):
            raw.append(raw[-offset])

    for i in range(24):
        lit(50000 + 977 * i)
        if i in (5, 6, 17):
            cp(64, 300000 + i)       # far back, across several tiles
        if i == 11:
            cp(40, 70000)
            cp(64, 131072)
    body = b"".join(parts)
    stream = np.frombuffer(oracle.encode32(len(raw)) + body, dtype=np.uint8).copy()
    want = np.frombuffer(bytes(raw), dtype=np.uint8)
    assert np.array_equal(oracle.uncompress_np(stream), want)
    got = device.uncompress_device(torch.from_numpy(stream).cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    assert _path(snappy) == 1
    assert snappy.uncompress(stream.tobytes()) == bytes(raw)            # host-buffer API
    # every tile reaches into its predecessor: no clean cut but 0 -- still correct (one long tile)
    parts, raw = [], bytearray()
    lit(70000)
    for i in range(40):
        lit(30000)
        cp(64, 65000)
    stream = np.frombuffer(oracle.encode32(len(raw)) + b"".join(parts), dtype=np.uint8).copy()
    want = np.frombuffer(bytes(raw), dtype=np.uint8)
    assert np.array_equal(oracle.uncompress_np(stream), want)
    got = device.uncompress_device(torch.from_numpy(stream).cuda())
    assert np.array_equal(got.cpu().numpy(), want)
    # an offset that reaches before the first byte: the reference's status (src/internal.jl:499)
    bad = np.frombuffer(oracle.encode32(70064) + _lit(bytes(70000)) + _copy4(64, 70001), dtype=np.uint8).copy()
    want_status = oracle.status_of_uncompress(bad)
    assert want_status != 0
    with pytest.raises(snappy.SnappyError) as e:
        device.uncompress_device(torch.from_numpy(bad).cuda())
    assert e.value.status == want_status


def test_corrupt_256_mib_stream_is_rejected_fast_with_the_reference_status(snappy, oracle):
    """one garbled spot in the middle of a 256 MiB stream: the reference's status at the cost of the parallel parse (the
    reference's per-element checks need the element header and its output position, never decoded bytes) -- not a
    serial pass over the stream"""
    import torch
    from snappy_jl_b200 import device, synth
    tile = synth.mix(512, seed=23)                    # 32 MiB
    reps = 8
    total = tile.size * reps
    parts, sizes = oracle.compress_fragments(tile, total, 0, 512)
    hdr = np.frombuffer(oracle.encode32(total), dtype=np.uint8)
    good = np.concatenate([hdr] + [parts] * reps)
    raw = np.tile(tile, reps)
    # garble 24 bytes somewhere in the middle; a spot that only hits literal bytes leaves the stream valid, so the
    # oracle picks the first spot that does not
    want_status, k = 0, 0
    while want_status == 0:
        bad = good.copy()
        pos = good.size // 2 + 12345 + 7919 * k
        bad[pos: pos + 24] ^= 0xA7
        want_status = oracle.status_of_uncompress(bad)
        k += 1
        assert k < 40
    d_bad = torch.from_numpy(bad).cuda()
    d_good = torch.from_numpy(good).cuda()
    out = torch.empty(total, dtype=torch.uint8, device="cuda")
    device.uncompress_device(d_good, out=out, claimed=total)          # warm-up, and the good stream still decodes
    assert torch.equal(out, torch.from_numpy(raw).cuda()) and _path(snappy) == 1
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with pytest.raises(snappy.SnappyError) as e:
        device.uncompress_device(d_bad, out=out, claimed=total)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert e.value.status == want_status
    assert _path(snappy) == 1   # the parse alone decides: header + output position of the failing element
    assert dt < 0.1, "rejecting the corrupt stream took %.1f ms" % (dt * 1e3)
    # a truncated stream, and one whose header lies about the length
    for s in (good[: good.size // 3], np.concatenate([np.frombuffer(oracle.encode32(total - 5), dtype=np.uint8), good[hdr.size:]])):
        want = oracle.status_of_uncompress(s)
        assert want != 0
        with pytest.raises(snappy.SnappyError) as e:
            device.uncompress_device(torch.from_numpy(np.ascontiguousarray(s)).cuda())
        assert e.value.status == want


def test_wrong_side_index_cannot_change_the_result(snappy, oracle):
    import torch
    from snappy_jl_b200 import device, synth
    raw = synth.mix(20, seed=3, tail=100)
    d = torch.from_numpy(raw).cuda()
    stream, index = device.compress_device(d, want_index=True)
    bogus = index.clone()
    bogus[5] += 3
    bogus[11] = 0
    back = device.uncompress_device(stream, index=bogus, claimed=raw.size)
    assert torch.equal(back, d) and _path(snappy) == 1
    back = device.uncompress_device(stream, index=index, claimed=raw.size)
    assert torch.equal(back, d) and _path(snappy) == 0
