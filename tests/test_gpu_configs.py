"""GPU: BASELINE.json's configs 3 and 4 at their stated sizes, and the 2^32 - 1 byte limit of src/Snappy.jl:21.

Inputs of several GiB are built on the GPU by tiling a smaller buffer from the synthetic generator (generating 4 GiB
with numpy would take minutes); every check still compares real bytes with the oracle: sampled pages / fragments are
copied back and compressed by the oracle, and whole buffers are compared on the device."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FRAGMENT = 65536


def test_config4_one_million_pages_sampled_against_the_oracle(oracle):
    """2^20 independent 4 KiB pages through the batched API; every page round-trips, >= 1000 sampled pages are
    byte-identical to the oracle's compress of that page (own varint, table sized from the page: Snappy.jl:26-27)"""
    import torch
    from snappy_jl_b200 import device, synth
    npages, psz = 1 << 20, 4096
    base = torch.from_numpy(synth.pages(32768, psz, seed=44).reshape(-1)).cuda()      # 128 MiB of pages
    d_in = base.repeat(npages // 32768)                                                # 4 GiB
    # make every page distinct: its number goes into its first 8 bytes
    ids = torch.arange(npages, dtype=torch.int64, device="cuda")
    d_in.view(npages, psz)[:, :8] ^= ids.view(-1, 1).view(torch.uint8).view(npages, 8)
    in_off = ids * psz
    in_sz = torch.full((npages,), psz, dtype=torch.int32, device="cuda")
    out, out_off, out_sz = device.compress_batched_device(d_in, in_off, in_sz)
    back = torch.empty_like(d_in)
    sizes, statuses = device.uncompress_batched_device(out, out_off, out_sz, back, in_off, in_sz)
    assert int(statuses.abs().sum().item()) == 0
    assert bool((sizes == psz).all()) and torch.equal(back, d_in)
    rng = np.random.default_rng(4)
    picks = sorted(set([0, 1, npages - 1] + [int(x) for x in rng.integers(0, npages, 1100)]))
    assert len(picks) >= 1000
    sel = torch.tensor(picks, dtype=torch.int64, device="cuda")
    src = d_in.view(npages, psz)[sel].cpu().numpy()
    offs = out_off[sel].cpu().numpy()
    szs = out_sz[sel].cpu().numpy()
    cap = int((out_off[1] - out_off[0]).item())
    rows = out.view(npages, cap)[sel].cpu().numpy()
    for i, p in enumerate(picks):
        want = oracle.compress_np(src[i])
        assert int(szs[i]) == want.size and np.array_equal(rows[i, : want.size], want), "page %d" % p
    assert int(offs[1]) > int(offs[0])


def test_config3_one_gib_oracle_stream_without_index(oracle):
    """a 1 GiB stream PRODUCED BY THE ORACLE, handed over with no side index: index-free parse + decode"""
    import torch
    from snappy_jl_b200 import device, synth
    tile = synth.mix(2048, seed=33)                      # 128 MiB
    reps = 8
    total = tile.size * reps
    # fragments are independent (table reset per fragment, Snappy.jl:30) and the table size comes from the TOTAL
    # length: the oracle's stream of the tiled input is its header + the tile's fragment bytes, 8 times over
    parts, _ = oracle.compress_fragments(np.concatenate([tile, np.zeros(0, dtype=np.uint8)]), total, 0, 2048)
    check, _ = oracle.compress_fragments(np.tile(tile, 2), total, 2048, 3)   # second tile, from a real 2-tile input
    assert np.array_equal(check, parts[: check.size])
    hdr = np.frombuffer(oracle.encode32(total), dtype=np.uint8)
    stream = torch.from_numpy(np.concatenate([hdr] + [parts] * reps)).cuda()
    want = torch.from_numpy(tile).cuda().repeat(reps)
    got = device.uncompress_device(stream, index=None, claimed=total)
    assert got.numel() == total and torch.equal(got, want)
    assert device.last_launch_count(1) > 5                # the parse kernels ran (no index was given)


def test_four_gib_minus_one_round_trip(snappy, oracle):
    """the largest legal input, 2^32 - 1 bytes (Snappy.jl:21): 65536 fragments, the last one ragged by one byte;
    offsets beyond 2^32 appear nowhere, fragment / offset arithmetic near the limit is exercised"""
    import torch
    from snappy_jl_b200 import device, synth
    tile = synth.mix(1024, seed=55)                       # 64 MiB
    n = (1 << 32) - 1
    d_in = torch.from_numpy(tile).cuda().repeat(64)[:n]
    stream, index = device.compress_device(d_in, want_index=True)
    idx = index.cpu().numpy()
    assert idx.size == 65536 + 1 and int(idx[-1]) == stream.numel() and int(idx[0]) == 5
    assert bytes(stream[:5].cpu().numpy().tobytes()) == oracle.encode32(n)
    # sampled fragments (always the first and the ragged last one) against the oracle
    for f in [0, 1, 1023, 1024, 40000, 65534, 65535]:
        lo, hi = f * FRAGMENT, min((f + 1) * FRAGMENT, n)
        want = oracle.compress_one_fragment(d_in[lo:hi].cpu().numpy(), n)
        got = stream[int(idx[f]): int(idx[f + 1])].cpu().numpy()
        assert got.size == want.size and np.array_equal(got, want), "fragment %d" % f
    back = device.uncompress_device(stream, index=index, claimed=n)
    assert torch.equal(back, d_in)
    del back
    back = device.uncompress_device(stream, index=None, claimed=n)   # index-free parse on a 1.9 GB stream
    assert torch.equal(back, d_in)
    assert device.last_launch_count(1) > 5                 # ... and not the single-warp decoder (nfrag in 64 bits)
