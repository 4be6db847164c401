"""Parquet-page adapter (BASELINE config 4's real-world caller): every SNAPPY page
body of a Parquet file written by pyarrow decodes on the GPU, in one batched call, to what pyarrow's own
(foreign) snappy decoder and the oracle produce; the writer side compresses pages to the oracle's bytes."""
import io

import numpy as np
import pytest

pa = pytest.importorskip("pyarrow")
pq = pytest.importorskip("pyarrow.parquet")


def make_file(version, page_size, rows=120_000):
    rng = np.random.default_rng(11)
    words = np.array(["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta"])
    t = pa.table({
        "id": np.arange(rows, dtype=np.int64),
        "bucket": rng.integers(0, 50, rows).astype(np.int32),
        "price": np.round(rng.normal(100, 15, rows), 2),
        "tag": words[rng.integers(0, len(words), rows)],
        "text": ["row %d of the %s partition" % (i, words[i % 8]) for i in range(rows)],
    })
    buf = io.BytesIO()
    pq.write_table(t, buf, compression="snappy", data_page_size=page_size, data_page_version=version,
                   use_dictionary=["bucket", "tag"], row_group_size=rows // 2)
    return buf.getvalue()


def test_thrift_page_headers_cover_every_column_chunk():
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file("1.0", 4096, rows=20_000)
    pages = pp.list_pages(f)
    md = pq.ParquetFile(io.BytesIO(f)).metadata
    total = sum(md.row_group(r).column(c).total_compressed_size for r in range(md.num_row_groups)
                for c in range(md.num_columns))
    # headers + bodies tile the column chunks exactly
    assert sum(p["compressed"] + p["prefix"] for p in pages) < total
    assert len(pages) > 50 and all(p["codec"] in ("SNAPPY", "UNCOMPRESSED") for p in pages)
    assert any(p["kind"] == pp.DICTIONARY_PAGE for p in pages)


@pytest.mark.gpu
@pytest.mark.parametrize("version,page_size", [("1.0", 4096), ("2.0", 4096), ("1.0", 1 << 20)])
def test_parquet_pages_decode_on_gpu(oracle, version, page_size):
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file(version, page_size)
    pages, out, offs = pp.uncompress_pages(f)
    codec = pa.Codec("snappy")
    n_snappy = 0
    for p, o in zip(pages, offs):
        if p["codec"] != "SNAPPY" or p["compressed"] == 0:
            continue
        n_snappy += 1
        body = f[p["stream"]: p["stream"] + p["compressed"]]
        want = codec.decompress(body, decompressed_size=p["uncompressed"]).to_pybytes()
        got = out[o: o + p["uncompressed"]].tobytes()
        assert got == want
        if n_snappy % 37 == 0:
            assert oracle.uncompress(body) == want
    assert n_snappy > (200 if page_size == 4096 else 8)


@pytest.mark.gpu
def test_parquet_pages_compress_on_gpu(oracle):
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file("1.0", 4096, rows=30_000)
    pages, out, offs = pp.uncompress_pages(f)
    raw_pages = [out[o: o + p["uncompressed"]].tobytes() for p, o in zip(pages, offs) if o >= 0]
    ours = pp.compress_pages(raw_pages)
    assert len(ours) == len(raw_pages)
    for a, b in zip(ours, raw_pages):
        assert a == oracle.compress(b)


@pytest.mark.gpu
def test_parquet_page_bodies_are_byte_identical_to_the_writer_under_google_rules(snappy):
    """rules = 2: recompressing the decoded pages gives back the very bytes pyarrow's Google snappy
    wrote into the file."""
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file("1.0", 4096, rows=30_000)
    pages, out, offs = pp.uncompress_pages(f)
    sel = [(p, o) for p, o in zip(pages, offs) if o >= 0]
    try:
        snappy.set_rules(2)
        ours = pp.compress_pages([out[o: o + p["uncompressed"]].tobytes() for p, o in sel])
    finally:
        snappy.set_rules(0)
    assert len(ours) > 50
    for body, (p, _) in zip(ours, sel):
        assert body == bytes(f[p["stream"]: p["stream"] + p["compressed"]])


def test_page_header_reader_rejects_malformed_input_quickly():
    """page headers come out of files: mutated headers parse or raise ValueError, never hang or blow the stack"""
    import time
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file("2.0", 4096, rows=5_000)
    pages = pp.list_pages(f)
    rng = np.random.default_rng(3)
    t0 = time.time()
    ok = bad = 0
    for p in pages[:12]:
        hdr = bytearray(f[p["header"]: p["body"]])
        assert pp.read_page_header(bytes(hdr), 0)[1] == len(hdr)
        for _ in range(400):
            b = bytearray(hdr)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            try:
                fields, hl = pp.read_page_header(bytes(b) + bytes(rng.integers(0, 256, 32, dtype=np.uint8)), 0)
                assert hl > 0 and fields[3] >= 0
                ok += 1
            except ValueError:
                bad += 1
    assert ok + bad == 12 * 400 and bad > 100
    assert time.time() - t0 < 60
    # adversarial lengths: a list / string / nesting that claims more than the buffer holds
    for blob in (b"\x19\xfc\xff\xff\xff\xff\x0f", b"\x18\xff\xff\xff\xff\x0f", b"\x1c" * 64, b"\x15" + b"\xff" * 12):
        with pytest.raises(ValueError):
            pp.read_page_header(blob, 0)



# ---- hostile / malformed files: every length the kernel is handed is checked on the host first (no GPU needed)
def _patched_headers(monkeypatch, pp, edit):
    real = pp.read_page_header
    state = {"n": 0}

    def fake(buf, pos):
        f, hl = real(buf, pos)
        state["n"] += 1
        edit(state["n"], f)
        return f, hl

    monkeypatch.setattr(pp, "read_page_header", fake)


def test_page_body_past_the_file_image_is_rejected(monkeypatch):
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file("1.0", 4096, rows=5_000)

    def edit(n, fields):
        if n == 3:
            fields[3] = len(f)  # compressed_page_size larger than what is left of the file

    _patched_headers(monkeypatch, pp, edit)
    with pytest.raises(ValueError, match="past the end"):
        pp.list_pages(f)


def test_oversized_and_negative_page_sizes_are_rejected(monkeypatch):
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file("2.0", 4096, rows=5_000)

    def huge(n, fields):
        if n == 2:
            fields[3] = 1 << 31

    _patched_headers(monkeypatch, pp, huge)
    with pytest.raises(ValueError, match="2 GiB"):
        pp.list_pages(f)

    def levels(n, fields):
        if fields[1] == pp.DATA_PAGE_V2:
            fields[8][5] = fields[3] + 1  # more level bytes than the page holds: usize - prefix would go negative

    _patched_headers(monkeypatch, pp, levels)
    with pytest.raises(ValueError, match="level bytes"):
        pp.list_pages(f)


def test_caller_supplied_page_list_is_checked():
    from snappy_jl_b200 import parquet_pages as pp
    f = make_file("1.0", 4096, rows=5_000)
    pages = pp.list_pages(f)
    k = next(i for i, p in enumerate(pages) if p["codec"] == "SNAPPY" and p["compressed"] > 0)
    bad = [dict(p) for p in pages]
    bad[k]["compressed"] = len(f)  # would make k_decode_pages read past d_file
    with pytest.raises(ValueError, match="does not fit"):
        pp.uncompress_pages(f, bad)
    truncated = f[: pages[k]["stream"] + 3]
    with pytest.raises(ValueError, match="does not fit"):
        pp.uncompress_pages(truncated, pages)
