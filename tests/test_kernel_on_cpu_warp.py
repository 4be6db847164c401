"""CPU: the REAL compress kernel source (csrc/compress_window.cuh + compress_chain.cuh, every instantiation of
k_compress_window) compiled with -DSB200_CPU_EMU and executed as one 32-lane warp of coroutines
(tools/cpu_warp/cuda_shim.h), fragment by fragment against the oracle.  Unlike tools/emulate_window.c this is not a
model of the algorithm: it is the kernel's own code (the few inline-PTX statements have C++ twins; the CUDA build's
SASS is unchanged by them), so lane logic, masks, ring addressing, table commits and emission are checked where no GPU
exists.  Timing, occupancy and inter-lane memory ordering are what only the GPU tests see."""
import os
import subprocess

import numpy as np
import pytest

from conftest import DATA, ROOT

FILES = ["html", "alice29.txt", "fireworks.jpeg", "geo.protodata", "smallrandom1.bin", "sample-tweet.json"]


@pytest.fixture(scope="module")
def warp(tmp_path_factory):
    d = tmp_path_factory.mktemp("cpu_warp")
    exe = str(d / "run_window_kernel")
    obj = str(d / "oracle.o")
    subprocess.check_call(["gcc", "-O2", "-c", "-o", obj, os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DSB200_CPU_EMU", "-DSB200_EXPERIMENTS", "-I" + os.path.join(ROOT, "tools", "cpu_warp"),
                           "-o", exe, os.path.join(ROOT, "tools", "cpu_warp", "run_window_kernel.cpp"), obj])
    return exe


def run(exe, files, table, rules, ring, kernel="window", slowcont=0, mixed=0, two=0, pipe=0, uni=0):
    env = dict(os.environ, TABLE=table, RULES=str(rules), RING=str(ring), KERNEL=kernel, SLOWCONT=str(slowcont),
               MIXED=str(mixed), TWO=str(two), PIPE=str(pipe), UNI=str(uni))
    p = subprocess.run([exe] + files, env=env, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if "fragments" in l]
    assert len(lines) == len(files)
    for l in lines:
        assert " 0 mismatches" in l, l


@pytest.mark.parametrize("table,ring", [("smem", 2048), ("global", 1024)])
@pytest.mark.parametrize("rules", [0, 1, 2])
def test_kernel_source_matches_oracle_on_fixtures(warp, table, ring, rules):
    run(warp, [os.path.join(DATA, f) for f in FILES], table, rules, ring)


@pytest.mark.parametrize("table,ring", [("smem", 2048), ("global", 1024)])
@pytest.mark.parametrize("rules", [0, 2])
def test_mixed_kernel_source_matches_oracle_on_fixtures(warp, table, ring, rules):
    """k_compress_window_mixed, the default launch form: one CTA holds both table placements (global-table warps on
    the low warp numbers, shared-table warps on the high ones); the harness runs either role of a two-warp CTA"""
    run(warp, [os.path.join(DATA, f) for f in FILES], table, rules, ring, mixed=1)


@pytest.mark.parametrize("table,ring", [("smem", 2048), ("global", 1024)])
def test_two_window_round_source_matches_oracle(warp, table, ring, tmp_path):
    """option `two`: 64 positions per round, the second window re-validated against the committed table instead of
    evaluated again -- every fixture, and the boundary sizes / adversarial patterns, byte-identical to the oracle"""
    run(warp, [os.path.join(DATA, f) for f in FILES + ["urls.10K", "kppkn.gtb"]], table, 0, ring, mixed=1, two=1)
    rng = np.random.default_rng(5)
    blobs = [bytes(rng.integers(0, 3, 70000, dtype=np.uint8)), b"ab" * 40000, b"\0" * 66000,
             (b"0123456789abcdef" * 5 + b"X") * 900, bytes(rng.integers(97, 101, 200000, dtype=np.uint8)),
             b"".join(bytes([65 + (i * 7) % 23]) * (1 + i % 40) for i in range(4000))]
    text = b" ".join(bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(30000))
    blobs += [text[:n] for n in (15, 16, 17, 47, 48, 49, 63, 64, 65, 79, 80, 81, 95, 96, 97, 4096, 65535, 65536, 65537)]
    files = []
    for i, b in enumerate(blobs):
        p = tmp_path / ("t%02d.bin" % i)
        p.write_bytes(b)
        files.append(str(p))
    run(warp, files, table, 0, ring, mixed=1, two=1)


@pytest.mark.parametrize("table,ring,pipe,uni", [("smem", 2048, 1, 0), ("global", 1024, 1, 0), ("global", 1024, 0, 1),
                                                 ("smem", 2048, 1, 1)])
def test_pipelined_round_and_shared_code_sources_match_oracle(warp, table, ring, pipe, uni, tmp_path):
    """options `pipe` (compress_pipe.cuh: the round's loads issued one round ahead, validated by re-reading the table
    entry) and `unified` (one copy of the round's code, table placement per warp at run time): experiments, exact"""
    run(warp, [os.path.join(DATA, f) for f in FILES + ["urls.10K", "kppkn.gtb"]], table, 0, ring, mixed=1, pipe=pipe, uni=uni)
    rng = np.random.default_rng(6)
    blobs = [bytes(rng.integers(0, 3, 70000, dtype=np.uint8)), b"ab" * 40000, b"\0" * 66000,
             (b"0123456789abcdef" * 5 + b"X") * 900, bytes(rng.integers(97, 101, 200000, dtype=np.uint8)),
             b"".join(bytes([65 + (i * 7) % 23]) * (1 + i % 40) for i in range(4000))]
    text = b" ".join(bytes(rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8)) for _ in range(30000))
    blobs += [text[:n] for n in (15, 16, 17, 47, 48, 49, 63, 64, 65, 79, 80, 81, 95, 96, 97, 4096, 65535, 65536, 65537)]
    files = []
    for i, b in enumerate(blobs):
        p = tmp_path / ("p%02d.bin" % i)
        p.write_bytes(b)
        files.append(str(p))
    run(warp, files, table, 0, ring, mixed=1, pipe=pipe, uni=uni)


def test_kernel_source_on_boundary_sizes_and_patterns(warp, tmp_path):
    rng = np.random.default_rng(77)
    words = [rng.integers(97, 123, int(rng.integers(2, 9)), dtype=np.uint8).tobytes() for _ in range(200)]
    text = b" ".join(words[i] for i in rng.integers(0, 200, 40000))
    blobs = [text[:n] for n in (1, 14, 15, 16, 17, 31, 32, 33, 60, 61, 76, 77, 255, 4096, 65535, 65536, 65537,
                                65536 + 15, 65536 + 16, 2 * 65536)]
    blobs += [bytes(range(60)), b"a" * 70 + bytes(range(100, 130)), b"ab" * 3000 + bytes(range(256)) * 3,
              b"\0" * 70000, bytes(rng.integers(0, 4, 70000, dtype=np.uint8)),
              (b"0123456789abcdef" * 5 + b"X") * 900]
    files = []
    for i, b in enumerate(blobs):
        p = tmp_path / ("b%02d.bin" % i)
        p.write_bytes(b)
        files.append(str(p))
    for table, ring in (("smem", 2048), ("global", 1024)):
        for rules in (0, 2):
            run(warp, files, table, rules, ring)


@pytest.mark.parametrize("table", ["smem", "global"])
def test_step_wise_chain_kernel_source_matches_oracle(warp, table):
    """k_compress_chain (option window=0; also the slow paths of the window kernel)"""
    run(warp, [os.path.join(DATA, f) for f in FILES], table, 0, 2048, kernel="chain")


@pytest.mark.parametrize("table,ring", [("smem", 2048), ("global", 1024)])
def test_slowcont_variant_source_matches_oracle(warp, table, ring):
    """option `slowcont` (off by default: byte-identical on the B200 but 3.5 % slower): long copies are extended inside
    the hop loop and the chain goes on in the same window"""
    run(warp, [os.path.join(DATA, f) for f in FILES + ["urls.10K"]], table, 0, ring, slowcont=1)


# ---------------------------------------------------------------------------------------------- the decoder
@pytest.fixture(scope="module")
def decode_warp(tmp_path_factory):
    d = tmp_path_factory.mktemp("cpu_warp_decode")
    exe = str(d / "run_decode_kernel")
    obj = str(d / "oracle.o")
    subprocess.check_call(["gcc", "-O2", "-c", "-o", obj, os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DSB200_CPU_EMU", "-DSB200_EXPERIMENTS", "-I" + os.path.join(ROOT, "tools", "cpu_warp"),
                           "-o", exe, os.path.join(ROOT, "tools", "cpu_warp", "run_decode_kernel.cpp"), obj])
    return exe


def run_decode(exe, mode, files):
    p = subprocess.run([exe, mode] + files, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert len(lines) == len(files)
    for l in lines:
        assert "0 mismatches" in l or "header rejected" in l, l
    return lines


def test_indexed_decoder_source_round_trips(decode_warp, tmp_path):
    """k_decode_fragments (the fused window-parallel decoder) on oracle-compressed fixtures and patterns:
    overlapping copies (runs, short periods), long literals, 64-byte copy splitting"""
    rng = np.random.default_rng(5)
    blobs = [b"\0" * 70000, b"ab" * 40000, (b"0123456789abcdef" * 5 + b"X") * 900, bytes(range(256)) * 300,
             bytes(rng.integers(0, 4, 70000, dtype=np.uint8)), bytes(rng.integers(0, 256, 66000, dtype=np.uint8)),
             b"x", b"a" * 70 + bytes(range(100, 130))]
    files = [os.path.join(DATA, f) for f in FILES + ["urls.10K", "kppkn.gtb"]]
    for i, b in enumerate(blobs):
        p = tmp_path / ("d%02d.bin" % i)
        p.write_bytes(b)
        files.append(str(p))
    run_decode(decode_warp, "indexed", files)


def test_exact_decoder_source_matches_oracle_statuses(decode_warp, oracle, tmp_path):
    """k_decode_serial: the reference's decoder semantics (src/internal.jl:411-527) -- foreign streams, the
    reference's must-throw streams (test/runtests.jl:62-123) with status AND failing output position, and streams the
    oracle made"""
    from conftest import corrupt_streams, read_data
    files = [os.path.join(DATA, f) for f in ("alice29.snappy", "baddata1.snappy", "baddata2.snappy", "baddata3.snappy")]
    for name, data in corrupt_streams(oracle):
        p = tmp_path / (name + ".snappy")
        p.write_bytes(data)
        files.append(str(p))
    for name in ("html", "sample-tweet.json", "fireworks.jpeg"):
        p = tmp_path / (name + ".oracle.snappy")
        p.write_bytes(oracle.compress(read_data(name)))
        files.append(str(p))
        q = tmp_path / (name + ".truncated.snappy")
        q.write_bytes(oracle.compress(read_data(name))[:-7])
        files.append(str(q))
    lines = run_decode(decode_warp, "exact", files)
    assert any("status 3 (oracle 3), produced 19791" in l for l in lines)   # baddata1


# ---------------------------------------------------------------------------------------------- the index-free parse
def test_clean_cut_retiling_matches_brute_force(oracle, tmp_path):
    """parse.cuh k_cut_low / k_cut_tiles / k_cut_fill (the re-tiling decode_parsed_locked runs when copies cross the
    64 KiB output boundaries): tile f must start at the greatest element start <= f * 65536 that no later copy reaches
    across -- block sizes that do and do not tile 64 KiB, hand-made 4-byte-offset copies hundreds of KiB back, a copy
    chain that leaves no clean cut but 0, the reference's foreign fixture, and ordinary streams (every boundary clean)"""
    from conftest import read_data
    from snappy_jl_b200 import synth
    exe = str(tmp_path / "run_parse_kernel")
    obj = str(tmp_path / "oracle.o")
    subprocess.check_call(["gcc", "-O2", "-c", "-o", obj, os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DSB200_CPU_EMU", "-I" + os.path.join(ROOT, "tools", "cpu_warp"),
                           "-o", exe, os.path.join(ROOT, "tools", "cpu_warp", "run_parse_kernel.cpp"), obj])
    raw = synth.mix(12, seed=17, tail=4321)
    files = []

    def put(name, data):
        p = tmp_path / name
        p.write_bytes(data)
        files.append(str(p))

    for block in (50000, 7777, 32768, 65536, 100000, 1000):
        parts = [oracle.encode32(raw.size)]
        for o in range(0, raw.size, block):
            s = oracle.compress(raw[o: o + block].tobytes())
            _, k = oracle.parse32(s, 0)
            parts.append(s[k:])
        put("blocky_%d.snappy" % block, b"".join(parts))
    rng = np.random.default_rng(99)

    def lit(parts, out, n):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        parts.append(bytes([62 << 2]) + (n - 1).to_bytes(3, "little") + b)
        out.extend(b)

    def cp(parts, out, length, offset):
        parts.append(bytes([((length - 1) << 2) | 3]) + offset.to_bytes(4, "little"))
        for _ in range(length):
            out.append(out[-offset])

    parts, out = [], bytearray()
    for i in range(24):
        lit(parts, out, 50000 + 977 * i)
        if i in (5, 6, 17):
            cp(parts, out, 64, 300000 + i)
        if i == 11:
            cp(parts, out, 40, 70000)
            cp(parts, out, 64, 131072)
    put("longrange.snappy", oracle.encode32(len(out)) + b"".join(parts))
    parts, out = [], bytearray()
    lit(parts, out, 70000)
    for i in range(40):
        lit(parts, out, 30000)
        cp(parts, out, 64, 65000)
    put("chained.snappy", oracle.encode32(len(out)) + b"".join(parts))
    put("mix.snappy", oracle.compress(raw.tobytes()))
    put("html.snappy", oracle.compress(read_data("html")))
    files.append(os.path.join(DATA, "alice29.snappy"))
    p = subprocess.run([exe] + files, env=dict(os.environ, CUTS="1"), capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert len(lines) == len(files)
    for l in lines:
        assert "[clean cuts ok]" in l and "0 mismatches" in l, l


def test_parse_kernel_sources_build_the_true_index(oracle, tmp_path):
    """parse.cuh passes A-E (one thread per 1 KiB of compressed bytes), launched as build_index_segment does, whole
    stream and cut into two segments: whenever the parse accepts a stream its index equals a sequential walk's, clean
    streams (Snappy.jl's, libsnappy's, current Google snappy's) are accepted, foreign / corrupt ones declined"""
    pa = pytest.importorskip("pyarrow")
    from conftest import corrupt_streams, read_data
    d = tmp_path
    exe = str(d / "run_parse_kernel")
    obj = str(d / "oracle.o")
    subprocess.check_call(["gcc", "-O2", "-c", "-o", obj, os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DSB200_CPU_EMU", "-DSB200_EXPERIMENTS", "-I" + os.path.join(ROOT, "tools", "cpu_warp"),
                           "-o", exe, os.path.join(ROOT, "tools", "cpu_warp", "run_parse_kernel.cpp"), obj])
    rng = np.random.default_rng(9)
    raws = {n: read_data(n) for n in ("html", "alice29.txt", "fireworks.jpeg", "urls.10K", "sample-tweet.json", "kppkn.gtb")}
    raws["zeros"] = b"\0" * 300000
    raws["period"] = (b"0123456789abcdef" * 5 + b"X") * 4000
    raws["random"] = bytes(rng.integers(0, 256, 200000, dtype=np.uint8))      # 64 KiB literals: chunks inside literals
    raws["mixed"] = raws["random"][:70000] + raws["html"][:50000] + raws["random"][70000:140001] + raws["zeros"][:65536]
    files, clean = [], []
    for name, raw in raws.items():
        for tag, comp in (("jl", oracle.compress(raw)), ("v1", oracle.compress_rules(raw, 1)),
                          ("google", pa.Codec("snappy").compress(raw, asbytes=True))):
            p = d / ("%s.%s.snappy" % (name, tag))
            p.write_bytes(comp)
            files.append(str(p))
            clean.append(str(p))
    for name, data in corrupt_streams(oracle):
        p = d / ("bad_%s.snappy" % name)
        p.write_bytes(data)
        files.append(str(p))
    files += [os.path.join(DATA, f) for f in ("alice29.snappy", "baddata1.snappy", "baddata2.snappy", "baddata3.snappy")]
    p = subprocess.run([exe] + files, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert len(lines) == len(files)
    for f, l in zip(files, lines):
        assert "0 mismatches" in l or "no body to parse" in l, l
        if f in clean:
            assert "stream valid/clean, parse accepted" in l, l
    assert any("alice29.snappy" in l and "not clean, parse declined" in l for l in lines)


# ---------------------------------------------------------------------------------------------- batched pages
@pytest.mark.parametrize("rules", [0, 2])
def test_page_kernel_sources_match_oracle_and_round_trip(tmp_path, rules):
    """k_compress_pages / k_decode_pages (one warp per independent stream, BASELINE config 4): 4 KiB pages, ragged and
    empty pages, one multi-fragment page; compressed bytes == the oracle's stream per page, decode gives the page back"""
    exe = str(tmp_path / "run_pages_kernel")
    obj = str(tmp_path / "oracle.o")
    subprocess.check_call(["gcc", "-O2", "-c", "-o", obj, os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DSB200_CPU_EMU", "-DSB200_EXPERIMENTS", "-I" + os.path.join(ROOT, "tools", "cpu_warp"),
                           "-o", exe, os.path.join(ROOT, "tools", "cpu_warp", "run_pages_kernel.cpp"), obj])
    files = [os.path.join(DATA, f) for f in ("html", "alice29.txt", "fireworks.jpeg", "sample-tweet.json", "geo.protodata")]
    p = subprocess.run([exe] + files, env=dict(os.environ, RULES=str(rules)), capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert len(lines) == len(files)
    for l in lines:
        assert "0 mismatches" in l, l


@pytest.mark.parametrize("rules,page", [(0, 4096), (2, 4096), (0, 1000), (0, 8192), (1, 5000)])
def test_page_window_kernel_source_matches_oracle_and_round_trips(tmp_path, rules, page):
    """k_compress_pages_window (pages <= 8 KiB: window round, table and whole page in shared memory, persistent warp):
    every page's stream equals the oracle's compress of that page; ragged and empty pages included"""
    exe = str(tmp_path / "run_pages_kernel")
    obj = str(tmp_path / "oracle.o")
    subprocess.check_call(["gcc", "-O2", "-c", "-o", obj, os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DSB200_CPU_EMU", "-DSB200_EXPERIMENTS", "-I" + os.path.join(ROOT, "tools", "cpu_warp"),
                           "-o", exe, os.path.join(ROOT, "tools", "cpu_warp", "run_pages_kernel.cpp"), obj])
    files = [os.path.join(DATA, f) for f in ("html", "alice29.txt", "fireworks.jpeg", "sample-tweet.json", "kppkn.gtb")]
    p = subprocess.run([exe] + files, env=dict(os.environ, RULES=str(rules), PAGE=str(page), KERNEL="window"),
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert len(lines) == len(files)
    for l in lines:
        assert "0 mismatches" in l, l
