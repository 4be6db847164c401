import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden", "compress_golden.json")

# the 15 round-trip files of test/runtests.jl:8-24
ROUNDTRIP_FILES = ["alice29.txt", "asyoulik.txt", "html", "html_x_4", "kppkn.gtb", "lcet10.txt",
                   "fireworks.jpeg", "geo.protodata", "paper-100k.pdf", "plrabn12.txt", "urls.10K",
                   "random1.bin", "random2.bin", "random3.bin", "smallrandom1.bin"]
ALL_FILES = ROUNDTRIP_FILES + ["sample-tweet.json"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def read_data(name):
    with open(os.path.join(DATA, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def snappy():
    import snappy_jl_b200
    return snappy_jl_b200


def edge_inputs():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    return make_golden.edge_vectors()


def corrupt_streams(oracle_mod):
    """The must-throw streams of test/runtests.jl:62-123, as (name, bytes)."""
    out = []
    dst = bytearray(oracle_mod.compress(b"making sure we don't crash with corrupted input"))
    dst[1] = (~dst[1]) & 0xFF   # runtests.jl:69-70 (1-based 2 and 4)
    dst[3] = dst[2]
    out.append(("flipped", bytes(dst)))
    dst = bytearray(oracle_mod.compress(b"A" * 100000))
    dst[0] = dst[1] = dst[2] = dst[3] = 0
    out.append(("header_lie_0", bytes(dst)))
    dst[0] = dst[1] = dst[2] = 0xFF
    dst[3] = 0x00
    out.append(("header_lie_2mb", bytes(dst)))
    for f in ("baddata1.snappy", "baddata2.snappy", "baddata3.snappy"):
        out.append((f, read_data(f)))
    out.append(("varint_f0", bytes([0xF0])))
    out.append(("varint_6", bytes([0x80, 0x80, 0x80, 0x80, 0x80, 0x0A])))
    out.append(("varint_7f", bytes([0xFB, 0xFF, 0xFF, 0xFF, 0x7F])))
    out.append(("copy_off0_a", bytes([0x40, 0x12, 0x00, 0x00])))
    out.append(("copy_off0_b", bytes([0x05, 0x12, 0x00, 0x00])))
    return out


def dictionary_fuzz(seed, count, maxwords=1 << 16):
    """test/runtests.jl:35-49: 64 words of 1..16 random bytes; inputs of 1..maxwords words."""
    rng = np.random.default_rng(seed)
    words = [rng.integers(0, 256, rng.integers(1, 17), dtype=np.uint8).tobytes() for _ in range(64)]
    for _ in range(count):
        picks = rng.integers(0, 64, rng.integers(1, maxwords + 1))
        yield b"".join(words[i] for i in picks)
