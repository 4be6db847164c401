"""CPU: tools/emulate_window.c, the plain-C model of the window-parallel compressor
(csrc/compress_window.cuh: 32 positions per round; csrc/compress_wide.cuh: WW windows per round with
table re-validation), must reproduce the oracle's bytes on the reference's fixtures.  This pins the
ALGORITHM of the kernels where no GPU is available; the kernels themselves are checked in
test_gpu_parity.py."""
import os
import subprocess

import pytest

from conftest import DATA, ROOT

FILES = ["alice29.txt", "html", "urls.10K", "geo.protodata", "kppkn.gtb", "fireworks.jpeg", "random1.bin",
         "smallrandom1.bin", "sample-tweet.json"]


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("model") / "emulate_window")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "emulate_window.c"),
                           os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    return exe


@pytest.mark.parametrize("ww", [None, "2", "4"])
def test_window_model_matches_oracle(model, ww):
    env = dict(os.environ)
    if ww:
        env["WW"] = ww
    out = subprocess.run([model] + [os.path.join(DATA, f) for f in FILES], env=env, capture_output=True,
                         text=True, check=True).stdout
    lines = [l for l in out.splitlines() if "fragments" in l]
    assert len(lines) == len(FILES)
    for l in lines:
        assert " 0 mismatches" in l, l


@pytest.mark.parametrize("rules", ["1", "2"])
def test_window_model_matches_oracle_under_google_rules(model, rules):
    """the kLib instantiations: ip_limit n - 15, short 60-byte literal, per-fragment table, and for
    rules 2 the masked 15-bit bucket with up to 32768 entries -- against sjo_compress_fragment_rules"""
    env = dict(os.environ, RULES=rules)
    out = subprocess.run([model] + [os.path.join(DATA, f) for f in FILES], env=env, capture_output=True,
                         text=True, check=True).stdout
    lines = [l for l in out.splitlines() if "fragments" in l]
    assert len(lines) == len(FILES)
    for l in lines:
        assert " 0 mismatches" in l, l


def test_window_model_is_exact_for_a_64_wide_window(tmp_path):
    """two positions per lane (the round-2 direction in DESIGN.md 9): same bytes as the oracle, ~35 % fewer rounds"""
    exe = str(tmp_path / "emulate_window64")
    subprocess.check_call(["gcc", "-O2", "-DW=64", "-o", exe, os.path.join(ROOT, "tools", "emulate_window.c"),
                           os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    out = subprocess.run([exe] + [os.path.join(DATA, f) for f in FILES], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if "fragments" in l]
    assert len(lines) == len(FILES)
    for l in lines:
        assert " 0 mismatches" in l, l


@pytest.mark.parametrize("width", ["32", "64"])
def test_window_model_variants_stay_exact(tmp_path, width):
    """SLOWCONT / PRECISE: candidate refinements of the round (tools/emulate_window.c), exact by construction"""
    exe = str(tmp_path / "emulate_window_v")
    subprocess.check_call(["gcc", "-O2", "-DW=" + width, "-o", exe, os.path.join(ROOT, "tools", "emulate_window.c"),
                           os.path.join(ROOT, "oracle", "snappy_oracle.c")])
    env = dict(os.environ, SLOWCONT="1", PRECISE="1")
    out = subprocess.run([exe] + [os.path.join(DATA, f) for f in FILES], env=env, capture_output=True, text=True,
                         check=True).stdout
    lines = [l for l in out.splitlines() if "fragments" in l]
    assert len(lines) == len(FILES)
    for l in lines:
        assert " 0 mismatches" in l, l
