"""Regenerates tests/golden/compress_golden.json.

Source of the vectors: oracle/py_restatement.py, the 1-based pure-Python transcription of the
Julia reference (src/Snappy.jl, src/internal.jl, src/varint.jl).  Julia is not available in this
image, so these are NOT outputs of the reference binary itself: they are the outputs of an
independent second transcription, cross-checked here against the SHA-256 table that the survey
derived separately.  The C oracle and the CUDA path are both tested against
this file.  Small vectors also carry the full compressed bytes (hex).

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import py_restatement as P  # noqa: E402

DATA = os.path.join(ROOT, "tests", "data")

FILES = ["alice29.txt", "asyoulik.txt", "html", "html_x_4", "kppkn.gtb", "lcet10.txt", "fireworks.jpeg",
         "geo.protodata", "paper-100k.pdf", "plrabn12.txt", "urls.10K", "random1.bin", "random2.bin",
         "random3.bin", "smallrandom1.bin", "sample-tweet.json"]

# SURVEY.md Appendix C (derived at survey time by a separate throw-away restatement)
SURVEY_SHA = {
    "alice29.txt": (88039, "b3707b512e4cf11c6bde9603be53dfef10f147eb5c7c547ffbeb93ed3aeeaf66"),
    "asyoulik.txt": (77503, "4bf8701f8c369f13e679f52e938c8630d2a2920eba4003bfeeced8522d984aa9"),
    "html": (22843, "c7c94425c2b3516cf3d1c9824391b8453beb544f38dfdfa90eb8126103234b5a"),
    "html_x_4": (92234, "11e53110e963fa6dd4ef3d726cf2d897a7ab689c8edb90ccab3f462ef21872f3"),
    "kppkn.gtb": (69526, "b6513d28c84b3715f02a2697ddb3f6b56aab8f09f0b5950075762912ae5ae8d9"),
    "lcet10.txt": (234662, "39bd4948c8743f4862e7a83fc36113f1db80712a0feb7545c048d43d26354268"),
    "fireworks.jpeg": (123034, "4da5e82d77ebe3d77e4f827a294562df17b5dcf37dcdb30d516ee8544d3164a6"),
    "geo.protodata": (23335, "84356d0f45f9cf8547834eabaa8d4ec569c3e71c505828ab3321ffbd35370d11"),
    "paper-100k.pdf": (85304, "ad668e5050689de4486cca4851a67b81731ff77ae920dc78da2e5fc9ca36d7e5"),
    "plrabn12.txt": (319267, "30915f0a26ae2b882e7d8a6951dc3e844c8dd615b0a1a21c6dd69e8c8f958337"),
    "urls.10K": (335506, "a0b5838bba64270a4fd4ca45eae8fe79469fd3ac1e3eab7208b5cbce9fd6d810"),
    "random1.bin": (127284, "022e2e41ae533e03b120d0f29178bc16ccd0fdac9ff987a63fee771935355941"),
    "random2.bin": (92416, "cd9d24141fa3f2273b60ac3cb636defa56f7f9d523302985890d4aa8daebbeab"),
    "random3.bin": (108560, "fb1d85c2dcbdba67c02b47c73333e0be99005082b5b39ef54bd064c524d326f8"),
    "smallrandom1.bin": (419, "d2b59e74172aa7fb67961f5d25fa16be95c7134f239e3650b88424f806ec0e51"),
    "sample-tweet.json": (3469, "3dd7721222ebdd07cc772ec3cf2e73da9b7523ef85884f737c0ded0b3496107d"),
}


def edge_vectors():
    """name -> input bytes; the reference's edge strings (test/runtests.jl:64,76,127-137) + quirk probes"""
    v = {
        "empty": b"",
        "a": b"a",
        "ab": b"ab",
        "abc": b"abc",
        "b16": b"aaaaaaa" + b"b" * 16 + b"aaaaa" + b"abc",
        "b256": b"aaaaaaa" + b"b" * 256 + b"aaaaa" + b"abc",
        "b2047": b"aaaaaaa" + b"b" * 2047 + b"aaaaa" + b"abc",
        "b65536": b"aaaaaaa" + b"b" * 65536 + b"aaaaa" + b"abc",
        "abc_b65536": b"abcaaaaaaa" + b"b" * 65536 + b"aaaaa" + b"abc",
        "sentence": b"making sure we don't crash with corrupted input",
        "A100000": b"A" * 100000,
        "literal60": bytes(range(60)),                       # 60-byte literal takes the 2-byte header
        "copy_tail_2byte": b"a" * 70 + bytes(range(100, 130)),  # remainder after a 64-copy uses the 2-byte form
        "xyz65536": b"xyz" * 21845 + b"x",                    # exactly one full fragment
        "len14": bytes(range(14)),                           # below K_INPUT_MARGIN_BYTES: literal only
        "len15": b"abcdabcdabcdabc",
        "len16": b"abcdabcdabcdabcd",
        "len17": b"abcdabcdabcdabcda",
        "len31": b"abcdefgh" * 3 + b"abcdefg",
        "len32": b"abcdefgh" * 4,
    }
    return v


def main():
    out = {"_generator": "oracle/py_restatement.py via tests/golden/make_golden.py", "files": {}, "edges": {}}
    for name in FILES:
        raw = open(os.path.join(DATA, name), "rb").read()
        c = P.compress(raw)
        assert P.uncompress(c) == raw, name
        sha = hashlib.sha256(c).hexdigest()
        assert SURVEY_SHA[name] == (len(c), sha), (name, len(c), sha)
        out["files"][name] = {"raw_len": len(raw), "raw_sha256": hashlib.sha256(raw).hexdigest(),
                              "comp_len": len(c), "comp_sha256": sha}
        print(name, len(raw), len(c), sha[:16], flush=True)
    for name, raw in edge_vectors().items():
        c = P.compress(raw)
        assert P.uncompress(c) == raw, name
        e = {"raw_len": len(raw), "comp_len": len(c), "comp_sha256": hashlib.sha256(c).hexdigest()}
        if len(raw) <= 256:
            e["raw_hex"] = raw.hex()
        if len(c) <= 5000:
            e["comp_hex"] = c.hex()
        out["edges"][name] = e
        print(name, len(raw), len(c), flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "compress_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
