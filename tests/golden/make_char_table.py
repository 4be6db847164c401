#!/usr/bin/env python
"""Extracts the reference's literal CHAR_TABLE (256 x UInt16, /root/reference/src/internal.jl:47-80) and WORDMASK
(:83-85) into tests/golden/char_table.json.  The three decoders in this repo (oracle/snappy_oracle.c,
oracle/py_restatement.py, csrc/decompress.cuh) regenerate the table by formula; tests/test_oracle.py holds them to the
reference's own 256 entries.  Run in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_char_table.py [/root/reference]
"""
import json
import os
import re
import sys


def extract(path):
    src = open(os.path.join(path, "src", "internal.jl")).read()
    m = re.search(r"const global CHAR_TABLE = UInt16\[(.*?)\]", src, re.S)
    table = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
    m = re.search(r"const global WORDMASK = UInt32\[(.*?)\]", src, re.S)
    mask = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
    assert len(table) == 256 and len(mask) == 5
    line = src[: src.index("const global CHAR_TABLE")].count("\n") + 1
    return {"source": "src/internal.jl:%d (CHAR_TABLE), WORDMASK behind it" % line, "char_table": table,
            "wordmask": mask}


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "char_table.json")
    with open(out, "w") as f:
        json.dump(extract(ref), f)
        f.write("\n")
    print("wrote", out)
