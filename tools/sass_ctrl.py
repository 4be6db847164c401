#!/usr/bin/env python
"""SASS of one kernel with the scheduling control bits decoded (stall count, yield, the scoreboard an instruction
sets on write / read, the scoreboards it waits for), next to the source line of each instruction.

  python tools/sass_ctrl.py <lib.so> <kernel-substring> [--file compress_pipe.cuh] [--from LINE --to LINE]

The control word sits in bits 105..125 of the 128-bit instruction (Volta and later):
  stall 4 bits | yield 1 | write scoreboard 3 (7 = none) | read scoreboard 3 (7 = none) | wait mask 6 | reuse 4
Used to see which instruction first WAITS for a long-latency load (what a software pipeline hides or does not)."""
import argparse
import os
import re
import subprocess
import tempfile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("lib")
    ap.add_argument("kernel")
    ap.add_argument("--file", default=None)
    ap.add_argument("--from", dest="lo", type=int, default=0)
    ap.add_argument("--to", dest="hi", type=int, default=1 << 30)
    a = ap.parse_args()
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-instruction-encoding", "--print-line-info", os.path.join(tmp, cubin)],
                         capture_output=True, text=True).stdout.splitlines()
    on, src, pend = False, ("", 0), None
    for ln in txt:
        m = re.match(r"\.text\.(\S+):", ln)
        if m:
            on = a.kernel in m.group(1)
            if on:
                print("#", m.group(1))
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            src = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", ln)
        if m:
            pend = (int(m.group(1), 16), m.group(2).strip())
            continue
        m = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", ln)
        if m and pend:
            hi = int(m.group(1), 16)
            ctrl = (hi >> 41) & 0x1fffff
            stall, yld, wb, rb, wait = ctrl & 15, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 63
            if (a.file is None or src[0] == a.file) and a.lo <= src[1] <= a.hi:
                waits = "".join(str(i) for i in range(6) if wait >> i & 1)
                print("%06x  %-22s st%-2d %s W%s R%s wait[%-6s]  %s" % (
                    pend[0], "%s:%d" % src, stall, "Y" if yld else " ", "-" if wb == 7 else wb, "-" if rb == 7 else rb,
                    waits, pend[1]))
            pend = None
        if ln.startswith(".L_") and on and a.file is None:
            print(ln)


if __name__ == "__main__":
    main()
