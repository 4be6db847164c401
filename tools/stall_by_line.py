#!/usr/bin/env python
"""Join an ncu source page (per-SASS-instruction stall samples and executed counts) with the line
info of the same kernel in the built library, and print where a kernel's time goes per SOURCE line.

  python tools/stall_by_line.py gpurun_out/prof.ncu-rep <kernel-substring> [--lib path.so] [--top 40]

ncu's CSV export of the source page has no file:line column, so the mapping comes from
`nvdisasm --print-line-info` of the cubin inside the .so (needs the same build as the capture: the
script checks that the opcode sequence of the two listings agrees and says so when it does not).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_listing(lib, want):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True,
                         text=True).stdout
    kernels = {}
    cur, line_at = None, None
    for ln in txt.splitlines():
        m = re.match(r"\.text\.(\S+):", ln)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            line_at = None
            continue
        if cur is None:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line_at = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            cur.append((int(m.group(1), 16), m.group(2).strip(), line_at))
    demangled = {}
    for k in kernels:
        d = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        demangled[k] = d
    base = want.split("<")[0]  # ncu and c++filt spell template arguments differently: match on the name
    hits = [k for k in kernels if base in demangled[k] or base in k]
    return kernels, demangled, hits


def ncu_source(rep, want):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "header": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["header"] is None:
            cur["header"] = row
        elif cur is not None:
            cur["rows"].append(row)
    return [b for b in blocks if want in b["name"]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--lib", default=os.path.join(ROOT, "snappy.jl_b200", "libsnappy_b200.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0, help="which captured launch of that kernel")
    ap.add_argument("--ranges", default="", help="file:lo-hi,... print totals for these line ranges too")
    a = ap.parse_args()
    blocks = ncu_source(a.report, a.kernel)
    if not blocks:
        sys.exit("no kernel matching %r in %s" % (a.kernel, a.report))
    b = blocks[min(a.launch, len(blocks) - 1)]
    h = b["header"]
    ia, isrc = h.index("Address"), h.index("Source")
    isamp = h.index("Warp Stall Sampling (All Samples)")
    iexe = h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    kernels, dem, hits = sass_listing(a.lib, a.kernel)
    # the template arguments are spelled differently by ncu and c++filt: take the hit whose length matches
    cand = [k for k in hits if len(kernels[k]) == len(b["rows"])]
    if not cand:
        sys.exit("no kernel in the library has %d instructions (capture from another build?)" % len(b["rows"]))
    best, agree = None, -1
    for k in cand:
        n = sum(1 for r, s in zip(b["rows"], kernels[k]) if r[isrc].split()[0:1] == s[1].split()[0:1]
                or r[isrc].strip().split(" ")[0] == s[1].split(" ")[0])
        if n > agree:
            best, agree = k, n
    lst = kernels[best]
    print("# %s\n# %d SASS instructions, opcode agreement %d/%d" % (dem[best], len(lst), agree, len(lst)))
    per_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot_s = tot_e = 0
    for r, s in zip(b["rows"], lst):
        samp, exe = int(r[isamp] or 0), int(r[iexe] or 0)
        key = s[2] or ("?", 0)
        per_line[key][0] += samp
        per_line[key][1] += exe
        for i, c in stall_cols:
            v = int(r[i] or 0)
            if v:
                per_line[key][2][c[6:]] += v
        tot_s += samp
        tot_e += exe
    print("# total samples %d, warp instructions executed %d" % (tot_s, tot_e))
    print("%-28s %8s %6s %12s %6s  top stalls" % ("file:line", "samples", "%", "inst", "%"))
    for key, (samp, exe, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[: a.top]:
        tops = ", ".join("%s %d" % kv for kv in st.most_common(3))
        print("%-28s %8d %6.2f %12d %6.2f  %s" % ("%s:%d" % key, samp, 100.0 * samp / max(tot_s, 1), exe,
                                                  100.0 * exe / max(tot_e, 1), tops))
    if a.ranges:
        print("# ranges")
        for spec in a.ranges.split(","):
            f, r = spec.split(":")
            lo, hi = [int(x) for x in r.split("-")]
            s = sum(v[0] for k, v in per_line.items() if k[0] == f and lo <= k[1] <= hi)
            e = sum(v[1] for k, v in per_line.items() if k[0] == f and lo <= k[1] <= hi)
            print("%-28s %8d %6.2f %12d %6.2f" % (spec, s, 100.0 * s / max(tot_s, 1), e, 100.0 * e / max(tot_e, 1)))


if __name__ == "__main__":
    main()
