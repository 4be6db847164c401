"""Turn the ncu captures in gpurun_out/ into the small text summaries committed under profiles/.
    python tools/summarize_profiles.py r01b
Reads gpurun_out/launches_<tag>.csv (launch list of bench.py) and gpurun_out/prof_*_<tag>.ncu-rep."""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]

def launches():
    src = os.path.join(G, "launches_%s.csv" % tag)
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg, unit = {}, ""
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[iu]
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "%s_launches_bench.csv" % tag), "w") as f:
        f.write("# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e` (%s)\n" % tag)
        f.write("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv (per-launch times are cold-cache and\n")
        f.write("# serialised - the two compress kernels normally run CONCURRENTLY and share the fragments; under ncu the first one does all the work: compare SHARES)\n")
        f.write("# unit seen: %s ; %d launches captured\nkernel,launches,total_ms,share\n" % (unit, sum(a[0] for a in agg.values())))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.3f,%.1f%%\n" % (k, n, t * scale, 100 * t / tot))

def report(name, note):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (name, tag))
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, "%s_%s.txt" % (tag, name)), "w") as f:
        f.write("# %s_%s : ncu --set full --clock-control none --import-source on ; %s\n" % (name, tag, note))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write("Kernel Name [] = %s\n" % d.get("Kernel Name", "?"))
            for k in KEYS:
                if k in d:
                    f.write("%s [%s] = %s\n" % (k, u.get(k, ""), d[k]))
            f.write("\n")

launches()
report("prof_window_smem", "tools/prof_run.py 16384 0 l2_chains=0 (1 GiB mix; the shared-table kernel alone, 6 warps per SM)")
report("prof_window_l2", "tools/prof_run.py 16384 0 smem_chains=0 (1 GiB mix; the global-table kernel alone, 14 warps per SM)")
report("prof_decode", "tools/prof_run.py 16384 0 (1 GiB mix, side index)")
print(os.listdir(P))
