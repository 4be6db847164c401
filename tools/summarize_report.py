"""One ncu report (any number of captured launches) -> a small text summary for profiles/.
    python tools/summarize_report.py gpurun_out/x.ncu-rep profiles/r02c_x.txt "how it was captured"
Launches of the same kernel are listed one after the other."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]

rep, dst, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
with open(dst, "w") as f:
    f.write("# %s\n# %s\n" % (rep.split("/")[-1], note))
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        f.write("\nKernel Name = %s\n" % d.get("Kernel Name", "?").split("(")[0])
        for k in KEYS:
            if k in d and d[k] != "":
                f.write("%s [%s] = %s\n" % (k, u.get(k, ""), d[k]))
print("wrote", dst, len(rows) - 2, "launches")
