#!/usr/bin/env python3
"""One line per kernel launch of an `ncu --set full` report: duration, DRAM traffic, issue / pipe utilisation, stalls.
    ncu_table.py <file.ncu-rep> [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
U = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
     "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9}
def num(r, name):
    i = col.get(name)
    if i is None or r[i] == "": return float("nan")
    return float(r[i].replace(",", "")) * U.get(units[i], 1.0)
stalls = [("bar", "barrier"), ("long", "long_scoreboard"), ("short", "short_scoreboard"), ("wait", "wait"), ("lg", "lg_throttle"),
          ("mio", "mio_throttle"), ("br", "branch_resolving"), ("math", "math_pipe_throttle"), ("noinst", "no_instruction")]
out = []
out.append("%-18s %8s %8s %7s %6s %6s %6s %5s %6s %6s | stalls per issue: %s" % ("kernel", "us", "dramMB", "GB/s", "issue%", "fp64%", "lsu%", "thr", "L1hit", "L2hit",
           " ".join(k for k, _ in stalls)))
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
    dur = num(r, "gpu__time_duration.sum")
    dram = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    st = " ".join("%.1f" % num(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % v) for _, v in stalls)
    out.append("%-18s %8.1f %8.1f %7.0f %6.1f %6.1f %6.1f %5.1f %6.1f %6.1f | %s" % (
        name[:18], dur * 1e6, dram / 1e6, dram / dur / 1e9, num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), num(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"), num(r, "l1tex__t_sector_hit_rate.pct"), num(r, "lts__t_sector_hit_rate.pct"), st))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
