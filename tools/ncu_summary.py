#!/usr/bin/env python3
"""Summaries of the ncu evidence under profiles/ (run here, no GPU needed):

  ncu_summary.py capture <file.ncu-rep> <out.json> [note]   one kernel launch of an `ncu --set full` capture
  ncu_summary.py launches <launches.csv> <out.txt> [title]  per-kernel totals of a gpu__time_duration launch list
"""
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3,
        "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9}


def capture(rep, out, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def num(name):
        i = col[name]
        return float(vals[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    def txt(name):
        return vals[col[name]]
    d = {"kernel": txt("Kernel Name"), "command": note,
         "duration_s": num("gpu__time_duration.sum"),
         "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
         "registers_per_thread": txt("launch__registers_per_thread"), "block_size": txt("launch__block_size"),
         "cluster_size": txt("launch__cluster_size"), "clusters_active": txt("launch__cluster_max_active"),
         "smem_per_block": "%s %s" % (txt("launch__shared_mem_per_block_dynamic"),
                                      units[col["launch__shared_mem_per_block_dynamic"]]),
         "issue_active_pct": txt("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "inst_executed": txt("smsp__inst_executed.sum"),
         "smem_wavefronts": txt("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
         "smem_bank_conflicts": txt("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
         "l2_hit_pct": txt("lts__t_sector_hit_rate.pct")}
    d["dram_bytes_total"] = d["dram_bytes_read"] + d["dram_bytes_write"]
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


def launches(path, out, title=""):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = {}
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0) * 1e3
        name = r["Kernel Name"].replace("void ", "").replace("<unnamed>::", "").split("(")[0]
        c, t = tot.get(name, (0, 0.0))
        tot[name] = (c + 1, t + v)
    total = sum(t for _, t in tot.values())
    with open(out, "w") as f:
        f.write(title + "\n")
        f.write("%-46s %6s %12s %7s\n" % ("kernel", "count", "total ms", "share"))
        for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-46s %6d %12.3f %6.2f%%\n" % (name[:46], c, t, 100 * t / total))
        f.write("%-46s %6d %12.3f\n" % ("total", sum(c for c, _ in tot.values()), total))
    print(open(out).read()[:1500])


if __name__ == "__main__":
    {"capture": capture, "launches": launches}[sys.argv[1]](*sys.argv[2:])
