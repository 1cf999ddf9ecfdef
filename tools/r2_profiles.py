#!/usr/bin/env python3
"""Turn the files of tools/gpu_r2_final.sh (gpurun_out/) into the committed summaries under profiles/ (runs here, no GPU)."""
import csv, json, shutil, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
U = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
     "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9}
short = lambda n: n.replace('void ', '').replace('<unnamed>::', '').split('(')[0]

bench_line = json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1])
NB = bench_line['config']['networks_per_gpu_per_step'] // max(1, bench_line['config']['batches_in_flight_per_gpu'])   # networks per batch
LIST_CMD = "python bench.py --engines 1 --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
# ---- launch list of one batch + DRAM traffic of the direct stage
lines = [l for l in open('gpurun_out/launches_final.csv') if not l.startswith('==')]
ids = {}
for r in csv.DictReader(lines):
    ids.setdefault(int(r['ID']), {'name': r['Kernel Name']})[r['Metric Name']] = float(r['Metric Value'].replace(',', '')) * U.get(r['Metric Unit'], 1.0)
batches, cur = [], None
for i in sorted(ids):
    if short(ids[i]['name']) == 'k_generate':
        cur = []
        batches.append(cur)
    if cur is not None:
        cur.append(i)
b = batches[1]
tot = {}
for i in b:
    d = ids[i]
    c = tot.setdefault(short(d['name']), [0, 0.0, 0.0])
    c[0] += 1
    c[1] += d['gpu__time_duration.sum']
    c[2] += d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
T, D = sum(c[1] for c in tot.values()), sum(c[2] for c in tot.values())
with open('profiles/r02_launch_list_final_summary.txt', 'w') as f:
    f.write("r02 final launch list: ONE batch of %d networks (the timed step) of `%s`\n" % (NB, LIST_CMD))
    f.write("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (serialised, cold cache); "
            "%d launches, %.2f ms summed, %.1f GB of DRAM traffic\n" % (len(b), T * 1e3, D / 1e9))
    f.write("%-40s %6s %10s %7s %10s %8s\n" % ("kernel", "count", "total ms", "share", "dram MB", "GB/s"))
    for n, (c, t, dd) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write("%-40s %6d %10.3f %6.2f%% %10.1f %8.0f\n" % (n[:40], c, t * 1e3, 100 * t / T, dd / 1e6, dd / t / 1e9 if t > 0 else 0))
direct = ('k_hoff', 'k_init', 'k_init_hash', 'k_mark', 'k_select', 'k_incident', 'k_pivots', 'k_items_a', 'k_items_b', 'k_items_scan',
          'k_items_c', 'k_final_a', 'k_final_nodes', 'k_final_edges', 'k_final_b', 'k_sm_edges0', 'k_sm_coupling', 'k_sm_pivot',
          'k_sm_update', 'k_sm_back', 'k_sm_dense_smem', 'k_sm_dense', 'k_sm_dense_fill')
ds = [i for i in b if short(ids[i]['name']).split('<')[0] in direct]
rd = sum(ids[i].get('dram__bytes_read.sum', 0) for i in ds)
wr = sum(ids[i].get('dram__bytes_write.sum', 0) for i in ds)
tt = sum(ids[i]['gpu__time_duration.sum'] for i in ds)
json.dump({"command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 " + LIST_CMD,
           "batch": NB, "launches": len(ds), "duration_s_sum": tt, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr,
           "note": "sum of dram__bytes_read+write over the %d launches of the direct stage (cnt_sm_symbolic + cnt_sm_solve) of ONE batch of %d networks x 4 "
                   "configurations (the timed step), ncu metrics pass of the bench command (profiles/r02_launch_list_final.csv); serialised, cold cache: "
                   "%.2f ms summed kernel time" % (len(ds), NB, tt * 1e3)}, open('profiles/r02_direct_stage_ncu.json', 'w'), indent=1)
shutil.copy('gpurun_out/launches_final.csv', 'profiles/r02_launch_list_final.csv')
shutil.copy('gpurun_out/bench_final.log', 'profiles/r02_bench_1gpu_final.json')
shutil.copy('gpurun_out/bench_reference.log', 'profiles/r02_bench_reference_arm.json')
print(open('profiles/r02_launch_list_final_summary.txt').read()[:2600])
print("direct stage: %.1f GB, %d launches, %.2f ms" % ((rd + wr) / 1e9, len(ds), tt * 1e3))

# ---- k_tiles capture
raw = subprocess.run(["ncu", "-i", "gpurun_out/r02_tiles_final.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
num = lambda n: float(vals[col[n]].replace(',', '')) * U.get(units[col[n]], 1.0)
bl = json.loads(open('profiles/r02_bench_1gpu_final.json').read().strip().splitlines()[-1])
tests = bl['junction_tests'] / bl.get('stages_batches', bl['steps'])     # per batch = per launch of k_tiles
d = {"kernel": "k_tiles (csrc/junctions.cu): one CTA of 256 threads per tile of 8x8 cells; one batch of %d networks of 100x100 um" % NB,
     "command": "ncu --set full --clock-control none --import-source on -k regex:k_tiles -c 1 " + LIST_CMD,
     "duration_s": num("gpu__time_duration.sum"), "grid": vals[col["launch__grid_size"]], "block": vals[col["launch__block_size"]],
     "registers_per_thread": vals[col["launch__registers_per_thread"]], "dynamic_smem_per_block_KB": num("launch__shared_mem_per_block_dynamic"),
     "occupancy_achieved_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
     "fp64_pipe_pct_of_peak": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
     "issue_slots_busy_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
     "l1_hit_rate_pct": num("l1tex__t_sector_hit_rate.pct"), "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
     "threads_active_per_instruction_of_32": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
     "inst_executed": num("smsp__inst_executed.sum"), "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
     "smem_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
     "smem_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
     "stalls_per_issued_instruction": {k: num("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k) for k in
                                       ("barrier", "short_scoreboard", "long_scoreboard", "wait", "branch_resolving", "math_pipe_throttle", "mio_throttle", "not_selected")},
     "reference_equivalent_tests_per_launch": tests}
d["dram_bytes_total"] = d["dram_bytes_read"] + d["dram_bytes_write"]
d["reference_equivalent_tests_per_s_kernel_only"] = tests / d["duration_s"]
d["fp64_flops_per_s_kernel_only_17_per_test"] = 17 * tests / d["duration_s"]
d["bench_junction_roofline"] = bl["junction_roofline"]
json.dump(d, open('profiles/r02_junctions_ncu.json', 'w'), indent=1)
print(json.dumps({k: v for k, v in d.items() if k not in ("kernel", "command", "bench_junction_roofline")}, indent=1))
print(bl['value'], bl['e2e']['value'], bl['ms_per_step'], "stages, ms per 56 networks (one stream):",
      {k: round(v / bl.get('stages_batches', bl['steps']) / (NB / 56.0), 2) for k, v in bl['stages_ms_rank0'].items()})
print(bl['roofline']['achieved'], bl['roofline']['frac'], bl['roofline']['traffic'], bl['cpu_baseline']['value'], bl['spmv']['achieved'], bl['junction_tests_per_s'], bl['junction_roofline']['frac'])
for k, v in bl['configs'].items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ('value', 'seconds', 'sweep_ms', 'ms_per_gate_step', 'solver', 'error')})
