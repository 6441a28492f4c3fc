#!/usr/bin/env python3
"""Key counters of one ncu report + instruction / stall share per source region.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep 'augment_team_kernel.*4096' [regions.json]
"""
import csv
import io
import re
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_requests_srcunit_tex_op_red.sum",
        "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "smsp__inst_executed_op_global_red.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_alu.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_uniform.sum",
        "sm__cycles_elapsed.avg", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        try:
            if h.startswith("smsp__average_warps") and float(v) < 0.05:
                continue
        except ValueError:
            pass
        print(f"{h:75s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
open("/tmp/sass.csv", "w").write(src)
out = subprocess.run([sys.executable, "tools/ncu_lines.py", "/tmp/sass.csv", "pantas_b200/libpantas_aug.so", pat, "100000"],
                     capture_output=True, text=True).stdout
print(out.splitlines()[0])
# regions of team_tiles.cuh by marker comments
marks = []
for i, line in enumerate(open("pantas_b200/csrc/team_tiles.cuh"), 1):
    m = re.search(r"// =================\s*(.*?)\s*=================", line)
    if m:
        marks.append((i, m.group(1)[:40]))
    elif "__global__ void" in line:
        marks.append((i, "kernel prologue"))
agg = {}
lines = []
for l in out.splitlines()[2:]:
    m = re.match(r"(\S+):(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(.*)", l)
    if not m:
        continue
    f, ln, i, sm, thr = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    lines.append((f, ln, i, sm, thr, m.group(6)))
    key = f
    if f == "team_tiles.cuh":
        key = "team: helpers"
        for a, n in marks:
            if ln >= a:
                key = "team: " + n
    agg.setdefault(key, [0, 0, 0])
    agg[key][0] += i
    agg[key][1] += sm
    agg[key][2] += i * thr
print("%-50s %8s %8s %8s" % ("region", "inst%", "samp%", "thr/inst"))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-50s %8.2f %8.2f %8.1f" % (k, v[0], v[1], v[2] / max(v[0], 1e-9)))
print("top lines by samples:")
for f, ln, i, sm, thr, rest in sorted(lines, key=lambda x: -x[3])[:25]:
    print(f"  {f}:{ln:<5d} inst {i:5.2f}% samp {sm:5.2f}% thr {thr:4.1f}  {rest}")
