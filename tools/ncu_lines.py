#!/usr/bin/env python3
"""Attribute an ncu SASS-level source page to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    python tools/ncu_lines.py sass.csv pantas_b200/libpantas_aug.so 'augment_tiles_kernelILi256' [top]

ncu's CLI prints per-instruction metrics only for SASS; this joins them, by
instruction order, with `nvdisasm -g` line info of the same kernel in the built
library (compile with -lineinfo), and prints instructions executed and stall
samples per source line (innermost inlined location).
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(so, kernel_pat):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, capture_output=True)
        cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    out = []
    inside = False
    cur = ("?", 0)
    for line in txt.splitlines():
        if line.startswith("//---") and ".text." in line:
            inside = re.search(kernel_pat, line) is not None
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((cur, m.group(2).strip()))
    return out


def main():
    sass_csv, so, pat = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    col = {n: i for i, n in enumerate(hdr)}
    lines = sass_lines(so, pat)
    if len(lines) != len(body):
        print(f"warning: {len(lines)} disassembled instructions vs {len(body)} profiled", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot_inst = tot_samp = 0
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for (loc, text), r in zip(lines, body):
        inst = int(r[col["Instructions Executed"]])
        thr = int(r[col["Thread Instructions Executed"]])
        samp = int(r[col["# Samples"]])
        a = agg[loc]
        a[0] += inst
        a[1] += thr
        a[2] += samp
        for sname in stalls:
            v = int(r[col[sname]] or 0)
            if v:
                a[3][sname] += v
        tot_inst += inst
        tot_samp += samp
    print(f"total warp instructions {tot_inst}, samples {tot_samp}")
    print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'thr/inst':>8s}  top stalls")
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        st = ", ".join(f"{k[6:]}={v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
        print(f"{loc[0] + ':' + str(loc[1]):28s} {100 * a[0] / max(tot_inst, 1):6.2f} {100 * a[2] / max(tot_samp, 1):6.2f} "
              f"{a[1] / max(a[0], 1):8.1f}  {st}")


if __name__ == "__main__":
    main()
