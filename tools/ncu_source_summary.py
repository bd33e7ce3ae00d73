#!/usr/bin/env python
"""Aggregates `ncu --page source --csv --print-source sass,cuda` output per kernel / source file / source line.

    ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda > src.csv
    python tools/ncu_source_summary.py src.csv [--kernel k_shade] [--top 40]

For each kernel: warp-level instructions executed, average active lanes and stall samples, summed over the
SASS rows that follow each CUDA source row (a SASS row belongs to the source row printed before it).
"""
import argparse
import collections
import csv
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--kernel", default="")
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()

csv.field_size_limit(1 << 30)


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


kern = None
path = None
hdr = None
cur_line = None
agg = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0, 0, ""]))  # kernel -> (file,line) -> [inst, thread, samples, nsass, text]
with open(a.csv, newline="") as f:
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] == "File Path":
            path = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            kern = row[1]
            continue
        if row[0] == "Line No":
            hdr = {h: i for i, h in enumerate(row)}
            # two columns are called "Source": the first is CUDA text, the second SASS
            continue
        if hdr is None:
            continue
        if row[0] != "":
            cur_line = (path, int(row[0]))
            agg[kern][cur_line][4] = row[1].strip()[:90]
            continue
        if cur_line is None:
            continue
        rec = agg[kern][cur_line]
        rec[0] += num(row[hdr["Instructions Executed"]])
        rec[1] += num(row[hdr["Thread Instructions Executed"]])
        rec[2] += num(row[hdr["# Samples"]])
        rec[3] += 1

for k, lines in agg.items():
    if a.kernel and a.kernel not in k:
        continue
    tot_i = sum(v[0] for v in lines.values())
    tot_t = sum(v[1] for v in lines.values())
    tot_s = sum(v[2] for v in lines.values())
    tot_n = sum(v[3] for v in lines.values())
    if tot_i == 0:
        continue
    print(f"=== {k[:100]}")
    print(f"    SASS instructions {tot_n}, warp-instr executed {tot_i:,}, avg lanes {tot_t / tot_i:.1f}, samples {tot_s:,}")
    per_file = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (p, ln), v in lines.items():
        for j in range(4):
            per_file[p][j] += v[j]
    for p, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"    {p:24s} sass {v[3]:6d}  inst {100 * v[0] / tot_i:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / max(tot_s, 1):5.1f}%")
    print("    --- top lines by warp-instructions")
    for (p, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print(f"    {p}:{ln:<5d} sass {v[3]:4d} inst {100 * v[0] / tot_i:5.1f}% lanes {v[1] / max(v[0], 1):5.1f} smp {100 * v[2] / max(tot_s, 1):5.1f}%  {v[4]}")
