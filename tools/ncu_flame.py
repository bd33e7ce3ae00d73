#!/usr/bin/env python
"""Inclusive / exclusive cost per source line of one kernel, from an ncu report plus the binary it profiled.

    ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_flame.py sass.csv gnxraytracer_b200/lib/libgnxrt.so --kernel k_shade --launch 0

ncu's per-SASS-instruction counters are exclusive; `nvdisasm -gi` of the same binary gives the inline
chain of every instruction (innermost frame first).  Joining the two by instruction offset yields, for each
source line, the cost of everything inlined beneath it — the view needed to see which call site of a
force-inlined function the time goes to.
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import tempfile

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("lib")
ap.add_argument("--kernel", required=True, help="substring of the (demangled) kernel name in the ncu csv")
ap.add_argument("--mangled", default="", help="substring of the mangled section name (default: derived)")
ap.add_argument("--launch", type=int, default=0, help="which matching launch of the report")
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--file", default="", help="only list inclusive lines of this source file")
a = ap.parse_args()


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


# ---- ncu: per-instruction counters of the chosen launch
csv.field_size_limit(1 << 30)
launches = []
cur = None
hdr = None
with open(a.csv, newline="") as f:
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] in ("Function Name", "Kernel Name"):
            cur = {"name": row[1], "rows": []}
            launches.append(cur)
            continue
        if row[0] == "Address":
            hdr = {h: i for i, h in enumerate(row)}
            continue
        if cur is None or hdr is None or not row[0].startswith("0x"):
            continue
        cur["rows"].append((int(row[0], 16), row[hdr["Source"]], num(row[hdr["Instructions Executed"]]),
                            num(row[hdr["Thread Instructions Executed"]]), num(row[hdr["# Samples"]])))
match = [l for l in launches if a.kernel in l["name"]]
if not match:
    raise SystemExit("kernel not found; have: " + ", ".join(sorted({l["name"][:50] for l in launches})))
L = match[a.launch]
rows = L["rows"]
base = min(r[0] for r in rows)

# ---- nvdisasm: inline chain per instruction offset
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout
m = re.search(r"(\w+)<\(int\)(\d+)>", L["name"])
want = a.mangled or (f"{m.group(1)}ILi{m.group(2)}E" if m else re.sub(r"\W.*", "", L["name"].split("::")[-1]))
chains = {}
sect = None
chain = []
pending = []
for line in dis.splitlines():
    s = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if s:
        sect = s.group(1)
        continue
    if sect is None or want not in sect:
        continue
    fl = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if fl:
        pending.append((fl.group(1).split("/")[-1], int(fl.group(2))))
        continue
    ins = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*)", line)
    if ins and not ins.group(2).startswith("."):
        if pending:
            chain = pending
            pending = []
        chains[int(ins.group(1), 16)] = chain
if not chains:
    raise SystemExit(f"no section matching {want}")

excl = collections.defaultdict(lambda: [0, 0, 0])
incl = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
missing = 0
for addr, sass, ie, te, smp in rows:
    ch = chains.get(addr - base)
    if ch is None:
        missing += 1
        ch = [("?", 0)]
    for j in range(3):
        tot[j] += (ie, te, smp)[j]
    if ch:
        e = excl[ch[0]]
        e[0] += ie; e[1] += te; e[2] += smp
        for fr in set(ch):
            v = incl[fr]
            v[0] += ie; v[1] += te; v[2] += smp

src_cache = {}


def text(fr):
    p = os.path.join(os.path.dirname(os.path.abspath(a.lib)), "..", "csrc", fr[0])
    if fr[0] not in src_cache:
        try:
            src_cache[fr[0]] = open(p).read().splitlines()
        except OSError:
            src_cache[fr[0]] = []
    ls = src_cache[fr[0]]
    return ls[fr[1] - 1].strip()[:95] if 0 < fr[1] <= len(ls) else ""


print(f"kernel {L['name'][:80]}\n  {len(rows)} SASS instr, warp-instr {tot[0]:,}, avg lanes {tot[1] / max(tot[0], 1):.1f}, samples {tot[2]:,}, unmatched {missing}")
print("--- inclusive (everything inlined beneath the line)")
items = [(k, v) for k, v in incl.items() if not a.file or k[0] == a.file]
for k, v in sorted(items, key=lambda kv: -kv[1][0])[: a.top]:
    print(f"  {k[0]}:{k[1]:<4d} inst {100 * v[0] / tot[0]:5.1f}% lanes {v[1] / max(v[0], 1):4.1f} smp {100 * v[2] / max(tot[2], 1):5.1f}%  {text(k)}")
print("--- exclusive (innermost frame)")
for k, v in sorted(excl.items(), key=lambda kv: -kv[1][0])[: a.top]:
    print(f"  {k[0]}:{k[1]:<4d} inst {100 * v[0] / tot[0]:5.1f}% lanes {v[1] / max(v[0], 1):4.1f} smp {100 * v[2] / max(tot[2], 1):5.1f}%  {text(k)}")
