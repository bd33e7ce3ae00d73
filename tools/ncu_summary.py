#!/usr/bin/env python
"""Condense the per-launch ncu metric lists (tools/final_evidence.sh, step 2) into profiles/r02_trace_metrics.json — the
lane / issue / DRAM figures bench.py attaches to its roofline object — and print a per-kernel table.
usage: python tools/ncu_summary.py [dir with r02_dram_trace_<workload>.csv, default profiles/]"""
import collections, csv, io, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")
out = {}
for w in ("c2", "u1p", "c4"):
    p = os.path.join(src, f"r02_dram_trace_{w}.csv")
    if not os.path.exists(p):
        continue
    txt = open(p).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    launches = collections.OrderedDict()
    for r in rows:
        launches.setdefault((r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace(", 0>", ">")), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    agg = collections.OrderedDict()
    for (_, name), m in launches.items():
        a = agg.setdefault(name, collections.defaultdict(float))
        t, inst = m["gpu__time_duration.sum"], m["sm__inst_executed.sum"]
        a["launches"] += 1
        a["ns"] += t
        a["dram_bytes"] += m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        a["warp_inst"] += inst
        a["_lanes"] += m["smsp__thread_inst_executed_per_inst_executed.ratio"] * inst
        a["_issue"] += m["smsp__issue_active.avg.pct_of_peak_sustained_active"] * t
        a["_occ"] += m["sm__warps_active.avg.pct_of_peak_sustained_active"] * t
        a["_l1"] += m["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"] * t
    out[w] = {}
    print(f"== {w} (one step under ncu: serialised, cold caches — shares, not absolutes)")
    for name, a in agg.items():
        rec = {"launches": int(a["launches"]), "ms_under_ncu": a["ns"] / 1e6, "dram_gb": a["dram_bytes"] / 1e9,
               "dram_gbs": a["dram_bytes"] / a["ns"], "warp_inst": int(a["warp_inst"]),
               "active_lanes_per_inst": a["_lanes"] / max(a["warp_inst"], 1), "issue_slot_pct": a["_issue"] / a["ns"],
               "warps_active_pct": a["_occ"] / a["ns"], "l1_wavefront_pct": a["_l1"] / a["ns"]}
        out[w][name] = rec
        print("  %-22s n=%4d %8.2f ms  dram %7.2f GB (%5.0f GB/s)  lanes %4.1f  issue %4.1f %%  warps %4.1f %%  l1 %4.1f %%" % (
            name, rec["launches"], rec["ms_under_ncu"], rec["dram_gb"], rec["dram_gbs"], rec["active_lanes_per_inst"],
            rec["issue_slot_pct"], rec["warps_active_pct"], rec["l1_wavefront_pct"]))
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_trace_metrics.json"), "w"), indent=1)
