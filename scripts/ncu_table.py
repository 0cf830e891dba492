#!/usr/bin/env python
"""Condense scripts/ncu_summary.py output into one line per launch + totals per kernel name."""
import re, sys, collections
txt = open(sys.argv[1]).read().split("\n== ")
rows = []
for blk in txt:
    lines = blk.strip().splitlines()
    if not lines: continue
    name = lines[0].replace("== ", "")
    name = re.sub(r"^ccd::", "", name)
    short = name.split("(")[0]
    d = {}
    for l in lines[1:]:
        p = l.split()
        if len(p) >= 2 and not l.strip().startswith("stalls"):
            try: d[p[0]] = (float(p[1]), p[2] if len(p) > 2 else "")
            except ValueError: pass
        if l.strip().startswith("stalls"): d["stalls"] = l.strip()[14:]
    def g(k, scale=None):
        v, u = d.get(k, (0, ""))
        if scale == "bytes":
            v *= {"Gbyte": 1e3, "Mbyte": 1, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1)
        if scale == "ms":
            v *= {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u, 1)
        return v
    rows.append((short, g("gpu__time_duration.sum", "ms"), g("smsp__thread_inst_executed_per_inst_executed.ratio"), g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                 g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("dram__bytes_read.sum", "bytes") + g("dram__bytes_write.sum", "bytes"),
                 g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g("launch__registers_per_thread"), d.get("stalls", "")))
print("%-44s %8s %6s %6s %6s %6s %9s %6s %4s" % ("kernel", "ms", "lanes", "fp64%", "issue%", "occ%", "dramMB", "dram%", "regs"))
tot = collections.OrderedDict()
for r in rows:
    print("%-44s %8.3f %6.1f %6.1f %6.1f %6.1f %9.1f %6.1f %4d  %s" % (r[0][:44], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9][:70]))
    t = tot.setdefault(r[0], [0, 0.0, 0.0]); t[0] += 1; t[1] += r[1]; t[2] += r[6]
print("\nTOTAL per kernel (ncu serialised, cold cache):")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-44s x%-3d %8.3f ms %9.1f MB" % (k[:44], v[0], v[1], v[2]))
print("sum %.3f ms, %.1f MB" % (sum(v[1] for v in tot.values()), sum(v[2] for v in tot.values())))
