#!/usr/bin/env python
"""Stall-reason breakdown and the most-sampled SASS instructions of one kernel in an ncu report.

    python tools/ncu_stalls.py <report.ncu-rep> <kernel-regex> [top]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[2:]:
    if len(r) > ix["stall_wait"] and r[ix["# Samples"]].isdigit() and r[0] not in seen:
        seen.add(r[0])
        data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(rows[0][1][:100], "| samples", tot, "| SASS instructions", len(data))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in reasons}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %6.1f%%" % (h, 100 * v / tot))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top_n]
for i in sorted(idx):
    r = data[i]
    top = sorted(((int(r[ix[h]]), h) for h in reasons), reverse=True)[:3]
    print("%5d %-58s smp %5.2f%% exec %10s thr %4s  %s" % (
        i, r[ix["Source"]].strip()[:58], 100 * int(r[ix["# Samples"]]) / tot, r[ix["Instructions Executed"]],
        r[ix["Avg. Threads Executed"]][:4], " ".join("%s:%d" % (h[6:], v) for v, h in top if v)))
