#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [lib.so]

Joins `ncu --page source --csv` (per-SASS-instruction executed counts) with `nvdisasm -g`
(SASS -> file:line of the -lineinfo build) by instruction order, and prints the hottest source
lines with their warp-level instruction count, share and average active threads.
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, pat = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "raytrace-miniapp_b200", "librtb200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
hdr = rows[1]
ia, ie, it = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
isamp = hdr.index("# Samples")
inst, seen = [], set()
for r in rows[2:]:
    if len(r) > it and r[ie].isdigit() and r[0] not in seen:  # ncu lists each address once per view
        seen.add(r[0])
        inst.append((r[ia].strip(), int(r[ie]), int(r[it]), int(r[isamp]) if r[isamp].isdigit() else 0))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
mangled = None
lines = []
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
    cur, loc = None, ("?", 0)
    for ln in dis:
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            lines.append((cur, loc, m.group(2).strip()))
# pick the function whose demangled name matches the kernel name best: by instruction count
byfn = collections.defaultdict(list)
for fn, loc, txt in lines:
    byfn[fn].append((loc, txt))
cands = [fn for fn, v in byfn.items() if len(v) == len(inst)]
short = re.sub(r"<.*", "", kname.split("(")[0].split("::")[-1])
cands = [c for c in cands if short in c] or cands
if not cands:
    sys.exit("no function with %d instructions found for %s" % (len(inst), kname))
fn = cands[0]
agg = collections.defaultdict(lambda: [0, 0, 0])
for (loc, txt), (src, e, t, sm) in zip(byfn[fn], inst):
    agg[loc][0] += e
    agg[loc][1] += t
    agg[loc][2] += sm
tot = sum(v[0] for v in agg.values())
tots = max(1, sum(v[2] for v in agg.values()))
print("kernel:", kname[:100])
print("function:", fn, "instructions:", len(inst), "executed warp-instr:", tot, "stall samples:", tots)
key = 2 if os.environ.get("BY") == "samples" else 0
for loc, (e, t, sm) in sorted(agg.items(), key=lambda kv: -kv[1][key])[:int(os.environ.get("TOP", "45"))]:
    print("%-28s %14d  %5.1f%%  avg threads %5.1f   samples %5.1f%%" % ("%s:%d" % loc, e, 100.0 * e / tot, t / max(e, 1), 100.0 * sm / tots))
if os.environ.get("SASS"):
    print("---- hottest SASS by samples")
    for (loc, txt), (src, e, t, sm) in sorted(zip(byfn[fn], inst), key=lambda kv: -kv[1][3])[:int(os.environ["SASS"])]:
        print("%-26s %-60s exec %11d samples %5.2f%%" % ("%s:%d" % loc, src[:60], e, 100.0 * sm / tots))
