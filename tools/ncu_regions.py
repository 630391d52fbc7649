#!/usr/bin/env python
"""SASS-order view of one kernel from an ncu report: consecutive instruction windows with their
executed count, average active threads and the source lines they come from (outer lines only).

    python tools/ncu_regions.py <report.ncu-rep> <kernel-regex> <mangled-substring> [window]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, pat, sub = sys.argv[1], sys.argv[2], sys.argv[3]
win = int(sys.argv[4]) if len(sys.argv) > 4 else 24
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "raytrace-miniapp_b200", "librtb200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ie, it = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
isamp = hdr.index("# Samples")
seen, inst = set(), []
for r in rows[2:]:
    if len(r) > it and r[ie].isdigit() and r[0] not in seen:
        seen.add(r[0])
        inst.append((int(r[ie]), int(r[it]), int(r[isamp]) if r[isamp].isdigit() else 0))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
fns = collections.defaultdict(list)
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    cur, loc = None, ("?", 0)
    for ln in subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            loc = (os.path.basename(m.group(1)).replace("rtb200_", ""), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            fns[cur].append((loc, m.group(2).strip()))
fn = [f for f in fns if sub in f and len(fns[f]) == len(inst)][0]
L = fns[fn]
tot = sum(i[0] for i in inst)
tots = max(1, sum(i[2] for i in inst))
print(fn, len(L), "warp-instr", tot)
for a in range(0, len(L), win):
    b = min(a + win, len(L))
    e = sum(i[0] for i in inst[a:b])
    t = sum(i[1] for i in inst[a:b])
    s = sum(i[2] for i in inst[a:b])
    srcs = collections.Counter("%s:%d" % l[0] for l in L[a:b] if "math" not in l[0][0] and "intrin" not in l[0][0] and ".hpp" not in l[0][0])
    print("%5d exec %7.2fM share %4.1f%% thr %5.1f samp %4.1f%%  %s" % (a, e / (b - a) / 1e6, 100.0 * e / tot, t / max(e, 1), 100.0 * s / tots,
                                                      " ".join(k for k, _ in srcs.most_common(6))))
