#!/usr/bin/env python
"""Per-kernel figures of an `ncu --set full` report as JSON (profiles/r02_traffic.json): DRAM
bytes, executed warp instructions, active threads per instruction, pipe / issue / L1-data-pipe
utilisation, duration under ncu.  Stamped with the hash of the library SOURCES the capture was
made from (raytrace_miniapp_b200.build.source_hash): bench.py quotes these static figures only
when its own sources carry the same hash.

    python tools/ncu_traffic.py <report.ncu-rep> ["note"] > profiles/r02_traffic.json
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytrace_miniapp_b200 import build  # noqa: E402

WANT = {
    "gpu__time_duration.sum": "launch_ms_under_ncu",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__inst_executed.sum": "warp_instr_per_launch",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instr",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_wavefronts_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
}
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3}

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
name = hdr.index("Kernel Name")
res = {"src_sha16": build.source_hash(),
       "source": "ncu --set full --clock-control none, B200; %s; %s" % (os.path.basename(rep),
                                                                        sys.argv[2] if len(sys.argv) > 2 else "")}
for r in data:
    k = r[name].split("(")[0].replace("void ", "").replace("rtb::", "").strip()
    k = k.split("<")[0]
    d = {}
    for m, key in WANT.items():
        if m in hdr:
            i = hdr.index(m)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if key.startswith("dram_bytes"):
                v = int(v * SCALE.get(units[i], 1.0))
            elif key == "launch_ms_under_ncu":
                v = v * SCALE.get(units[i], 1.0)
            elif key in ("warp_instr_per_launch", "registers"):
                v = int(v)
            d[key] = v
    res[k] = d
print(json.dumps(res, indent=1))
