#!/usr/bin/env python
"""Kernel times of ONE device's share of the row-cyclic decomposition over W devices, measured
on a single GPU (the share of rank r is what that rank's GPU runs in a W-GPU job): the strong-
scaling curve of the kernels without needing W GPUs.

    python tools/time_share.py [W ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytrace_miniapp_b200 import lib, problem_io, synth  # noqa: E402

small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
p = synth.ase_medium_synth(small)
e = p.euv_beam
ctx = lib.Context(0)
ctx.stage(p)
dev = torch.device("cuda", 0)
image = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
I_ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
base = None
for W in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    res = []
    for r in sorted({0, W - 1}):
        best = None
        for it in range(6):
            flush.zero_()
            ctx.reset_timings()
            ctx.launch_rows(r, W, image, I_ang)
            ctx.sync()
            t = ctx.timings()
            if it >= 2 and (best is None or t["march_ms"] + t["integrate_ms"] < best[0] + best[1]):
                best = (t["march_ms"], t["integrate_ms"])
        res.append((r, best))
    worst = max(res, key=lambda x: x[1][0] + x[1][1])
    tot = worst[1][0] + worst[1][1]
    if W == 1:
        base = tot
    print("W=%d  slowest share r=%d: march %.3f ms integrate %.3f ms total %.3f ms%s"
          % (W, worst[0], worst[1][0], worst[1][1], tot,
             "  kernel speed-up %.2fx (%.0f%%)" % (base / tot, 100 * base / tot / W) if base else ""))
