#!/usr/bin/env python
"""Times rtb200_create_image (host buffers, H2D + kernels + D2H) on the fixture problems."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytrace_miniapp_b200 import lib, problem_io, synth  # noqa: E402

small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
seed, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "seed_small.npz"))
cases = [("ASE_small", small), ("seed_small", seed), ("ASE_medium-synth", synth.ase_medium_synth(small))]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
ctx = lib.Context(0)
for name, p in cases:
    ctx.create_image(p)
    best, t = 1e9, None
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.create_image(p)
        dt = time.perf_counter() - t0
        if dt < best:
            best, t = dt, ctx.timings()
    print("%-18s rays %9d  wall %8.3f ms | h2d %.3f march %.3f integrate %.3f d2h %.3f ms | launches %d | %.3e ray-seg/s"
          % (name, p.n_rays, best * 1e3, t["h2d_ms"], t["march_ms"], t["integrate_ms"], t["d2h_ms"],
             t["kernel_launches"], p.ray_segments / best))
