#!/usr/bin/env python
"""Wall time of rtb200_multi_create_image (every device of the box behind one call, host buffers
in and out) on the fixture problems, best of 7, with parity against the single-device result."""
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else lib.load().rtb200_device_count()
one, many = lib.Context(0), lib.MultiContext(n)
for name, p in cases:
    ref_img, ref_ang = one.create_image(p)
    many.create_image(p)
    best = 1e9
    for _ in range(7):
        t0 = time.perf_counter()
        img, ang = many.create_image(p)
        best = min(best, time.perf_counter() - t0)
    e_img = np.linalg.norm(img - ref_img) / np.linalg.norm(ref_img)
    e_ang = np.linalg.norm(ang - ref_ang) / np.linalg.norm(ref_ang)
    print("%-18s %d devices: wall %8.3f ms | vs one device: image relL2 %.2e I_ang relL2 %.2e"
          % (name, n, best * 1e3, e_img, e_ang))
