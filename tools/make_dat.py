#!/usr/bin/env python
"""Writes reference-format .dat inputs from the committed fixtures / synthetic generators.

    python tools/make_dat.py <out_dir> [ase_small] [seed_small] [ase_medium_synth]
The files are what the reference's CreateImage driver (and oracle/_ref/CreateImageB200) reads.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytrace_miniapp_b200 import problem_io, synth, write_dat  # noqa: E402

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
names = sys.argv[2:] or ["ase_small", "seed_small"]
for name in names:
    if name == "ase_medium_synth":
        small, _ = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", "ase_small.npz"))
        write_dat(os.path.join(out, "ASE_medium_synth.dat"), synth.ase_medium_synth(small))
    else:
        p, extra = problem_io.load_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        write_dat(os.path.join(out, {"ase_small": "ASE_small", "seed_small": "seed_small"}[name] + ".dat"),
                  p, extra["dat_golden_image"], extra["dat_golden_I_ang"])
print(sorted(os.listdir(out)))
