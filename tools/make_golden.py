#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference's own inputs and its own CPU implementation.

Run in the build container (needs /root/reference and `make -C oracle ref`):
    python tools/make_golden.py
For each of ASE_small.dat / seed_small.dat the fixture holds the problem arrays, the golden
image / I_ang embedded in the .dat, the image / I_ang produced by the UNMODIFIED reference's
RayTrace::create_image(info, "cpu") in this container, and per-ray outputs of the reference's
RayTrace_calc_ray (Iv, ray2, error, RAY_DEBUG path) for a strided ray sample.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytrace_miniapp_b200 as rt  # noqa: E402
from raytrace_miniapp_b200 import problem_io  # noqa: E402
from oracle import pyoracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    pyoracle.build(ref=True)
    os.makedirs(OUT, exist_ok=True)
    for name, n_sample in (("ASE_small", 1500), ("seed_small", 800)):
        path = os.path.join(pyoracle.REF_ROOT, name + ".dat")
        p, gimg, gang = rt.read_dat(path)
        R = pyoracle.Reference(path)
        img, ang, sec = R.create_image("cpu")
        rays = p.rays()
        idx = np.linspace(0, rays.size - 1, n_sample).astype(np.int64)
        pr = R.calc_rays(rays[idx], p.method, debug=False)
        # the RAY_DEBUG trajectory comes from a second call: with debug != NULL the reference
        # switches to the emission-style integration (RayTraceImageHelper.h:543), so its Iv differ
        pr["debug"] = R.calc_rays(rays[idx], p.method, debug=True)["debug"]
        problem_io.save_npz(os.path.join(OUT, name.lower() + ".npz"), p,
                            dat_golden_image=gimg, dat_golden_I_ang=gang,
                            ref_cpu_image=img, ref_cpu_I_ang=ang,
                            sample_index=idx, sample_Iv=pr["Iv"],
                            sample_ray2=pr["ray2"].view(np.float32).reshape(-1, 4),
                            sample_error=pr["error"], sample_debug=pr["debug"])
        print("%s: reference cpu %.2f s, |image| %.17g, |I_ang| %.17g" %
              (name, sec, np.linalg.norm(img), np.linalg.norm(ang)))
        R.close()


if __name__ == "__main__":
    main()
