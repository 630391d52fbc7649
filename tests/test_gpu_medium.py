"""BASELINE.json configs[1]: the ASE_medium stand-in on one B200, validated against the CPU
oracle (restatement of RayTraceImageCPU, run multi-threaded) at full size, plus size-independent
properties of the path at that size."""
import os

import numpy as np
import pytest

from raytrace_miniapp_b200 import synth
from conftest import max_rel, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def medium(ase_small):
    return synth.ase_medium_synth(ase_small[0])


def test_ase_medium_synth_matches_cpu_oracle(medium, oracle, ctx):
    p = medium
    assert p.n_rays == 2994600 and p.N == 6 and p.euv_beam.nv == 52
    img, ang = ctx.create_image(p)
    o = oracle.create_image(p, threads=max(2, os.cpu_count() or 2))
    assert o["failure_code"] == 0 and ctx.failure_code == 0
    assert rel_l2(img, o["image"]) <= 1e-10 and rel_l2(ang, o["I_ang"]) <= 1e-10
    assert max_rel(img, o["image"]) <= 1e-9 and max_rel(ang, o["I_ang"]) <= 1e-9


HAVE_REF = os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       "oracle", "_ref", "libref_oracle.so"))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref (the reference built from its own sources) not present")
def test_ase_medium_synth_matches_the_live_reference(medium, ctx, tmp_path):
    """The same full-size image against the UNMODIFIED reference itself: the stand-in is written
    in the reference's wire format, loaded by the reference's own unpack and traced by its own
    `threads` method (RayTraceImageCPULoop on all host cores) in this process."""
    import raytrace_miniapp_b200 as rt
    from oracle import pyoracle
    f = str(tmp_path / "ASE_medium_synth.dat")
    rt.write_dat(f, medium)
    R = pyoracle.Reference(f)
    ref_img, ref_ang, _ = R.create_image("threads")
    R.close()
    img, ang = ctx.create_image(medium)
    assert np.linalg.norm(ref_img) > 0
    assert rel_l2(img, ref_img) <= 1e-10 and rel_l2(ang, ref_ang) <= 1e-10
    assert max_rel(img, ref_img) <= 1e-9 and max_rel(ang, ref_ang) <= 1e-9


@pytest.fixture(scope="module")
def seed_medium(seed_small):
    return synth.seed_medium_synth(seed_small[0])


def test_seed_medium_synth_sample_and_full_size_properties(seed_medium, oracle, ctx):
    """BASELINE.json configs[1], seeded half: the seed_medium stand-in (60 993 450 rays, N = 6).
    (1) every 32-nd ray against the CPU oracle; (2) at FULL size, properties that do not need the
    oracle: the image is linear in the seed amplitude (x2 is exact), and the strided workers of
    the reference's own N_start / N_parallel decomposition add up to the full image."""
    p = seed_medium
    assert p.n_rays == 60993450 and p.N == 6 and p.euv_beam.nv == 82
    p.N_start, p.N_parallel = 3, 32
    try:
        img_s, ang_s = ctx.create_image(p)
        o = oracle.create_image(p, threads=max(2, os.cpu_count() or 2))
    finally:
        p.N_start, p.N_parallel = 0, 1
    assert o["failure_code"] == 0 and ctx.failure_code == 0 and np.linalg.norm(o["image"]) > 0
    assert rel_l2(img_s, o["image"]) <= 1e-10 and rel_l2(ang_s, o["I_ang"]) <= 1e-10
    assert max_rel(img_s, o["image"]) <= 1e-9 and max_rel(ang_s, o["I_ang"]) <= 1e-9
    full, full_ang = ctx.create_image(p)
    f0 = p.seed.f0
    try:
        p.seed.f0 = 2.0 * f0
        twice, _ = ctx.create_image(p)
    finally:
        p.seed.f0 = f0
    assert rel_l2(twice, 2.0 * full) < 1e-13  # atomics: summation order differs between runs
    parts = np.zeros_like(full)
    parts_ang = np.zeros_like(full_ang)
    for k in range(2):
        p.N_start, p.N_parallel = k, 2
        try:
            a, b = ctx.create_image(p)
        finally:
            p.N_start, p.N_parallel = 0, 1
        parts += a
        parts_ang += b
    assert rel_l2(parts, full) < 1e-13 and rel_l2(parts_ang, full_ang) < 1e-13


def test_emission_linearity_is_exact(medium, ctx):
    """I is linear in the emissivity and scaling E0 by 2 is exact in binary floating point, so
    image(2*E0) == 2*image(E0) bit for bit at any size."""
    p = medium
    img1, ang1 = ctx.create_image(p)
    saved = [g.E0.copy() for g in p.gain]
    try:
        for g in p.gain:
            g.E0 *= 2.0
        img2, ang2 = ctx.create_image(p)
    finally:
        for g, s in zip(p.gain, saved):
            g.E0[:] = s
    assert np.array_equal(img2, 2.0 * img1)
    assert rel_l2(ang2, 2.0 * ang1) < 1e-14  # I_ang is summed with atomics: order may differ


def test_repeatability_and_mirror_symmetry(medium, ctx):
    """Two runs give the identical image (owner kernel: no atomics on the image)."""
    a, _ = ctx.create_image(medium)
    b, _ = ctx.create_image(medium)
    assert np.array_equal(a, b) and np.isfinite(a).all() and (a >= 0).all()
