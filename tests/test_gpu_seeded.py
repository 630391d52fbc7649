"""Seeded (method 2, forward) path on the GPU: scatter binning by the exit ray, separable seed
tabulated per grid index, explicit ray lists with the seed interpolated on the device
(src/common/RayTraceImageHelper.h:168-247, :523-533, :569-581; src/RayTraceImageCPU.cpp:40-68)."""
import numpy as np
import pytest

from raytrace_miniapp_b200 import abi
from conftest import max_rel, rel_l2

pytestmark = pytest.mark.gpu


def test_strided_worker_matches_oracle(seed_small, oracle, ctx):
    p, _ = seed_small
    p.N_start, p.N_parallel = 3, 97
    try:
        img, ang = ctx.create_image(p)
        o = oracle.create_image(p)
    finally:
        p.N_start, p.N_parallel = 0, 1
    assert o["failure_code"] == 0 and ctx.failure_code == 0
    assert rel_l2(img, o["image"]) <= 1e-10 and rel_l2(ang, o["I_ang"]) <= 1e-10
    assert max_rel(img, o["image"]) <= 1e-9 and max_rel(ang, o["I_ang"]) <= 1e-9
    assert np.linalg.norm(img) > 0


def test_explicit_ray_list_with_seed(seed_small, oracle, ctx):
    """RayTraceImage<B200>Loop with a seed: the seed amplitude is interpolated per ray on the
    device (interp_pchip) instead of being tabulated per grid index."""
    p, _ = seed_small
    e, s = p.euv_beam, p.seed_beam
    rays = p.rays()[11::193]
    scale = (s.dx * s.dy * s.da * s.db) / (e.dx * e.dy)
    o = oracle.trace_rays(p, rays, 2, scale)
    img, ang = ctx.trace_rays(p, rays, 2, scale)
    assert rel_l2(img, o["image"]) <= 1e-10 and rel_l2(ang, o["I_ang"]) <= 1e-10
    assert np.linalg.norm(o["image"]) > 0


def test_per_ray_spectra_with_seed(seed_small, oracle, ctx):
    """rtb200_calc_rays (RayTrace::calc_ray) in seeded mode: Iv, exit ray, error per ray."""
    p, _ = seed_small
    rays = p.rays()[5::2503]
    g = ctx.calc_rays(p, rays)
    o = oracle.calc_rays(p, rays)
    assert np.array_equal(g["error"], o["error"])
    ok = o["error"] == 0
    for f in "xyab":
        assert np.array_equal(g["ray2"][f][ok].view(np.uint32), o["ray2"][f][ok].view(np.uint32)), f
    scale = np.abs(o["Iv"]).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(g["Iv"] - o["Iv"]) / scale) < 1e-12
    assert (o["Iv"] > 0).any()


def test_backward_method_with_seed(seed_small, oracle, ctx):
    """method 1 with a seed (reachable through calc_ray only): the seed is evaluated at the
    EXIT ray (:525-529)."""
    p, _ = seed_small
    rays = p.rays()[7::9001]
    # rays that start on the euv grid and run backward
    g = ctx.calc_rays(p, rays, method=1)
    o = oracle.calc_rays(p, rays, method=1)
    assert np.array_equal(g["error"], o["error"])
    scale = np.abs(o["Iv"]).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(g["Iv"] - o["Iv"]) / scale) < 1e-11
