"""Randomised parity sweep (fixed seeds): each case perturbs ASE_small along several axes at once
- stronger / weaker refraction, non-uniform gain grids, number of planes, spectral resolution,
beam refinement, dz, strided decomposition - and checks the CUDA path against the CPU oracle:
march intermediates bit for bit on a ray sample, image / I_ang to 1e-10."""
import os

import numpy as np
import pytest

from raytrace_miniapp_b200 import abi, synth
from conftest import max_rel, rel_l2

pytestmark = pytest.mark.gpu

# RTB200_FUZZ=<n> widens the sweep (n ASE and n/2 seeded configurations) for a one-off hunt.
N_FUZZ = int(os.environ.get("RTB200_FUZZ", "16"))


def _case(small, seed):
    rng = np.random.default_rng(seed)
    p = small
    K = int(rng.choice([17, 32, 52, 64, 77, 130, 200]))
    if K != p.euv_beam.nv:
        p = synth.spectral(p, K)
    euv = synth.scale_beam(p.euv_beam, float(rng.choice([0.6, 0.8, 1.0, 1.25])))
    if rng.random() < 0.5:
        euv = abi.BeamGrid(euv.x, euv.y, euv.a, euv.b, euv.dx, euv.dy, euv.da, euv.db, dv=euv.dv,
                           dz=euv.dz * float(rng.choice([0.5, 2.0])), extra={})
    alpha = float(rng.choice([0.25, 1.0, 3.0]))  # refraction strength: n -> 1 - alpha*(1 - n)
    gscale = np.float32(rng.choice([0.3, 1.0, 2.5]))  # gain strength (moves the Taylor/exp split)
    ax, ay = float(rng.uniform(-0.7, 0.7)), float(rng.uniform(-0.7, 0.7))
    warp = rng.random() < 0.6

    def tweak(g):
        g = abi.Gain(g.x, g.y, 1.0 - alpha * (1.0 - g.n), g.g0 * gscale,
                     None if g.E0 is None else g.E0 * gscale, g.gv, g.gv0)
        return synth.warp_gain_grid(g, ax, ay) if warp else g

    g = [tweak(q) for q in p.gain]
    n_planes = int(rng.integers(2, 8))
    planes = [g[0]] + [synth.blend_planes(g[1], g[2], w) for w in np.linspace(0, 1, n_planes - 1)]
    out = abi.Problem(euv, planes)
    out.N_parallel = int(rng.integers(5, 23))
    out.N_start = int(rng.integers(0, out.N_parallel))
    return out


@pytest.mark.parametrize("seed", range(N_FUZZ))
def test_random_configuration(seed, ase_small, oracle, ctx):
    p = _case(ase_small[0], seed)
    rays = p.rays()
    sample = rays[:: max(1, rays.size // 1500)]
    g = ctx.calc_rays(p, sample)
    o = oracle.calc_rays(p, sample)
    assert np.array_equal(g["error"], o["error"])
    for f in ("gvl", "evl"):
        assert np.array_equal(g[f].view(np.uint32), o[f].view(np.uint32)), (seed, f)
    assert np.array_equal(g["ivl"], o["ivl"])
    img, ang = ctx.create_image(p, flags=abi.FLAG_NO_LIMITS)
    oi = oracle.create_image(p, flags=abi.FLAG_NO_LIMITS)
    assert oi["rc"] in (abi.OK, abi.RAYS_FAILED) and ctx.failure_code == oi["failure_code"]
    if np.linalg.norm(oi["image"]) > 0:
        assert rel_l2(img, oi["image"]) <= 1e-10 and max_rel(img, oi["image"]) <= 1e-9, seed
    if np.linalg.norm(oi["I_ang"]) > 0:
        assert rel_l2(ang, oi["I_ang"]) <= 1e-10 and max_rel(ang, oi["I_ang"]) <= 1e-9, seed
    # the same rays as an explicit list (RayTraceImage<Backend>Loop), accumulated on top of a
    # non-zero image: scatter binning instead of pixel ownership
    base_img = np.full_like(oi["image"], 0.25 * np.abs(oi["image"]).max())
    base_ang = np.zeros_like(oi["I_ang"])
    gi, ga = base_img.copy(), base_ang.copy()
    wi, wa = base_img.copy(), base_ang.copy()
    ctx.trace_rays(p, rays, 1, 0.5, gi, ga)
    oracle.trace_rays(p, rays, 1, 0.5, wi, wa)
    assert rel_l2(gi, wi) <= 1e-10 and rel_l2(ga, wa + 1e-300) <= 1e-10, seed


@pytest.mark.parametrize("seed", range(max(10, N_FUZZ // 2)))
def test_random_seeded_configuration(seed, seed_small, oracle, ctx):
    """The seeded (gain-only, scatter-binned) path under the same kind of perturbation.  (Seed 1
    is the case that exposed rays leaving the image leaking into the next pixel of their run.)"""
    rng = np.random.default_rng(100 + seed)
    p0 = seed_small[0]
    alpha = float(rng.choice([0.3, 1.0, 2.5]))
    gscale = np.float32(rng.choice([0.4, 1.0, 2.0]))
    ax, ay = float(rng.uniform(-0.6, 0.6)), float(rng.uniform(-0.6, 0.6))
    warp = rng.random() < 0.5

    def tweak(g):
        g = abi.Gain(g.x, g.y, 1.0 - alpha * (1.0 - g.n), g.g0 * gscale,
                     None if g.E0 is None else g.E0 * gscale, g.gv, g.gv0)
        return synth.warp_gain_grid(g, ax, ay) if warp else g

    p = abi.Problem(p0.euv_beam, [tweak(g) for g in p0.gain], p0.seed_beam, p0.seed)
    p.N_parallel = int(rng.integers(150, 400))
    p.N_start = int(rng.integers(0, p.N_parallel))
    img, ang = ctx.create_image(p)
    oi = oracle.create_image(p)
    assert ctx.failure_code == oi["failure_code"]
    assert np.linalg.norm(oi["image"]) > 0
    assert rel_l2(img, oi["image"]) <= 1e-10 and max_rel(img, oi["image"]) <= 1e-9, seed
    assert rel_l2(ang, oi["I_ang"]) <= 1e-10 and max_rel(ang, oi["I_ang"]) <= 1e-9, seed
    # explicit ray lists through the same kernels: forward (seed at the entry point, binned by the
    # exit ray) and backward (seed at the exit point, binned by the ray itself)
    rays = p.rays()[:: 7]
    for method in (2, 1):
        gi, ga = np.zeros_like(img), np.zeros_like(ang)
        wi, wa = np.zeros_like(img), np.zeros_like(ang)
        ctx.trace_rays(p, rays, method, 2.0, gi, ga)
        oracle.trace_rays(p, rays, method, 2.0, wi, wa)
        if np.linalg.norm(wi) > 0:
            assert rel_l2(gi, wi) <= 1e-10 and max_rel(gi, wi) <= 1e-9, (seed, method)
        if np.linalg.norm(wa) > 0:
            assert rel_l2(ga, wa) <= 1e-10 and max_rel(ga, wa) <= 1e-9, (seed, method)


def _unmirrored(p):
    """The same plasma with the y < 0 half spelled out (gain planes and beam grid symmetric about
    y = 0) instead of mirrored by |y|: the abs_y = 0 branches of the march and of the binning."""
    def full(g):
        assert g.y[0] == 0.0
        y = np.concatenate([-g.y[:0:-1], g.y])
        m = lambda a: None if a is None else np.concatenate([a[:0:-1], a], axis=0)  # noqa: E731
        return abi.Gain(g.x, y, m(g.n), m(g.g0), m(g.E0), m(g.gv), m(g.gv0))
    e = p.euv_beam
    euv = abi.BeamGrid(e.x, np.concatenate([-e.y[::-1], e.y]), e.a, e.b, e.dx, e.dy, e.da, e.db,
                       dv=e.dv, dz=e.dz, extra={})
    return abi.Problem(euv, [full(g) for g in p.gain], None, None, p.N_start, p.N_parallel)


def test_unmirrored_plasma(ase_small, oracle, ctx):
    p = _unmirrored(ase_small[0])
    assert p.gain[1].y[0] < 0 and p.euv_beam.y[0] < 0
    p.N_start, p.N_parallel = 4, 17
    rays = p.rays()[::13]
    g, o = ctx.calc_rays(p, rays), oracle.calc_rays(p, rays)
    assert np.array_equal(g["error"], o["error"])
    for f in ("gvl", "evl"):
        assert np.array_equal(g[f].view(np.uint32), o[f].view(np.uint32)), f
    img, ang = ctx.create_image(p)
    oi = oracle.create_image(p)
    assert np.linalg.norm(oi["image"]) > 0
    assert rel_l2(img, oi["image"]) <= 1e-10 and max_rel(img, oi["image"]) <= 1e-9
    assert rel_l2(ang, oi["I_ang"]) <= 1e-10 and max_rel(ang, oi["I_ang"]) <= 1e-9


def test_seed_narrower_than_the_beam(seed_small, oracle, ctx, monkeypatch):
    """A seed profile that covers only part of the seed beam in x and in a: rays outside it carry
    no seed (calc_seed_inline's range test; NaN entries of the per-index factor tables), and a
    hand-off arena of 2 MB so that the image is assembled from many chunks."""
    p0 = seed_small[0]
    sd = p0.seed
    x0 = np.linspace(0.0022, 0.0051, 97)
    a0 = np.linspace(-1.1, 0.9, 61)
    x = [x0, sd.x[1], a0, sd.x[3], sd.x[4]]
    f = [np.interp(x0, sd.x[0], sd.f[0]), sd.f[1], np.interp(a0, sd.x[2], sd.f[2]), sd.f[3], sd.f[4]]
    p = abi.Problem(p0.euv_beam, p0.gain, p0.seed_beam, abi.SeedProfile(x, f, sd.f0), 11, 173)
    oi = oracle.create_image(p)
    assert np.linalg.norm(oi["image"]) > 0
    from raytrace_miniapp_b200 import lib
    monkeypatch.setenv("RTB200_HANDOFF_MB", "2")
    small_arena = lib.Context(0)
    for c in (ctx, small_arena):
        img, ang = c.create_image(p)
        assert rel_l2(img, oi["image"]) <= 1e-10 and max_rel(img, oi["image"]) <= 1e-9
        assert rel_l2(ang, oi["I_ang"]) <= 1e-10 and max_rel(ang, oi["I_ang"]) <= 1e-9
    small_arena.close()
    g = ctx.calc_rays(p, p.rays()[::5])
    o = oracle.calc_rays(p, p.rays()[::5])
    assert (np.abs(o["Iv"]).max(axis=1) == 0).sum() > 100  # rays without seed exist ...
    assert (np.abs(o["Iv"]).max(axis=1) > 0).sum() > 100  # ... and rays with seed
    sc = np.abs(o["Iv"]).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(g["Iv"] - o["Iv"]) / sc) < 1e-12
