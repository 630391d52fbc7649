"""Parity of the CUDA path (through the C ABI of include/rtb200.h) against the CPU oracle and
the committed reference outputs.  Tolerances: march intermediates bit-exact; image / I_ang
relative L2 <= 1e-10 and max element-wise relative error <= 1e-9 over entries > 1e-6*max
(BASELINE.json north_star)."""
import numpy as np
import pytest

from raytrace_miniapp_b200 import abi
from conftest import max_rel, rel_l2

pytestmark = pytest.mark.gpu

TOL_L2 = 1e-10
TOL_MAX = 1e-9


def check_image(got, want):
    assert rel_l2(got[0], want[0]) <= TOL_L2, ("image relL2", rel_l2(got[0], want[0]))
    assert rel_l2(got[1], want[1]) <= TOL_L2, ("I_ang relL2", rel_l2(got[1], want[1]))
    assert max_rel(got[0], want[0]) <= TOL_MAX, ("image max rel", max_rel(got[0], want[0]))
    assert max_rel(got[1], want[1]) <= TOL_MAX, ("I_ang max rel", max_rel(got[1], want[1]))


def test_march_intermediates_bit_exact(ase_small, oracle, ctx):
    """gvl / evl / ivl of >= 10^4 rays are bit-identical to the oracle's (FP32 march)."""
    p, _ = ase_small
    rays = p.rays()[3::37]
    assert rays.size >= 10000
    g = ctx.calc_rays(p, rays)
    o = oracle.calc_rays(p, rays)
    assert np.array_equal(g["error"], o["error"])
    assert np.array_equal(g["gvl"].view(np.uint32), o["gvl"].view(np.uint32))
    assert np.array_equal(g["evl"].view(np.uint32), o["evl"].view(np.uint32))
    assert np.array_equal(g["ivl"], o["ivl"])
    ok = o["error"] == 0
    assert np.array_equal(g["ray2"]["x"][ok], o["ray2"]["x"][ok])
    assert np.array_equal(g["ray2"]["y"][ok], o["ray2"]["y"][ok])
    # per-ray spectra: FP64, libm exp differs by <= 1 ulp between host and device
    scale = np.abs(o["Iv"]).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(g["Iv"] - o["Iv"]) / scale) < 1e-12


def test_exit_angles_bit_exact(ase_small, seed_small, oracle, ctx):
    """ray2.a / ray2.b = atanf(s.x/s.z)*1e3f feed a discrete bin in seeded mode."""
    for (p, _), sl in ((ase_small, slice(5, None, 211)), (seed_small, slice(7, None, 4001))):
        rays = p.rays()[sl]
        p2 = abi.Problem(p.euv_beam, p.gain)  # calc_rays without the seed: march only
        g = ctx.calc_rays(p2, rays, method=p.method)
        o = oracle.calc_rays(p2, rays, method=p.method)
        ok = o["error"] == 0
        for f in "xyab":
            assert np.array_equal(g["ray2"][f][ok].view(np.uint32), o["ray2"][f][ok].view(np.uint32)), f


def test_ase_small_image_vs_oracle_and_reference(ase_small, oracle, ctx):
    p, extra = ase_small
    img, ang = ctx.create_image(p)
    assert ctx.failure_code == 0
    o = oracle.create_image(p)
    check_image((img, ang), (o["image"], o["I_ang"]))
    check_image((img, ang), (extra["ref_cpu_image"], extra["ref_cpu_I_ang"]))
    # the reference's own acceptance test (check_ans, src/CreateImageHelpers.cpp:66-100)
    g0, g1 = np.linalg.norm(extra["dat_golden_image"]), np.linalg.norm(img)
    assert (g0 - g1) / g0 <= 5e-6
    t = ctx.timings()
    assert t["kernel_launches"] >= 1 and t["march_ms"] + t["integrate_ms"] > 0


def test_seed_small_image_vs_reference(seed_small, ctx):
    p, extra = seed_small
    img, ang = ctx.create_image(p)
    assert ctx.failure_code == 0
    check_image((img, ang), (extra["ref_cpu_image"], extra["ref_cpu_I_ang"]))


def test_strided_decomposition(ase_small, oracle, ctx):
    """N_start / N_parallel (src/RayTraceImage.cpp:300-308): each worker matches the oracle and
    the workers add up to the full image."""
    p, extra = ase_small
    tot_i, tot_a = 0, 0
    try:
        for start in range(3):
            p.N_start, p.N_parallel = start, 3
            img, ang = ctx.create_image(p)
            tot_i, tot_a = tot_i + img, tot_a + ang
            if start == 1:
                o = oracle.create_image(p)
                check_image((img, ang), (o["image"], o["I_ang"]))
    finally:
        p.N_start, p.N_parallel = 0, 1
    check_image((tot_i, tot_a), (extra["ref_cpu_image"], extra["ref_cpu_I_ang"]))


def test_explicit_ray_list_accumulates(ase_small, oracle, ctx):
    """RayTraceImage<B200>Loop semantics: arbitrary ray list, += into caller buffers."""
    p, _ = ase_small
    rays = p.rays()[::53]
    o = oracle.trace_rays(p, rays, 1, 1.0)
    img = np.full(o["image"].size, 1.0)
    ang = np.full(o["I_ang"].size, 2.0)
    ctx.trace_rays(p, rays, 1, 1.0, image=img, I_ang=ang)
    check_image((img - 1.0, ang - 2.0), (o["image"], o["I_ang"]))
    # empty list: nothing added, no error
    ctx.trace_rays(p, rays[:0], 1, 1.0, image=img, I_ang=ang)
    check_image((img - 1.0, ang - 2.0), (o["image"], o["I_ang"]))


def test_failures_are_reported_like_the_reference(ase_small, oracle, ctx, rtlib):
    """Negative emissivity cannot occur (clamped) but a NaN gain makes rays fail with code -3;
    failed rays are excluded from the image and listed (src/RayTraceImageCPU.cpp:32-36)."""
    p, _ = ase_small
    g = p.gain[2]
    saved = g.g0.copy()
    g.g0[5:8, 40:60] = np.nan
    p.N_parallel = 11
    try:
        o = oracle.create_image(p)
        assert o["failure_code"] == 8 and o["n_failed"] > 0
        img, ang = ctx.create_image(p, raise_on_failed=False)
        assert ctx.failure_code == o["failure_code"]
        assert ctx.n_failed == o["n_failed"]
        check_image((img, ang), (o["image"], o["I_ang"]))
        got = set(map(tuple, ctx.failed.view(np.float32).reshape(-1, 4).tolist()))
        allf = set(map(tuple, p.rays().view(np.float32).reshape(-1, 4).tolist()))
        assert got and got <= allf
        with pytest.raises(rtlib.RaysFailed):
            ctx.create_image(p)
    finally:
        g.g0[:] = saved
        p.N_parallel = 1


def test_limits_and_grid_errors(ase_small, ctx, rtlib):
    p, _ = ase_small
    x = p.euv_beam.x
    p.euv_beam.x = x.copy()
    p.euv_beam.x[7] *= 1.0 + 1e-9
    try:
        with pytest.raises(rtlib.RTB200Error) as e:
            ctx.create_image(p)
        assert e.value.code == abi.ERR_GRID and "uniform grid" in str(e.value)
    finally:
        p.euv_beam.x = x
    many = abi.Problem(p.euv_beam, [p.gain[i % 3] for i in range(21)])
    with pytest.raises(rtlib.RTB200Error) as e:
        ctx.create_image(many)
    assert e.value.code == abi.ERR_LIMITS


def test_single_plane_problem_is_empty(ase_small, ctx, oracle):
    """N = 1: no length segments, the image is identically zero."""
    p, _ = ase_small
    one = abi.Problem(p.euv_beam, p.gain[:1])
    one.N_parallel = 17
    img, ang = ctx.create_image(one)
    assert not img.any() and not ang.any()
    o = oracle.create_image(one)
    assert not o["image"].any()


def test_device_tiles_equal_whole_image(ase_small, ctx):
    """Pixel-tile sharding (the multi-GPU decomposition) on one device: tiles reproduce the
    single-launch image bit for bit in ASE mode (disjoint pixels, same per-pixel order)."""
    import torch
    p, _ = ase_small
    e = p.euv_beam
    n = ctx.stage(p)
    assert n == e.nx * e.ny and ctx.staged_rays == p.n_rays
    whole_i = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device="cuda")
    whole_a = torch.zeros(e.na * e.nb, dtype=torch.float64, device="cuda")
    ctx.launch(0, n, whole_i, whole_a)
    ctx.sync()
    tile_i, tile_a = torch.zeros_like(whole_i), torch.zeros_like(whole_a)
    cuts = [0, 100, 101, n // 2, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ctx.launch(a, b, tile_i, tile_a)
    ctx.sync()
    assert torch.equal(tile_i, whole_i)
    assert rel_l2(tile_a.cpu().numpy(), whole_a.cpu().numpy()) < 1e-14


def test_row_cyclic_shares_equal_whole_image(ase_small, ctx):
    """The multi-GPU decomposition (rows r, r + W, ...) emulated on one device: the W shares sum
    to the single-launch image bit for bit."""
    import torch
    p, _ = ase_small
    e = p.euv_beam
    n = ctx.stage(p)
    whole_i = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device="cuda")
    whole_a = torch.zeros(e.na * e.nb, dtype=torch.float64, device="cuda")
    ctx.launch(0, n, whole_i, whole_a)
    ctx.sync()
    for world in (2, 3, 8):
        acc_i, acc_a = torch.zeros_like(whole_i), torch.zeros_like(whole_a)
        for r in range(world):
            part_i, part_a = torch.zeros_like(whole_i), torch.zeros_like(whole_a)
            ctx.launch_rows(r, world, part_i, part_a)
            ctx.sync()
            rows = part_i.view(e.ny, e.nx * e.nv)
            other = [j for j in range(e.ny) if j % world != r]
            assert not rows[other].any()  # a rank never touches another rank's rows
            acc_i += part_i
            acc_a += part_a
        assert torch.equal(acc_i, whole_i)
        assert rel_l2(acc_a.cpu().numpy(), whole_a.cpu().numpy()) < 1e-14


def test_ray_trajectories(ase_small, seed_small, oracle, ctx):
    """rtb200_calc_ray_paths == RayTrace::calc_ray_path (RAY_DEBUG trajectories): positions at the
    sub-segment boundaries bit-exact, running intensity (a float sum over bins) to 1e-5."""
    for (p, extra), c in ((ase_small, 0.5), (seed_small, 0.5), (ase_small, 0.3)):
        rays = p.rays()[extra["sample_index"]][:400]
        g = ctx.calc_ray_paths(p, rays, c=c)
        o = oracle.calc_ray_paths(p, rays, c=c)
        assert np.array_equal(g["error"], o["error"])
        assert np.array_equal(g["x"].view(np.uint32), o["x"].view(np.uint32))
        assert np.array_equal(g["y"].view(np.uint32), o["y"].view(np.uint32))
        scale = np.abs(o["I"]).max(axis=1, keepdims=True) + 1e-30
        assert np.max(np.abs(g["I"] - o["I"]) / scale) < 1e-5
        assert o["I"].max() > 0
        if c == 0.5:  # the committed trajectories of the unmodified reference
            ref = extra["sample_debug"][:400].reshape(400, -1, 3)
            assert np.array_equal(g["x"], ref[:, :, 0]) and np.array_equal(g["y"], ref[:, :, 1])


@pytest.mark.parametrize("env", [{"RTB200_HANDOFF_MB": "8"}, {"RTB200_IEEE_DIV": "1"},
                                 {"RTB200_MARCH_BLOCKS": "7"}],
                         ids=["small-handoff-chunks", "ieee-divisions", "tiny-march-grid"])
def test_alternative_kernel_paths(env, ase_small, rtlib, monkeypatch):
    """A hand-off arena that forces many chunks and a march grid of a few CTAs give the same image
    as the default path (bit for bit: same march, same per-pixel summation order); so do plain
    IEEE divisions in place of the reciprocal-table divisions."""
    p, extra = ase_small
    base = rtlib.Context(0)
    img0, ang0 = base.create_image(p)
    base.close()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    alt = rtlib.Context(0)
    img1, ang1 = alt.create_image(p)
    alt.close()
    check_image((img1, ang1), (extra["ref_cpu_image"], extra["ref_cpu_I_ang"]))
    assert np.array_equal(img0, img1)


def _warped(p):
    from raytrace_miniapp_b200 import synth
    q = abi.Problem(p.euv_beam, [p.gain[0]] + [synth.warp_gain_grid(g) for g in p.gain[1:]],
                    p.seed_beam, p.seed, p.N_start, p.N_parallel)
    return q


def test_non_uniform_gain_grid(ase_small, oracle, ctx, rtlib, monkeypatch):
    """Gain planes on a non-uniform (x, y) grid: every cell has its own width, the index guess of
    the interval table is wrong for most cells (generic search), and the reciprocal-table
    divisions meet ~130 distinct divisors per plane.  March bit-exact, image within tolerance,
    and identical bits with IEEE divisions."""
    p = _warped(ase_small[0])
    assert np.unique(np.diff(p.gain[1].x).round(12)).size > 50
    rays = p.rays()[11::97]
    g = ctx.calc_rays(p, rays)
    o = oracle.calc_rays(p, rays)
    assert np.array_equal(g["error"], o["error"]) and (o["error"] == 0).sum() > 1000
    for f in ("gvl", "evl"):
        assert np.array_equal(g[f].view(np.uint32), o[f].view(np.uint32)), f
    assert np.array_equal(g["ivl"], o["ivl"])
    p.N_start, p.N_parallel = 5, 19
    img, ang = ctx.create_image(p)
    oi = oracle.create_image(p)
    assert np.linalg.norm(oi["image"]) > 0
    check_image((img, ang), (oi["image"], oi["I_ang"]))
    monkeypatch.setenv("RTB200_IEEE_DIV", "1")
    alt = rtlib.Context(0)
    g2 = alt.calc_rays(p, rays)
    img2, ang2 = alt.create_image(p)
    alt.close()
    for f in ("gvl", "evl"):
        assert np.array_equal(g[f].view(np.uint32), g2[f].view(np.uint32)), f
    assert np.array_equal(img, img2)
    assert rel_l2(ang, ang2) < 1e-13  # I_ang is summed with FP64 atomics: order varies run to run


def test_vacuum_planes_and_empty_inputs(ase_small, oracle, ctx):
    """Zero gain and emission in one plane (identity records on the integration side) and the
    empty cases of the explicit-ray entry points."""
    p0 = ase_small[0]
    planes = list(p0.gain)
    for i in (1,):
        g = planes[i]
        planes[i] = abi.Gain(g.x, g.y, g.n, np.zeros_like(g.g0), np.zeros_like(g.E0), g.gv, g.gv0)
    p = abi.Problem(p0.euv_beam, planes, None, None, 3, 31)
    img, ang = ctx.create_image(p)
    o = oracle.create_image(p)
    assert np.linalg.norm(o["image"]) > 0
    check_image((img, ang), (o["image"], o["I_ang"]))
    none = p.rays()[:0]
    g = ctx.calc_rays(p, none)
    assert g["error"].size == 0 and g["Iv"].shape[0] == 0
    img0 = np.zeros_like(img)
    ang0 = np.zeros_like(ang)
    ctx.trace_rays(p, none, p.method, 1.0, img0.ravel(), ang0.ravel())
    assert not img0.any() and not ang0.any()


def test_launch_after_create_image_uses_the_staged_problem(ase_small, ctx):
    """create_image stages the problem in two uploads (the lineshape tables follow while the
    march runs); a launch on the same context afterwards must find the staging complete, and a
    second create_image with other gains must not see the tables of the first."""
    import torch
    p, _ = ase_small
    e = p.euv_beam
    img, ang = ctx.create_image(p)
    dev_i = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device="cuda")
    dev_a = torch.zeros(e.na * e.nb, dtype=torch.float64, device="cuda")
    ctx.launch(0, ctx.staged_pixels, dev_i, dev_a)
    ctx.sync()
    assert np.array_equal(dev_i.cpu().numpy(), img)
    assert rel_l2(dev_a.cpu().numpy(), ang) < 1e-14
    planes = [abi.Gain(g.x, g.y, g.n, g.g0, g.E0, g.gv * np.float32(0.5), g.gv0) for g in p.gain]
    q = abi.Problem(p.euv_beam, planes, None, None, p.N_start, p.N_parallel)
    img2, _ = ctx.create_image(q)
    fresh = type(ctx)(0)
    img3, _ = fresh.create_image(q)
    fresh.close()
    assert np.array_equal(img2, img3) and not np.array_equal(img2, img)


@pytest.mark.parametrize("name", ["ase_small", "seed_small"])
def test_create_image_from_dat_bytes(name, request, ctx):
    """rtb200_create_image_from_dat: the serialized problem (unaligned byte stream; the large
    arrays are read once, from the stream into the pinned blob) gives the bits of the array form."""
    import raytrace_miniapp_b200 as rt
    p, extra = request.getfixturevalue(name)
    stride = 1 if name == "ase_small" else 53
    p.N_start, p.N_parallel = (0, 1) if stride == 1 else (5, stride)
    try:
        payload = b"\x00" + rt.pack_payload(p)  # one more byte of misalignment for good measure
        e = p.euv_beam
        img, ang = ctx.create_image_from_dat(memoryview(payload)[1:], e.nx * e.ny * e.nv, e.na * e.nb)
        ref_img, ref_ang = ctx.create_image(p)
    finally:
        p.N_start, p.N_parallel = 0, 1
    if name == "ase_small":
        assert np.array_equal(img, ref_img)
    else:  # scatter binning: atomics in another order
        assert rel_l2(img, ref_img) < 1e-13
    assert rel_l2(ang, ref_ang) < 1e-13
    with pytest.raises(Exception):
        ctx.create_image_from_dat(payload[1:200], 8, 8)


def test_overlapped_launch_matches_serial(ase_small, monkeypatch):
    """The default ASE launch starts the integration while the march drains (programmatic dependent
    launch, per-pixel completion counts; with rtb200_create_image's lazily uploaded lineshape
    tables also the epoch word behind their copy).  Same image bits as RTB200_OVERLAP=0, image
    after image on one context, also when the hand-off is cut into pixel-range chunks."""
    from raytrace_miniapp_b200 import lib as rl
    p, _ = ase_small
    monkeypatch.setenv("RTB200_OVERLAP", "0")
    c0 = rl.Context(0)
    img0, ang0 = c0.create_image(p)
    t0 = c0.timings()
    assert t0["march_ms"] > 0 and t0["integrate_ms"] > 0
    c0.close()
    monkeypatch.setenv("RTB200_OVERLAP", "1")
    c1 = rl.Context(0)
    for _ in range(3):
        img1, ang1 = c1.create_image(p)
        assert np.array_equal(img0, img1)
        assert rel_l2(ang1, ang0) < 1e-13  # (atomics: order of the additions)
        t1 = c1.timings()
        assert t1["march_ms"] > 0 and t1["integrate_ms"] == 0 and t1["kernel_launches"] == 2
    c1.close()
    monkeypatch.setenv("RTB200_HANDOFF_MB", "16")  # several chunks, each march + integration overlapped
    c2 = rl.Context(0)
    img2, ang2 = c2.create_image(p)
    assert c2.timings()["kernel_launches"] > 2
    assert np.array_equal(img0, img2) and rel_l2(ang2, ang0) < 1e-13
    c2.close()
