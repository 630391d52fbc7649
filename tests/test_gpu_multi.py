"""Several devices behind the C ABI (rtb200_multi_*: one context per device, NCCL exchange) and
the compact-row form of the row-cyclic decomposition (rtb200_launch_rows_compact +
rtb200_unpermute_rows).  The multi-device tests skip on a box with one GPU; the compact-row
test emulates the shares of W devices on one."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _n_devices(rtlib):
    return rtlib.device_count()


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_compact_rows_and_unpermute_equal_whole_image(world, ase_small, ctx):
    """Each emulated device writes its rows compactly; the gathered blocks un-permuted give the
    single-launch image bit for bit (ragged last rows included: ny = 25)."""
    import torch
    p, _ = ase_small
    e = p.euv_beam
    ref_img, ref_ang = ctx.create_image(p)
    n_pix = ctx.stage(p)
    info = ctx.staged_info()
    assert info["owner"] == 1 and info["sny"] == e.ny and n_pix == e.nx * e.ny
    per = (e.ny + world - 1) // world
    n_rows = per * e.nx * e.nv
    dev = torch.device("cuda", 0)
    gathered = torch.full((world * n_rows,), float("nan"), dtype=torch.float64, device=dev)
    I_ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
    for r in range(world):
        ctx.launch_rows_compact(r, world, gathered[r * n_rows:(r + 1) * n_rows], I_ang)
        ctx.sync()
    image = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
    ctx.unpermute_rows(gathered, world, per, image)
    ctx.sync()
    assert np.array_equal(image.cpu().numpy(), ref_img)
    assert rel_l2(I_ang.cpu().numpy(), ref_ang) < 1e-13


def test_lazy_tables_staging_matches(ase_small, ctx, rtlib):
    """rtb200_stage with RTB200_FLAG_LAZY_TABLES uploads the lineshape tables at the first launch."""
    import torch
    from raytrace_miniapp_b200 import abi
    p, _ = ase_small
    e = p.euv_beam
    ref_img, ref_ang = ctx.create_image(p)
    c = rtlib.Context(0)
    n_pix = c.stage(p, flags=abi.FLAG_LAZY_TABLES)
    dev = torch.device("cuda", 0)
    image = torch.zeros(e.nx * e.ny * e.nv, dtype=torch.float64, device=dev)
    I_ang = torch.zeros(e.na * e.nb, dtype=torch.float64, device=dev)
    c.launch(0, n_pix, image, I_ang)
    c.sync()
    assert np.array_equal(image.cpu().numpy(), ref_img)
    # a second launch on the same staging finds the tables in place
    image.zero_()
    I_ang.zero_()
    c.launch(0, n_pix, image, I_ang)
    c.sync()
    assert np.array_equal(image.cpu().numpy(), ref_img)
    c.close()


def test_multi_on_one_device_equals_single(ase_small, seed_small, ctx, rtlib):
    """rtb200_multi with one device needs no NCCL and must reproduce rtb200_create_image."""
    m = rtlib.MultiContext(1)
    p, _ = ase_small
    img, ang = m.create_image(p)
    ref_img, ref_ang = ctx.create_image(p)
    assert np.array_equal(img, ref_img) and rel_l2(ang, ref_ang) < 1e-13
    q, _ = seed_small
    q.N_start, q.N_parallel = 3, 97
    try:
        img2, ang2 = m.create_image(q)
        ref2, refa2 = ctx.create_image(q)
    finally:
        q.N_start, q.N_parallel = 0, 1
    assert rel_l2(img2, ref2) < 1e-12 and rel_l2(ang2, refa2) < 1e-12
    t = m.timings()
    assert t["total_ms"] > 0 and len(t["per_device"]) == 1
    m.close()


def test_multi_devices_equal_single(ase_small, seed_small, ctx, rtlib):
    """All devices of the box through rtb200_multi_create_image (NCCL gather of owned rows for
    ASE, NCCL sum for the seeded path): ASE image bit-identical to one device."""
    n = _n_devices(rtlib)
    if n < 2:
        pytest.skip("needs at least 2 CUDA devices")
    m = rtlib.MultiContext(n)
    p, _ = ase_small
    img, ang = m.create_image(p)
    ref_img, ref_ang = ctx.create_image(p)
    assert np.array_equal(img, ref_img)
    assert rel_l2(ang, ref_ang) < 1e-13
    q, _ = seed_small
    q.N_start, q.N_parallel = 3, 97
    try:
        img2, ang2 = m.create_image(q)
        ref2, refa2 = ctx.create_image(q)
    finally:
        q.N_start, q.N_parallel = 0, 1
    assert rel_l2(img2, ref2) < 1e-12 and rel_l2(ang2, refa2) < 1e-12
    # failures are collected from every device
    bad = ase_small[0]
    t = m.timings()
    assert t["exchange_ms"] > 0 and len(t["per_device"]) == n
    m.close()
    del bad
