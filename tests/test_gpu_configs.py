"""BASELINE.json configs 4 and 5: the synthetic finer-grid / higher-resolution
family and the high-spectral-resolution sweep, each validated against the CPU oracle on a
strided ray sample (the oracle has no N or K limit; the reference itself stops at K < 100)."""
import numpy as np
import pytest

from raytrace_miniapp_b200 import abi, synth
from conftest import max_rel, rel_l2

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, p, flags=0):
    img, ang = ctx.create_image(p, flags=flags)
    o = oracle.create_image(p, flags=flags)
    assert o["rc"] == abi.OK and ctx.failure_code == 0
    assert np.linalg.norm(o["image"]) > 0
    assert rel_l2(img, o["image"]) <= 1e-10 and rel_l2(ang, o["I_ang"]) <= 1e-10
    assert max_rel(img, o["image"]) <= 1e-9 and max_rel(ang, o["I_ang"]) <= 1e-9


@pytest.mark.parametrize("gain_factor,image_factor", [(2, 2), (4, 1)], ids=["S4", "S4b-gain"])
def test_finer_gain_grid_and_image(gain_factor, image_factor, ase_small, oracle, ctx):
    p = synth.s4(ase_small[0], gain_factor=gain_factor, image_factor=image_factor)
    assert p.gain[1].Nx == (106 - 1) * gain_factor + 1
    p.N_start, p.N_parallel = 2, 23
    _check(ctx, oracle, p)


@pytest.mark.parametrize("K,angle_factor", [(52, 2), (99, 1), (128, 1), (200, 1), (512, 1)])
def test_spectral_sweep(K, angle_factor, ase_small, oracle, ctx):
    p = synth.spectral(ase_small[0], K, angle_factor=angle_factor)
    assert p.euv_beam.nv == K
    p.N_start, p.N_parallel = 1, 29 * angle_factor * angle_factor
    _check(ctx, oracle, p, flags=abi.FLAG_NO_LIMITS)


def test_reference_limits_are_enforced_by_default(ase_small, ctx, rtlib):
    p = synth.spectral(ase_small[0], 100)
    with pytest.raises(rtlib.RTB200Error) as e:
        ctx.create_image(p)
    assert e.value.code == abi.ERR_LIMITS and "frequencies" in str(e.value)


def test_deep_stack_of_planes(ase_small, oracle, ctx):
    """N = 20 length planes (the reference's N_MAX): 57 hand-off records per ray."""
    small = ase_small[0]
    g = small.gain
    planes = [g[0]] + [synth.blend_planes(g[1], g[2], w) for w in np.linspace(0, 1, 19)]
    p = abi.Problem(small.euv_beam, planes)
    p.N_start, p.N_parallel = 0, 37
    _check(ctx, oracle, p)
