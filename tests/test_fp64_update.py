"""The FP64 frequency-bin update (csrc/rtb200_fp64.cuh) compiled for the host: the fast exp and
the two update branches against libm / the reference formula
(src/common/RayTraceImageHelper.h:549-557).  Budget: the path's tolerance is 1e-10 on the
image; the update is held to <= 4e-16 + 3.4e-17*|x| (exp: one correctly rounded ln2/128 in the
range reduction) and <= 1e-12 (one emission update)."""
import ctypes as C
import math

import numpy as np
import pytest

from test_march_hostsim import hostsim  # noqa: F401  (fixture)


def test_fast_exp_matches_libm(hostsim):
    hostsim.hostsim_exp.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-3, 3, 400000),
                        rng.uniform(-1e-3, 1e-3, 50000), [0.0, 1e-3, -1e-3, 699.9, -699.9, 1e-300],
                        np.arange(-128, 129) * math.log(2) / 128, (np.arange(-128, 129) + 0.5) * math.log(2) / 128])
    y = np.zeros_like(x)
    hostsim.hostsim_exp(x.ctypes.data, y.ctypes.data, x.size)
    ref = np.exp(x)
    rel = np.abs(y - ref) / ref
    budget = 4e-16 + 3.4e-17 * np.abs(x)
    assert (rel < budget).all(), (rel / budget).max()
    assert rel[np.abs(x) <= 3].max() < 5e-16, rel[np.abs(x) <= 3].max()
    # outside the fast range the library routine's semantics apply
    x2 = np.array([710.0, -750.0, np.inf, -np.inf, np.nan, 800.0])
    y2 = np.zeros_like(x2)
    hostsim.hostsim_exp(x2.ctypes.data, y2.ctypes.data, x2.size)
    assert y2[0] == np.inf and y2[1] == 0.0 and y2[2] == np.inf and y2[3] == 0.0 and np.isnan(y2[4])


def _reference_update(Iv, gvl, evl, g):
    gl = float(np.float32(gvl) * np.float32(g))
    el = float(np.float32(evl) * np.float32(g))
    if abs(gl) < 1e-3:
        return el * (1.0 + 0.5 * gl * (1.0 + 0.3333333333 * gl)) + Iv * (1.0 + gl * (1.0 + 0.5 * gl))
    e = math.exp(gl)
    return el / gl * (e - 1.0) + Iv * e


def test_update_matches_reference_formula(hostsim):
    f = hostsim.hostsim_ase_update
    f.restype = C.c_double
    f.argtypes = [C.c_double, C.c_float, C.c_float, C.c_float]
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(40000):
        gvl = float(np.float32(rng.choice([-1, 1]) * 10 ** rng.uniform(-6, 1.5)))
        evl = float(np.float32(10 ** rng.uniform(-8, -1)))
        g = float(np.float32(10 ** rng.uniform(-3, 0)))
        Iv = float(10 ** rng.uniform(-9, 2)) if rng.random() < 0.8 else 0.0
        got, want = f(Iv, gvl, evl, g), _reference_update(Iv, gvl, evl, g)
        scale = max(abs(want), 1e-300)
        worst = max(worst, abs(got - want) / scale)
    # the emission term suffers cancellation in exp(gl) - 1 near |gl| = 1e-3 in BOTH
    # implementations (2.2e-16 / 1e-3); everything else is at rounding level
    assert worst < 1e-12, worst


def test_small_branch_threshold_is_the_double_comparison():
    """(float) |glf| < 1e-3f  <=>  (double) |glf| < 1e-3 for every float."""
    t = np.float32(1e-3)
    assert float(t) > 1e-3 and float(np.nextafter(t, np.float32(0))) < 1e-3
    assert float(np.float32(0.05)) > 0.05 and float(np.float32(0.01)) < 0.01 < float(np.nextafter(np.float32(0.01), np.float32(1)))
