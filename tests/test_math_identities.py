"""Exactness of the arithmetic shortcuts the march relies on (csrc/rtb200_math.cuh), checked on
the host build of the same header.  Each shortcut must give the bit pattern of the reference's
plain C++ expression for EVERY operand it can meet, because the march branches on these values."""
import ctypes as C
import os

import numpy as np
import pytest

from raytrace_miniapp_b200 import problem_io, synth
from test_march_hostsim import hostsim  # noqa: F401  (fixture)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(x):
    return int(np.float32(x).view(np.uint32))


def test_markstein_division_by_precomputed_reciprocal(hostsim):
    f = hostsim.hostsim_check_ddiv_by
    f.restype = C.c_longlong
    f.argtypes = [C.c_longlong, C.c_ulonglong]
    assert f(20_000_000, 12345) == 0
    assert f(20_000_000, 987654321) == 0


def test_markstein_divisor_proof(hostsim):
    """markstein_safe: the per-divisor search for numerators on which ddiv_by could round wrongly.

    Random operands never land within 2^-104 of a rounding midpoint, so the statistical test above
    cannot see a wrong reciprocal; the search does: with RN(1/b) one ulp off it finds a numerator
    that fails, with the correctly rounded reciprocal it finds none, for every cell width of the
    fixtures and for random divisors."""
    safe = hostsim.hostsim_markstein_safe
    safe.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_double]
    div = hostsim.hostsim_ddiv_by_rb
    div.restype = C.c_double
    div.argtypes = [C.c_double] * 3
    rng = np.random.default_rng(5)
    rejected = 0
    for B in rng.integers(2**52, 2**53, 400):
        b = float(B)
        assert safe(b, None, 0.0) == 1
        w = C.c_double(0.0)
        rb_off = float(np.nextafter(1.0 / b, 1.0))
        if safe(b, C.byref(w), rb_off) == 0:
            rejected += 1
            assert div(w.value, b, rb_off) != w.value / b  # the witness really fails ...
            assert div(w.value, b, 1.0 / b) == w.value / b  # ... and only because of the reciprocal
    assert rejected > 100  # a third or more of the divisors have such a numerator
    small, _ = problem_io.load_npz(os.path.join(GOLDEN, "ase_small.npz"))
    seed, _ = problem_io.load_npz(os.path.join(GOLDEN, "seed_small.npz"))
    for p in (small, seed, synth.ase_medium_synth(small)):
        for g in p.gain:
            for ax in (g.x, g.y):
                for w in np.diff(ax):
                    assert safe(float(w), None, 0.0) == 1
                    assert safe(float(np.float32(w)), None, 0.0) == 1
    assert safe(0.0, None, 0.0) == 0 and safe(float("inf"), None, 0.0) == 0 and safe(5e-324, None, 0.0) == 0


@pytest.mark.parametrize("c", [3.0, 6.0, 12.0])
def test_division_by_step_constants(c, hostsim):
    f = hostsim.hostsim_check_fdiv_const
    f.restype = C.c_longlong
    f.argtypes = [C.c_float, C.c_uint, C.c_uint]
    # the magnitudes st = step*t and st*st take in the march (1e-12 .. 1e3): every float, both signs
    assert f(c, _bits(1e-12), _bits(1e3)) == 0
    # the guarded ends: zero, denormals, the fast-range boundaries, huge, inf, NaN
    assert f(c, 0, 0x00800100) == 0
    assert f(c, _bits(1e-30) - 4096, _bits(1e-30) + 4096) == 0
    assert f(c, _bits(1e30) - 4096, _bits(1e30) + 4096) == 0
    assert f(c, 0x7f7ff000, 0x7f800010) == 0


def test_reciprocal_square_root_identity(hostsim):
    """normalize_s: (float)(1.0 / (double)sqrtf(x)) == 1.0f / sqrtf(x) (|s|^2 is close to 1)."""
    f = hostsim.hostsim_check_rsqrt_identity
    f.restype = C.c_longlong
    f.argtypes = [C.c_uint, C.c_uint]
    assert f(_bits(0.25), _bits(4.0)) == 0
