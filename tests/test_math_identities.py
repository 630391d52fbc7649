"""Exactness of the arithmetic shortcuts the march relies on (csrc/rtb200_math.cuh), checked on
the host build of the same header.  Each shortcut must give the bit pattern of the reference's
plain C++ expression for EVERY operand it can meet, because the march branches on these values."""
import ctypes as C

import numpy as np
import pytest

from test_march_hostsim import hostsim  # noqa: F401  (fixture)


def _bits(x):
    return int(np.float32(x).view(np.uint32))


def test_markstein_division_by_precomputed_reciprocal(hostsim):
    f = hostsim.hostsim_check_ddiv_by
    f.restype = C.c_longlong
    f.argtypes = [C.c_longlong, C.c_ulonglong]
    assert f(20_000_000, 12345) == 0
    assert f(20_000_000, 987654321) == 0


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
