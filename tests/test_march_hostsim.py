"""The march's device source (csrc/rtb200_march.cuh, rtb200_pack.h) compiled for the HOST and
checked bit for bit against the oracle — the same functions the GPU kernel inlines, so index
search, sub-segment bookkeeping and packing are covered without a GPU.  tests/hostsim is test
infrastructure, not part of librtb200.so."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from raytrace_miniapp_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostsim():
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    so = os.path.join(HERE, "hostsim", "libhostsim.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", so, src],
                   check=True)
    L = C.CDLL(so)
    L.hostsim_march.restype = C.c_longlong
    L.hostsim_march.argtypes = [C.POINTER(abi.CProblem), C.c_longlong, C.c_longlong] + [C.c_void_p] * 5 + [C.c_int]
    L.hostsim_pchip.restype = C.c_double
    L.hostsim_pchip.argtypes = [C.c_size_t, abi.c_double_p, abi.c_double_p, C.c_double]
    L.hostsim_get_index.argtypes = [C.c_int, abi.c_double_p, C.c_double, C.c_double]
    L.hostsim_find_cell.argtypes = [abi.c_double_p, C.c_int, C.c_double]
    L.hostsim_find_cell_fast.argtypes = [abi.c_double_p, C.c_int, C.c_double]
    L.hostsim_tables.argtypes = [C.POINTER(abi.CProblem)] + [C.c_void_p] * 7
    return L


def _march(L, p, flat=0):
    n, S = p.n_rays, (p.N - 1) * 3
    cp, keep = p.c_struct()
    gvl, evl = np.zeros((n, S), np.float32), np.zeros((n, S), np.float32)
    ivl, ex, meta = np.zeros((n, S), np.int32), np.zeros((n, 6), np.float32), np.zeros((n, 2), np.int32)
    steps = L.hostsim_march(C.byref(cp), 0, n, gvl.ctypes.data, evl.ctypes.data, ivl.ctypes.data,
                            ex.ctypes.data, meta.ctypes.data, flat)
    return gvl, evl, ivl, ex, meta, steps


@pytest.mark.parametrize("flat", [0, 1], ids=["nested", "flat"])
@pytest.mark.parametrize("name,stride,start", [("ase_small", 41, 3), ("seed_small", 1999, 11)])
def test_march_bit_identical_to_oracle(name, stride, start, flat, request, oracle, hostsim):
    """flat=0: the literal nested form (rtb200_march.cuh); flat=1: the flat state machine of
    the fused kernel (rtb200_march_flat.cuh).  Both must reproduce the oracle bit for bit."""
    p, _ = request.getfixturevalue(name)
    p.N_start, p.N_parallel = start, stride
    try:
        gvl, evl, ivl, ex, meta, steps = _march(hostsim, p, flat)
        o = oracle.calc_rays(p, p.rays())
    finally:
        p.N_start, p.N_parallel = 0, 1
    assert steps == o["steps"] and steps > 0
    assert np.array_equal(gvl.view(np.uint32), o["gvl"].view(np.uint32))
    assert np.array_equal(evl.view(np.uint32), o["evl"].view(np.uint32))
    assert np.array_equal(ivl, o["ivl"])
    assert np.array_equal(ex[:, 5].astype(np.int32), o["escaped"])
    ok = o["error"] == 0
    assert np.array_equal(ex[ok, 0], o["ray2"]["x"][ok]) and np.array_equal(ex[ok, 1], o["ray2"]["y"][ok])
    # records outside the visited range are exactly the zero records
    S = gvl.shape[1]
    s = np.arange(S)[None, :]
    outside = (s < meta[:, :1]) | (s >= meta[:, 1:])
    assert not gvl[outside].any() and not evl[outside].any() and not ivl[outside].any()
    assert o["escaped"].any() or name == "seed_small"


def test_find_cell_equals_reference_bisection(oracle, hostsim):
    rng = np.random.default_rng(3)
    for n in (2, 3, 26, 106):
        for uniform in (True, False):
            X = np.linspace(0.0, 6.95e-3, n) if uniform else np.cumsum(rng.uniform(0.1, 3.0, n)) * 1e-4
            Xp = X.ctypes.data_as(abi.c_double_p)
            ys = np.concatenate([rng.uniform(X[0] - 1e-3, X[-1] + 1e-3, 400), X, np.nextafter(X, 1), np.nextafter(X, -1),
                                 [np.nan, np.inf, -np.inf]])
            for Y in ys:
                assert hostsim.hostsim_find_cell(Xp, n, Y) == oracle.L.rt_oracle_findindex(Xp, n, Y), (n, uniform, Y)
                Yf = float(np.float32(Y))  # the interval-table variant is used on float coordinates
                assert hostsim.hostsim_find_cell_fast(Xp, n, Y) == oracle.L.rt_oracle_findindex(Xp, n, Yf), (n, uniform, Y)


def test_owner_tables_and_seed_factors(seed_small, ase_small, oracle, hostsim):
    for p, _ in (ase_small, seed_small):
        g = p.ray_grid
        cp, keep = p.c_struct()
        pixI, pixJ = np.zeros(g.nx, np.int32), np.zeros(g.ny, np.int32)
        binA, binB = np.zeros(g.na, np.int32), np.zeros(g.nb, np.int32)
        tanA, tanB = np.zeros(g.na, np.float32), np.zeros(g.nb, np.float32)
        sf = np.zeros(g.nx + g.ny + g.na + g.nb)
        m = hostsim.hostsim_tables(C.byref(cp), pixI.ctypes.data, pixJ.ctypes.data, binA.ctypes.data,
                                   binB.ctypes.data, tanA.ctypes.data, tanB.ctypes.data, sf.ctypes.data)
        assert m == p.method
        e = p.euv_beam
        if p.method == 1:  # every source coordinate owns its own cell
            assert np.array_equal(pixI, np.arange(e.nx)) and np.array_equal(pixJ, np.arange(e.ny))
            assert np.array_equal(binA, np.arange(e.na)) and np.array_equal(binB, np.arange(e.nb))
        else:  # separable seed factors == calc_seed_inline on the grid points
            s = p.seed
            Iv = np.zeros(s.x[4].size)
            sd = s.c_struct()
            rng = np.random.default_rng(0)
            for _ in range(200):
                i, j, k, mm = rng.integers(g.nx), rng.integers(g.ny), rng.integers(g.na), rng.integers(g.nb)
                x, y = float(np.float32(g.x[i])), float(np.float32(g.y[j]))
                a, b = float(np.float32(g.a[k])), float(np.float32(g.b[mm]))
                oracle.L.rt_oracle_calc_seed(C.byref(sd), x, y, a, b, Iv.ctypes.data_as(abi.c_double_p))
                f = [sf[i], sf[g.nx + j], sf[g.nx + g.ny + k], sf[g.nx + g.ny + g.na + mm]]
                if any(np.isnan(f)):
                    want = 0.0
                else:
                    want = max(s.f0 * f[0] * f[1] * f[2] * f[3], 0.0)
                assert np.array_equal(Iv, want * s.f[4])
