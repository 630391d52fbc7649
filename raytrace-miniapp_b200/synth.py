"""Deterministic synthetic inputs derived from an existing problem.

`ASE_medium.dat` / `seed_medium.dat` are not part of the reference checkout
(.MISSING_LARGE_BLOBS), so the "ASE_medium" configuration of BASELINE.json is a documented
stand-in built from ASE_small: no random numbers anywhere.
"""
import numpy as np

from .abi import BeamGrid, Gain, Problem


def scale_beam(g, scale):
    """scale_beam (src/CreateImageHelpers.cpp:104-142): refine nx, ny, na, nb by `scale` over the
    same physical extent; cell centres at x0 + (0.5 + i)*dx."""
    def refine(x, dx, n_new):
        lo, hi = x[0] - 0.5 * dx, x[-1] + 0.5 * dx
        d = (hi - lo) / n_new
        return lo + (0.5 + np.arange(n_new)) * d, d

    n = [int(v * scale) for v in (g.nx, g.ny, g.na, g.nb)]
    x, dx = refine(g.x, g.dx, n[0])
    y, dy = refine(g.y, g.dy, n[1])
    a, da = refine(g.a, g.da, n[2])
    b, db = refine(g.b, g.db, n[3])
    return BeamGrid(x, y, a, b, dx, dy, da, db, dv=g.dv if g.dv.size else None, dz=g.dz,
                    extra=g.extra)


def refine_rows(g, factor):
    """Refine only ny by an integer factor (the sharded axis of the weak-scaling runs)."""
    lo, hi = g.y[0] - 0.5 * g.dy, g.y[-1] + 0.5 * g.dy
    n = g.ny * factor
    dy = (hi - lo) / n
    y = lo + (0.5 + np.arange(n)) * dy
    return BeamGrid(g.x, y, g.a, g.b, g.dx, dy, g.da, g.db, dv=g.dv if g.dv.size else None,
                    dz=g.dz, extra=g.extra)


def blend_planes(g1, g2, w):
    """Linear blend of two gain planes (all arrays; n in double, the rest rounded to float)."""
    mix = lambda a, b: (1.0 - w) * a.astype(np.float64) + w * b.astype(np.float64)  # noqa: E731
    return Gain(g1.x, g1.y, mix(g1.n, g2.n), mix(g1.g0, g2.g0).astype(np.float32),
                None if g1.E0 is None else mix(g1.E0, g2.E0).astype(np.float32),
                mix(g1.gv, g2.gv).astype(np.float32), mix(g1.gv0, g2.gv0).astype(np.float32))


def ase_medium_synth(small, rows_factor=1):
    """ASE_medium stand-in: euv grids refined as `-scale=8` does
    (nx,ny,na,nb x 8^0.25 -> 100 x 42 x 31 x 23 = 2 994 600 rays) and N raised from 3 to 6 by
    inserting linear blends of gain[1], gain[2].  rows_factor > 1 refines ny further (weak
    scaling: rays grow with the GPU count, rows are the sharded axis)."""
    euv = scale_beam(small.euv_beam, 8.0 ** 0.25)
    if rows_factor > 1:
        euv = refine_rows(euv, rows_factor)
    g = small.gain
    planes = [g[0]] + [blend_planes(g[1], g[2], w) for w in (0.0, 0.25, 0.5, 0.75, 1.0)]
    return Problem(euv, planes, None, None, 0, 1)


def seed_medium_synth(small):
    """seed_medium stand-in (BASELINE.json config 2; seed_medium.dat is not in the reference
    checkout, .MISSING_LARGE_BLOBS:2): seed_small with BOTH beams refined as `-scale=8` does
    (scale_problem, src/CreateImageHelpers.cpp:144-150: euv_beam and seed_beam x 8^0.25 per
    axis -> seed grid 201 x 42 x 85 x 85 = 60 994 350 rays) and N raised from 3 to 6 with the
    same linear blends of gain[1], gain[2] as ase_medium_synth.  The seed profile is unchanged."""
    f = 8.0 ** 0.25
    euv = scale_beam(small.euv_beam, f)
    sb = scale_beam(small.seed_beam, f)
    g = small.gain
    planes = [g[0]] + [blend_planes(g[1], g[2], w) for w in (0.0, 0.25, 0.5, 0.75, 1.0)]
    return Problem(euv, planes, sb, small.seed, 0, 1)


def resample_gain(g, fx, fy):
    """Bilinear resampling of a gain plane to fx x fy finer cells (S4 family, config 4)."""
    Nx, Ny = (g.Nx - 1) * fx + 1, (g.Ny - 1) * fy + 1
    x = np.interp(np.arange(Nx) / fx, np.arange(g.Nx), g.x)
    y = np.interp(np.arange(Ny) / fy, np.arange(g.Ny), g.y)
    ix, iy = np.arange(Nx) / fx, np.arange(Ny) / fy
    i0 = np.minimum(ix.astype(int), g.Nx - 2)
    j0 = np.minimum(iy.astype(int), g.Ny - 2)
    tx, ty = (ix - i0)[None, :], (iy - j0)[:, None]

    def bil(a):
        a = a.astype(np.float64)
        if a.ndim == 3:
            txx, tyy = tx[..., None], ty[..., None]
        else:
            txx, tyy = tx, ty
        a00 = a[j0][:, i0]
        a01 = a[j0][:, i0 + 1]
        a10 = a[j0 + 1][:, i0]
        a11 = a[j0 + 1][:, i0 + 1]
        return (a00 * (1 - txx) + a01 * txx) * (1 - tyy) + (a10 * (1 - txx) + a11 * txx) * tyy

    return Gain(x, y, bil(g.n), bil(g.g0).astype(np.float32),
                None if g.E0 is None else bil(g.E0).astype(np.float32),
                bil(g.gv).astype(np.float32), bil(g.gv0).astype(np.float32))


def warp_gain_grid(g, ax=0.6, ay=-0.5):
    """Same node data on a NON-uniform grid: u -> u + a*u*(1-u) on each axis (monotone for
    |a| < 1, end points fixed).  Exercises the generic cell search and cell widths that all
    differ (the reference's findindex is a bisection, any monotone grid is legal)."""
    def warp(c, a):
        u = (c - c[0]) / (c[-1] - c[0])
        return c[0] + (c[-1] - c[0]) * (u + a * u * (1.0 - u))
    return Gain(warp(g.x, ax), warp(g.y, ay), g.n, g.g0, g.E0, g.gv, g.gv0)


def s4(small, gain_factor=2, image_factor=2):
    """Config 4: gain planes resampled to gain_factor x finer cells per axis, image nx, ny x
    image_factor, angles unchanged."""
    e = small.euv_beam

    def refine(x, dx, f):
        lo, hi = x[0] - 0.5 * dx, x[-1] + 0.5 * dx
        n = x.size * f
        d = (hi - lo) / n
        return lo + (0.5 + np.arange(n)) * d, d

    x, dx = refine(e.x, e.dx, image_factor)
    y, dy = refine(e.y, e.dy, image_factor)
    euv = BeamGrid(x, y, e.a, e.b, dx, dy, e.da, e.db, dv=e.dv, dz=e.dz, extra=e.extra)
    planes = [small.gain[0]] + [resample_gain(g, gain_factor, gain_factor) for g in small.gain[1:]]
    planes[0] = resample_gain(small.gain[0], gain_factor, gain_factor)
    return Problem(euv, planes, None, None, 0, 1)


def spectral(small, K, angle_factor=1):
    """Config 5: lineshape and dv resampled linearly to K frequency bins, renormalised so that
    sum(2*dv*gv) per cell is preserved; angles na, nb x angle_factor."""
    e = small.euv_beam
    K0 = e.nv
    pos = np.linspace(0, K0 - 1, K)
    dv = np.interp(pos, np.arange(K0), e.dv) * (K0 / K)
    planes = []
    for g in small.gain:
        k0 = np.minimum(pos.astype(int), K0 - 2)
        t = pos - k0
        gv = g.gv.astype(np.float64)
        new = gv[..., k0] * (1 - t) + gv[..., k0 + 1] * t
        old_sum = (2 * e.dv * gv).sum(-1, keepdims=True)
        new_sum = (2 * dv * new).sum(-1, keepdims=True)
        new = new * np.where(new_sum > 0, old_sum / np.where(new_sum > 0, new_sum, 1), 1.0)
        planes.append(Gain(g.x, g.y, g.n, g.g0, g.E0, new.astype(np.float32), g.gv0))

    def refine(x, dx, f):
        lo, hi = x[0] - 0.5 * dx, x[-1] + 0.5 * dx
        n = x.size * f
        d = (hi - lo) / n
        return lo + (0.5 + np.arange(n)) * d, d

    a, da = refine(e.a, e.da, angle_factor)
    b, db = refine(e.b, e.db, angle_factor)
    euv = BeamGrid(e.x, e.y, a, b, e.dx, e.dy, da, db, dv=dv, dz=e.dz, extra={})
    return Problem(euv, planes, None, None, 0, 1)
