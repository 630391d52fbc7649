"""Compact .npz container for problems (fixtures under tests/golden/, synthetic inputs)."""
import numpy as np

from .abi import BeamGrid, Gain, Problem, SeedProfile


def _beam_dict(prefix, g):
    d = {prefix + k: getattr(g, k) for k in ("x", "y", "a", "b")}
    d[prefix + "d"] = np.array([g.dx, g.dy, g.da, g.db, g.dz])
    if g.dv.size:
        d[prefix + "dv"] = g.dv
    return d


def _beam_from(prefix, z):
    d = z[prefix + "d"]
    dv = z[prefix + "dv"] if prefix + "dv" in z else None
    return BeamGrid(z[prefix + "x"], z[prefix + "y"], z[prefix + "a"], z[prefix + "b"], d[0], d[1],
                    d[2], d[3], dv=dv, dz=d[4])


def problem_arrays(p):
    d = {"N": np.array([p.N, p.N_start, p.N_parallel])}
    d.update(_beam_dict("euv_", p.euv_beam))
    if p.seed_beam is not None:
        d.update(_beam_dict("sb_", p.seed_beam))
    for i, g in enumerate(p.gain):
        d.update({"g%d_x" % i: g.x, "g%d_y" % i: g.y, "g%d_n" % i: g.n, "g%d_g0" % i: g.g0,
                  "g%d_gv" % i: g.gv})
        if g.E0 is not None:
            d["g%d_E0" % i] = g.E0
    if p.seed is not None:
        for i in range(5):
            d["seed_x%d" % i] = p.seed.x[i]
            d["seed_f%d" % i] = p.seed.f[i]
        d["seed_scale"] = np.array([p.seed.f0])
    return d


def save_npz(path, p, **extra):
    d = problem_arrays(p)
    d.update(extra)
    np.savez_compressed(path, **d)


def load_npz(path):
    """Returns (Problem, dict of the extra arrays)."""
    z = dict(np.load(path))
    N, N_start, N_parallel = [int(v) for v in z.pop("N")]
    euv = _beam_from("euv_", z)
    sb = _beam_from("sb_", z) if "sb_x" in z else None
    gain = []
    for i in range(N):
        gain.append(Gain(z["g%d_x" % i], z["g%d_y" % i], z["g%d_n" % i], z["g%d_g0" % i],
                         z.get("g%d_E0" % i), z["g%d_gv" % i]))
    seed = None
    if "seed_scale" in z:
        seed = SeedProfile([z["seed_x%d" % i] for i in range(5)], [z["seed_f%d" % i] for i in range(5)],
                           float(z["seed_scale"][0]))
    used = ("euv_", "sb_", "seed_") + tuple("g%d_" % i for i in range(N))
    extra = {k: v for k, v in z.items() if not k.startswith(used)}
    return Problem(euv, gain, sb, seed, N_start, N_parallel), extra
