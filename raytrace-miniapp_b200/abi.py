"""ctypes mirror of include/rtb200.h plus a numpy-backed problem container.

The same `Problem` object drives the product library (librtb200.so, CUDA only) and, in the
tests, the CPU oracle (oracle/librt_oracle.so): both take the POD descriptors of
include/rtb200.h, which flatten the reference's structs (src/RayTraceStructures.h:26-56,
:218-230, :276-281, :321-338).
"""
import ctypes as C

import numpy as np

N_MAX, K_MAX, N_SUB, N_FAILED_MAX = 20, 100, 3, 32
OK, RAYS_FAILED = 0, 1
ERR_LIMITS, ERR_GRID, ERR_CUDA, ERR_ARG, ERR_FORMAT = -1, -2, -3, -4, -5
FLAG_NO_LIMITS = 0x1
FLAG_LAZY_TABLES = 0x2

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)


class Ray(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("a", C.c_float), ("b", C.c_float)]


ray_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("a", "<f4"), ("b", "<f4")])


class Beam(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("na", C.c_int32), ("nb", C.c_int32),
                ("nv", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double), ("da", C.c_double), ("db", C.c_double),
                ("dz", C.c_double),
                ("x", c_double_p), ("y", c_double_p), ("a", c_double_p), ("b", c_double_p),
                ("dv", c_double_p)]


class GainPlane(C.Structure):
    _fields_ = [("Nx", C.c_int32), ("Ny", C.c_int32), ("Nv", C.c_int32),
                ("x", c_double_p), ("y", c_double_p), ("n", c_double_p),
                ("g0", c_float_p), ("E0", c_float_p), ("gv", c_float_p)]


class Seed(C.Structure):
    _fields_ = [("dim", C.c_int32 * 5), ("x", c_double_p * 5), ("f", c_double_p * 5),
                ("f0", C.c_double)]


class CProblem(C.Structure):
    _fields_ = [("N", C.c_int32), ("N_start", C.c_int32), ("N_parallel", C.c_int32),
                ("euv_beam", C.POINTER(Beam)), ("seed_beam", C.POINTER(Beam)),
                ("gain", C.POINTER(GainPlane)), ("seed", C.POINTER(Seed))]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("march_ms", C.c_float), ("integrate_ms", C.c_float),
                ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("kernel_launches", C.c_int32),
                ("reserved", C.c_int32), ("n_rays", C.c_uint64), ("march_steps", C.c_uint64)]


class Staged(C.Structure):
    _fields_ = [("method", C.c_int32), ("owner", C.c_int32), ("snx", C.c_int32), ("sny", C.c_int32),
                ("nx", C.c_int32), ("ny", C.c_int32), ("na", C.c_int32), ("nb", C.c_int32),
                ("nv", C.c_int32), ("reserved", C.c_int32)]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


class BeamGrid:
    """Output (euv) beam or seed-beam grid: cell centres x, y, a, b (+ dv, dz for the euv beam)."""

    def __init__(self, x, y, a, b, dx, dy, da, db, dv=None, dz=0.0, extra=None):
        self.x, self.y, self.a, self.b = _f64(x), _f64(y), _f64(a), _f64(b)
        self.dx, self.dy, self.da, self.db, self.dz = float(dx), float(dy), float(da), float(db), float(dz)
        self.dv = _f64(dv) if dv is not None else np.zeros(0)
        self.extra = dict(extra or {})  # fields of the wire format that are off the path

    nx = property(lambda s: s.x.size)
    ny = property(lambda s: s.y.size)
    na = property(lambda s: s.a.size)
    nb = property(lambda s: s.b.size)
    nv = property(lambda s: s.dv.size)

    def c_struct(self):
        return Beam(self.nx, self.ny, self.na, self.nb, self.nv, self.dx, self.dy, self.da,
                    self.db, self.dz, _dp(self.x), _dp(self.y), _dp(self.a), _dp(self.b),
                    _dp(self.dv) if self.dv.size else None)


class Gain:
    """One length plane of the gain medium (ray_gain_struct)."""

    def __init__(self, x, y, n, g0, E0, gv, gv0=None):
        self.x, self.y = _f64(x), _f64(y)
        Nx, Ny = self.x.size, self.y.size
        self.n = _f64(n).reshape(Ny, Nx)
        self.g0 = _f32(g0).reshape(Ny, Nx)
        self.E0 = None if E0 is None else _f32(E0).reshape(Ny, Nx)
        self.gv = _f32(gv).reshape(Ny, Nx, -1)
        self.gv0 = np.zeros((Ny, Nx), np.float32) if gv0 is None else _f32(gv0).reshape(Ny, Nx)

    Nx = property(lambda s: s.x.size)
    Ny = property(lambda s: s.y.size)
    Nv = property(lambda s: s.gv.shape[2])

    def c_struct(self):
        return GainPlane(self.Nx, self.Ny, self.Nv, _dp(self.x), _dp(self.y), _dp(self.n),
                         _fp(self.g0), _fp(self.E0) if self.E0 is not None else None,
                         _fp(self.gv))


class SeedProfile:
    """Separable seed beam f0*fx(x)*fy(y)*fa(a)*fb(b)*fv[k] (ray_seed_struct)."""

    def __init__(self, x, f, f0):
        self.x = [_f64(v) for v in x]
        self.f = [_f64(v) for v in f]
        self.f0 = float(f0)

    def c_struct(self):
        s = Seed()
        for i in range(5):
            s.dim[i] = self.x[i].size
            s.x[i] = _dp(self.x[i])
            s.f[i] = _dp(self.f[i])
        s.f0 = self.f0
        return s


class Problem:
    """create_image_struct: everything one create_image call consumes."""

    def __init__(self, euv_beam, gain, seed_beam=None, seed=None, N_start=0, N_parallel=1):
        self.euv_beam, self.gain, self.seed_beam, self.seed = euv_beam, list(gain), seed_beam, seed
        self.N_start, self.N_parallel = int(N_start), int(N_parallel)

    N = property(lambda s: len(s.gain))

    @property
    def ray_grid(self):
        return self.seed_beam if self.seed is not None else self.euv_beam

    @property
    def method(self):
        return 2 if self.seed is not None else 1

    @property
    def n_rays_total(self):
        g = self.ray_grid
        return g.nx * g.ny * g.na * g.nb

    @property
    def n_rays(self):
        """Rays of this worker: ijkm = N_start + it*N_parallel < Nt (src/RayTraceImage.cpp:300-308)."""
        Nt, s, o = self.n_rays_total, self.N_parallel, self.N_start
        return len(range(o, Nt, s)) if o < Nt else 0

    @property
    def ray_segments(self):
        """Unit of work of BASELINE.json's metric: rays * (N-1) * N_SUB."""
        return self.n_rays * (self.N - 1) * N_SUB

    def c_struct(self):
        """Returns (CProblem, keepalive).  Keep `keepalive` referenced while the struct is used."""
        eb = self.euv_beam.c_struct()
        sb = self.seed_beam.c_struct() if self.seed_beam is not None else None
        planes = (GainPlane * self.N)(*[g.c_struct() for g in self.gain])
        sd = self.seed.c_struct() if self.seed is not None else None
        p = CProblem(self.N, self.N_start, self.N_parallel, C.pointer(eb),
                     C.pointer(sb) if sb is not None else None, planes,
                     C.pointer(sd) if sd is not None else None)
        return p, (eb, sb, planes, sd, self)

    def marshal(self):
        """The C structure built ONCE (what an application that holds a create_image_struct passes
        to every call): a view of this problem whose c_struct() costs nothing.  The arrays and the
        N_start / N_parallel of the moment are captured; marshal again after changing them."""
        return Marshalled(self)

    def rays(self):
        """The ray list create_image builds (src/RayTraceImage.cpp:300-328), as a structured array."""
        g = self.ray_grid
        c = self.seed_beam if self.seed_beam is not None else self.euv_beam
        ijkm = np.arange(self.N_start, self.n_rays_total, self.N_parallel, dtype=np.int64)
        m = ijkm % g.nb
        k = (ijkm // g.nb) % g.na
        j = (ijkm // (g.na * g.nb)) % g.ny
        i = ijkm // (g.ny * g.na * g.nb)
        r = np.empty(ijkm.size, ray_dtype)
        r["x"], r["y"], r["a"], r["b"] = c.x[i], c.y[j], c.a[k], c.b[m]
        return r


class Marshalled:
    """Problem + its rtb200_problem structure (Problem.marshal); accepted wherever a Problem is."""

    def __init__(self, problem):
        self.problem = problem
        self._c = problem.c_struct()

    def c_struct(self):
        return self._c

    def __getattr__(self, name):
        return getattr(self.problem, name)
