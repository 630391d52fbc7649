"""Python doors to the parity checkers — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (raytrace-miniapp_b200) never does.

  Oracle     our plain-C restatement (oracle/rt_oracle.c -> oracle/librt_oracle.so)
  Reference  the unmodified reference compiled from /root/reference (oracle/_ref/libref_oracle.so);
             available only when that prebuilt file exists.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from raytrace_miniapp_b200 import abi  # noqa: E402

ORACLE_SO = os.path.join(_HERE, "librt_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")
REF_ROOT = "/root/reference"


def build(ref=True, quiet=True, integration=False):
    """Compile the C restatement, and the reference library when /root/reference is present.
    integration=True also links the reference's driver with the B200 back-end registered
    (CreateImage_b200, CreateImageB200) and the legacy-CUDA comparison binary; these need
    raytrace-miniapp_b200/librtb200.so to exist."""
    targets = ["oracle"]
    if ref and os.path.isdir(os.path.join(REF_ROOT, "src")):
        targets.append("ref")
        if integration:
            targets += ["b200", "legacy"]
    out = subprocess.run(["make", "-C", _HERE, "-j8"] + targets, capture_output=quiet, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n%s\n%s" % (out.stdout, out.stderr))


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.L = C.CDLL(ORACLE_SO)
        L.rt_oracle_findindex.restype = C.c_uint32
        L.rt_oracle_findindex.argtypes = [abi.c_double_p, C.c_uint32, C.c_double]
        L.rt_oracle_findfirstsingle.restype = C.c_size_t
        L.rt_oracle_findfirstsingle.argtypes = [abi.c_double_p, C.c_size_t, C.c_double]
        L.rt_oracle_bilinear.restype = C.c_float
        L.rt_oracle_bilinear.argtypes = [C.c_float] * 6
        L.rt_oracle_interp_pchip.restype = C.c_double
        L.rt_oracle_interp_pchip.argtypes = [C.c_size_t, abi.c_double_p, abi.c_double_p, C.c_double]
        L.rt_oracle_calc_seed.restype = None
        L.rt_oracle_calc_seed.argtypes = [C.POINTER(abi.Seed)] + [C.c_double] * 4 + [abi.c_double_p]
        L.rt_oracle_calc_ray.restype = C.c_int
        L.rt_oracle_calc_ray.argtypes = [C.POINTER(abi.Ray), C.c_int, C.c_float,
                                         C.POINTER(abi.GainPlane), C.POINTER(abi.Seed), C.c_int,
                                         C.c_int, C.c_float, abi.c_double_p, C.POINTER(abi.Ray),
                                         abi.c_float_p, abi.c_float_p, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        L.rt_oracle_calc_ray_debug.restype = C.c_int
        L.rt_oracle_calc_ray_debug.argtypes = [C.POINTER(abi.Ray), C.c_int, C.c_float,
                                               C.POINTER(abi.GainPlane), C.POINTER(abi.Seed),
                                               C.c_int, C.c_int, C.c_float, abi.c_double_p,
                                               abi.c_double_p, C.POINTER(abi.Ray), abi.c_float_p]
        L.rt_oracle_trace_rays.restype = None
        L.rt_oracle_trace_rays.argtypes = [C.c_int, C.POINTER(abi.Beam), C.POINTER(abi.GainPlane),
                                           C.POINTER(abi.Seed), C.c_int, C.POINTER(abi.Ray),
                                           C.c_size_t, C.c_double, abi.c_double_p, abi.c_double_p,
                                           C.POINTER(C.c_uint), C.POINTER(abi.Ray), C.c_int,
                                           C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        L.rt_oracle_create_image.restype = C.c_int
        L.rt_oracle_create_image.argtypes = [C.POINTER(abi.CProblem), C.c_uint, abi.c_double_p,
                                             abi.c_double_p, C.POINTER(C.c_uint),
                                             C.POINTER(abi.Ray), C.c_int, C.POINTER(C.c_int),
                                             C.POINTER(C.c_uint64)]
        L.rt_oracle_create_image_threads.restype = C.c_int
        L.rt_oracle_create_image_threads.argtypes = [C.POINTER(abi.CProblem), C.c_uint, C.c_int,
                                                     abi.c_double_p, abi.c_double_p,
                                                     C.POINTER(C.c_uint), C.POINTER(C.c_size_t)]
        L.rt_oracle_ray_count.restype = C.c_size_t
        L.rt_oracle_ray_count.argtypes = [C.POINTER(abi.CProblem)]

    def create_image(self, problem, flags=0, threads=0):
        """Returns dict(rc, image, I_ang, failure_code, n_failed, failed, steps)."""
        cp, keep = problem.c_struct()
        e = problem.euv_beam
        image = np.zeros(e.nx * e.ny * e.nv)
        I_ang = np.zeros(e.na * e.nb)
        fc = C.c_uint(0)
        if threads and threads > 1:
            done = C.c_size_t(0)
            rc = self.L.rt_oracle_create_image_threads(C.byref(cp), flags, threads,
                                                       _ptr(image, C.c_double), _ptr(I_ang, C.c_double),
                                                       C.byref(fc), C.byref(done))
            return dict(rc=rc, image=image, I_ang=I_ang, failure_code=fc.value, n_rays=done.value)
        failed = np.zeros(abi.N_FAILED_MAX, abi.ray_dtype)
        nf = C.c_int(0)
        steps = C.c_uint64(0)
        rc = self.L.rt_oracle_create_image(C.byref(cp), flags, _ptr(image, C.c_double),
                                           _ptr(I_ang, C.c_double), C.byref(fc),
                                           _ptr(failed, abi.Ray), abi.N_FAILED_MAX, C.byref(nf),
                                           C.byref(steps))
        return dict(rc=rc, image=image, I_ang=I_ang, failure_code=fc.value, n_failed=nf.value,
                    failed=failed[:min(nf.value, abi.N_FAILED_MAX)], steps=steps.value)

    def trace_rays(self, problem, rays, method, scale, image=None, I_ang=None):
        e = problem.euv_beam
        eb = e.c_struct()
        planes = (abi.GainPlane * problem.N)(*[g.c_struct() for g in problem.gain])
        sd = problem.seed.c_struct() if problem.seed is not None else None
        rays = np.ascontiguousarray(rays, abi.ray_dtype)
        image = np.zeros(e.nx * e.ny * e.nv) if image is None else image
        I_ang = np.zeros(e.na * e.nb) if I_ang is None else I_ang
        fc, nf, steps = C.c_uint(0), C.c_int(0), C.c_uint64(0)
        failed = np.zeros(abi.N_FAILED_MAX, abi.ray_dtype)
        self.L.rt_oracle_trace_rays(problem.N, C.byref(eb), planes, C.byref(sd) if sd else None,
                                    method, _ptr(rays, abi.Ray), rays.size, scale,
                                    _ptr(image, C.c_double), _ptr(I_ang, C.c_double), C.byref(fc),
                                    _ptr(failed, abi.Ray), abi.N_FAILED_MAX, C.byref(nf),
                                    C.byref(steps))
        return dict(image=image, I_ang=I_ang, failure_code=fc.value, n_failed=nf.value,
                    failed=failed[:min(nf.value, abi.N_FAILED_MAX)], steps=steps.value)

    def calc_rays(self, problem, rays, method=None, K=None):
        """Per-ray outputs incl. the march intermediates gvl/evl/ivl ([n, (N-1)*3])."""
        rays = np.ascontiguousarray(rays, abi.ray_dtype)
        n, N = rays.size, problem.N
        K = problem.euv_beam.nv if K is None else K
        method = problem.method if method is None else method
        S = (N - 1) * abi.N_SUB
        planes = (abi.GainPlane * N)(*[g.c_struct() for g in problem.gain])
        sd = problem.seed.c_struct() if problem.seed is not None else None
        Iv = np.zeros((n, K))
        ray2 = np.zeros(n, abi.ray_dtype)
        err = np.zeros(n, np.int32)
        gvl = np.zeros((n, S), np.float32)
        evl = np.zeros((n, S), np.float32)
        ivl = np.zeros((n, S), np.int32)
        esc = np.zeros(n, np.int32)
        steps = C.c_uint64(0)
        rp = _ptr(rays, abi.Ray)
        r2 = _ptr(ray2, abi.Ray)
        for i in range(n):
            e_ = C.c_int(0)
            err[i] = self.L.rt_oracle_calc_ray(
                C.byref(rp[i]), N, np.float32(problem.euv_beam.dz), planes,
                C.byref(sd) if sd else None, K, method, 0.5, _ptr(Iv[i], C.c_double),
                C.byref(r2[i]), _ptr(gvl[i], C.c_float), _ptr(evl[i], C.c_float),
                _ptr(ivl[i], C.c_int32), C.byref(e_), C.byref(steps))
            esc[i] = e_.value
        return dict(Iv=Iv, ray2=ray2, error=err, gvl=gvl, evl=evl, ivl=ivl, escaped=esc,
                    steps=steps.value)


    def calc_ray_paths(self, problem, rays, method=None, c=0.5):
        """RAY_DEBUG trajectories: (x, y, I)[n_rays, N_SUB*(N-1)+1] and the error codes."""
        rays = np.ascontiguousarray(rays, abi.ray_dtype)
        n, N, K = rays.size, problem.N, problem.euv_beam.nv
        method = problem.method if method is None else method
        planes = (abi.GainPlane * N)(*[g.c_struct() for g in problem.gain])
        sd = problem.seed.c_struct() if problem.seed is not None else None
        N2 = abi.N_SUB * (N - 1) + 1
        dbg = np.zeros((n, N2, 3), np.float32)
        err = np.zeros(n, np.int32)
        Iv = np.zeros(K)
        r2 = abi.Ray()
        rp = _ptr(rays, abi.Ray)
        dv = np.ascontiguousarray(problem.euv_beam.dv)
        for i in range(n):
            err[i] = self.L.rt_oracle_calc_ray_debug(
                C.byref(rp[i]), N, np.float32(problem.euv_beam.dz), planes,
                C.byref(sd) if sd else None, K, method, c, _ptr(dv, C.c_double),
                _ptr(Iv, C.c_double), C.byref(r2), _ptr(dbg[i], C.c_float))
        return dict(x=dbg[:, :, 0].copy(), y=dbg[:, :, 1].copy(), I=dbg[:, :, 2].copy(), error=err)


class Reference:
    """The unmodified reference (oracle/_ref/libref_oracle.so) on a .dat file."""

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def __init__(self, dat_path):
        if not self.available():
            raise RuntimeError("oracle/_ref/libref_oracle.so not built (needs /root/reference)")
        L = self.L = C.CDLL(REF_SO)
        L.ref_load.restype = C.c_void_p
        L.ref_load.argtypes = [C.c_char_p]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.ref_golden.argtypes = [C.c_void_p, abi.c_double_p, abi.c_double_p]
        L.ref_create_image.restype = C.c_double
        L.ref_create_image.argtypes = [C.c_void_p, C.c_char_p, abi.c_double_p, abi.c_double_p]
        L.ref_calc_rays.argtypes = [C.c_void_p, C.c_int, abi.c_float_p, C.c_int, abi.c_double_p,
                                    abi.c_float_p, C.POINTER(C.c_int), abi.c_float_p]
        L.ref_findindex.restype = C.c_uint
        L.ref_findindex.argtypes = [abi.c_double_p, C.c_uint, C.c_double]
        L.ref_findfirstsingle.restype = C.c_size_t
        L.ref_findfirstsingle.argtypes = [abi.c_double_p, C.c_size_t, C.c_double]
        L.ref_bilinear.restype = C.c_float
        L.ref_bilinear.argtypes = [C.c_float] * 6
        L.ref_interp_pchip.restype = C.c_double
        L.ref_interp_pchip.argtypes = [C.c_size_t, abi.c_double_p, abi.c_double_p, C.c_double]
        L.ref_calc_seed.argtypes = [C.c_void_p] + [C.c_double] * 4 + [abi.c_double_p]
        L.ref_hardware_threads.restype = C.c_int
        self.h = L.ref_load(dat_path.encode())
        if not self.h:
            raise RuntimeError("reference could not load %s" % dat_path)
        d = (C.c_int * 11)()
        L.ref_info(self.h, d)
        (self.N, self.N_start, self.N_parallel, self.nx, self.ny, self.na, self.nb, self.nv,
         self.has_seed, self.has_gimg, self.has_gang) = list(d)

    def close(self):
        if self.h:
            self.L.ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def golden(self):
        img = np.zeros(self.nx * self.ny * self.nv)
        ang = np.zeros(self.na * self.nb)
        self.L.ref_golden(self.h, _ptr(img, C.c_double), _ptr(ang, C.c_double))
        return img, ang

    def create_image(self, method="cpu"):
        img = np.zeros(self.nx * self.ny * self.nv)
        ang = np.zeros(self.na * self.nb)
        sec = self.L.ref_create_image(self.h, method.encode(), _ptr(img, C.c_double),
                                      _ptr(ang, C.c_double))
        return img, ang, sec

    def calc_rays(self, rays, method, debug=False):
        rays = np.ascontiguousarray(rays, abi.ray_dtype)
        n = rays.size
        Iv = np.zeros((n, self.nv))
        ray2 = np.zeros(n, abi.ray_dtype)
        err = np.zeros(n, np.int32)
        dbg = np.zeros((n, 3 * (abi.N_SUB * (self.N - 1) + 1)), np.float32) if debug else None
        self.L.ref_calc_rays(self.h, method, rays.view(np.float32).ctypes.data_as(abi.c_float_p), n,
                             _ptr(Iv, C.c_double), ray2.view(np.float32).ctypes.data_as(abi.c_float_p),
                             _ptr(err, C.c_int), _ptr(dbg, C.c_float) if debug else None)
        return dict(Iv=Iv, ray2=ray2, error=err, debug=dbg)

    def hardware_threads(self):
        return self.L.ref_hardware_threads()
