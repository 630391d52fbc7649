"""ctypes binding of librtb200.so — the host-side mirror of the reference's interface.

`Context.create_image(problem)` is the call a user of RayTrace::create_image(info, "b200")
makes (src/RayTrace.h:93, src/RayTraceImage.cpp:227-434): host arrays in, host image / I_ang
out, failures reported the reference's way.  Everything computes on the GPU through the C ABI
of include/rtb200.h; there is no CPU path here, and a missing library or device raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_LIB = None


class RTB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rtb200 error %d: %s" % (code, msg))
        self.code = code


class RaysFailed(RuntimeError):
    """The reference aborts with "Some rays failed" (src/RayTraceImage.cpp:427-430)."""

    def __init__(self, failure_code, failed):
        msgs = []
        if failure_code & 2:
            msgs.append("Invalid ray detected")
        if failure_code & 4:
            msgs.append("Negitive intensity detected")
        if failure_code & 8:
            msgs.append("NaNs detected in intensity")
        super().__init__("Some rays failed: " + "; ".join(msgs))
        self.failure_code, self.failed = failure_code, failed


def library_path():
    """In-tree librtb200.so; RTB200_LIB selects another build of the same library (A/B tuning)."""
    return os.environ.get("RTB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                        "librtb200.so")


def load():
    """Load librtb200.so (never builds implicitly, never falls back)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `python __graft_entry__.py build` (the rtb200 path "
                           "has no CPU fallback)" % path)
    L = C.CDLL(path)
    ctx = C.c_void_p
    P = C.POINTER
    L.rtb200_version.restype = C.c_char_p
    L.rtb200_device_count.restype = C.c_int
    L.rtb200_create.argtypes = [C.c_int, P(ctx)]
    L.rtb200_destroy.argtypes = [ctx]
    L.rtb200_destroy.restype = None
    L.rtb200_last_error.argtypes = [ctx]
    L.rtb200_last_error.restype = C.c_char_p
    L.rtb200_create_image.argtypes = [ctx, P(abi.CProblem), C.c_uint, C.c_void_p, C.c_void_p,
                                      P(C.c_uint), P(abi.Ray), C.c_int, P(C.c_int)]
    L.rtb200_trace_rays.argtypes = [ctx, C.c_int, P(abi.Beam), P(abi.GainPlane), P(abi.Seed),
                                    C.c_int, P(abi.Ray), C.c_size_t, C.c_double, C.c_void_p,
                                    C.c_void_p, P(C.c_uint), P(abi.Ray), C.c_int, P(C.c_int)]
    L.rtb200_calc_rays.argtypes = [ctx, C.c_int, C.c_double, P(abi.GainPlane), P(abi.Seed),
                                   C.c_int, C.c_int, P(abi.Ray), C.c_size_t, C.c_void_p,
                                   P(abi.Ray), P(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
    L.rtb200_calc_ray_paths.argtypes = [ctx, C.c_int, C.c_double, P(abi.GainPlane), P(abi.Seed),
                                        C.c_int, abi.c_double_p, C.c_int, C.c_double, P(abi.Ray),
                                        C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, P(C.c_int)]
    L.rtb200_stage.argtypes = [ctx, P(abi.CProblem), C.c_uint]
    L.rtb200_staged_pixels.argtypes = [ctx]
    L.rtb200_staged_pixels.restype = C.c_int64
    L.rtb200_staged_rays.argtypes = [ctx]
    L.rtb200_staged_rays.restype = C.c_int64
    L.rtb200_launch.argtypes = [ctx, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.rtb200_launch_rows.argtypes = [ctx, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.rtb200_launch_rows_compact.argtypes = [ctx, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.rtb200_unpermute_rows.argtypes = [ctx, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.rtb200_staged_info.argtypes = [ctx, P(abi.Staged)]
    L.rtb200_multi_create.argtypes = [P(C.c_int), C.c_int, P(ctx)]
    L.rtb200_multi_destroy.argtypes = [ctx]
    L.rtb200_multi_destroy.restype = None
    L.rtb200_multi_last_error.argtypes = [ctx]
    L.rtb200_multi_last_error.restype = C.c_char_p
    L.rtb200_multi_device_count.argtypes = [ctx]
    L.rtb200_multi_create_image.argtypes = [ctx, P(abi.CProblem), C.c_uint, C.c_void_p, C.c_void_p,
                                            P(C.c_uint), P(abi.Ray), C.c_int, P(C.c_int)]
    L.rtb200_multi_get_timings.argtypes = [ctx, P(abi.Timings), P(C.c_float), P(C.c_float)]
    L.rtb200_sync.argtypes = [ctx, P(C.c_uint), P(abi.Ray), C.c_int, P(C.c_int)]
    L.rtb200_get_timings.argtypes = [ctx, P(abi.Timings)]
    L.rtb200_reset_timings.argtypes = [ctx]
    L.rtb200_parse_dat.argtypes = [C.c_void_p, C.c_size_t, P(P(abi.CProblem)),
                                   P(abi.c_double_p), P(abi.c_double_p)]
    L.rtb200_free_problem.argtypes = [P(abi.CProblem)]
    L.rtb200_free_problem.restype = None
    L.rtb200_write_dat.argtypes = [P(abi.CProblem), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   P(C.c_size_t)]
    L.rtb200_create_image_from_dat.argtypes = [ctx, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p,
                                               C.c_void_p, P(C.c_uint), P(abi.Ray), C.c_int, P(C.c_int)]
    L.rtb200_measure_fp64_peak.argtypes = [ctx, P(C.c_double)]
    L.rtb200_check_fdiv.argtypes = [ctx, C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_int, P(C.c_ulonglong),
                                    P(C.c_float), P(C.c_float)]
    _LIB = L
    return L


def device_count():
    return load().rtb200_device_count()


def write_dat_payload(problem, image=None, I_ang=None):
    """rtb200_write_dat (the C++ writer of the reference's wire format): the payload of a .dat
    file as bytes; a file is struct.pack("<Q", len(payload)) + payload."""
    L = load()
    cp, keep = problem.c_struct()
    img = None if image is None else np.ascontiguousarray(image, np.float64)
    ang = None if I_ang is None else np.ascontiguousarray(I_ang, np.float64)
    n = C.c_size_t(0)
    rc = L.rtb200_write_dat(C.byref(cp), _addr(img), _addr(ang), None, 0, C.byref(n))
    if rc != abi.OK:
        raise RTB200Error(rc, "rtb200_write_dat")
    buf = np.empty(n.value, np.uint8)
    rc = L.rtb200_write_dat(C.byref(cp), _addr(img), _addr(ang), buf.ctypes.data, buf.size, C.byref(n))
    if rc != abi.OK:
        raise RTB200Error(rc, "rtb200_write_dat")
    return buf.tobytes()


def _check_out(a, n, name):
    """An output buffer handed to the C ABI must be n contiguous float64 values."""
    if a is None or isinstance(a, int):
        return
    if hasattr(a, "data_ptr"):  # torch tensor
        import torch
        ok = a.dtype == torch.float64 and a.is_contiguous() and a.numel() == n
    else:
        ok = a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.size == n
    if not ok:
        raise ValueError("%s must be a contiguous float64 buffer of %d elements" % (name, n))


def _addr(a):
    """Host numpy array / torch tensor (host or device) / int -> raw address."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


class Context:
    """One CUDA device + stream + staging arena (rtb200_ctx).  Not shared between threads."""

    def __init__(self, device=0):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.rtb200_create(device, C.byref(self.h))
        if rc != abi.OK:
            raise RTB200Error(rc, "no usable CUDA device %d (the rtb200 path has no CPU fallback)" % device)
        self.device = device
        self._keep = None

    def close(self):
        if self.h:
            self.L.rtb200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise RTB200Error(rc, self.L.rtb200_last_error(self.h).decode())
        return rc

    # ---- reference-facing calls (host buffers) -------------------------------------------------
    def create_image(self, problem, flags=0, image=None, I_ang=None, raise_on_failed=True):
        """RayTrace::create_image(info, "b200").  Returns (image[ny,nx,nv], I_ang[nb,na]) as flat
        arrays in the reference's layout: image[nv*(i + j*nx) + k], I_ang[ia + ib*na]."""
        e = problem.euv_beam
        cp, keep = problem.c_struct()
        image = np.empty(e.nx * e.ny * e.nv) if image is None else image
        I_ang = np.empty(e.na * e.nb) if I_ang is None else I_ang
        _check_out(image, e.nx * e.ny * e.nv, "image")
        _check_out(I_ang, e.na * e.nb, "I_ang")
        fc, nf = C.c_uint(0), C.c_int(0)
        failed = np.zeros(abi.N_FAILED_MAX, abi.ray_dtype)
        rc = self._check(self.L.rtb200_create_image(
            self.h, C.byref(cp), flags, _addr(image), _addr(I_ang), C.byref(fc),
            failed.ctypes.data_as(C.POINTER(abi.Ray)), abi.N_FAILED_MAX, C.byref(nf)))
        self.failure_code, self.n_failed = fc.value, nf.value
        self.failed = failed[:min(nf.value, abi.N_FAILED_MAX)]
        if rc == abi.RAYS_FAILED and raise_on_failed:
            raise RaysFailed(fc.value, self.failed)
        return image, I_ang

    def create_image_from_dat(self, payload, n_image, n_ang, flags=0, raise_on_failed=True):
        """rtb200_create_image_from_dat: the serialized create_image_struct (payload of a .dat file)
        straight to the device.  n_image / n_ang: sizes of the outputs (nx*ny*nv, na*nb)."""
        buf = np.frombuffer(payload, np.uint8)
        image, I_ang = np.empty(n_image), np.empty(n_ang)
        fc, nf = C.c_uint(0), C.c_int(0)
        failed = np.zeros(abi.N_FAILED_MAX, abi.ray_dtype)
        rc = self._check(self.L.rtb200_create_image_from_dat(
            self.h, buf.ctypes.data, buf.size, flags, _addr(image), _addr(I_ang), C.byref(fc),
            failed.ctypes.data_as(C.POINTER(abi.Ray)), abi.N_FAILED_MAX, C.byref(nf)))
        self.failure_code, self.n_failed = fc.value, nf.value
        if rc == abi.RAYS_FAILED and raise_on_failed:
            raise RaysFailed(fc.value, failed[:min(nf.value, abi.N_FAILED_MAX)])
        return image, I_ang

    def trace_rays(self, problem, rays, method, scale, image=None, I_ang=None):
        """RayTraceImage<B200>Loop: explicit ray list, accumulates into image / I_ang."""
        e = problem.euv_beam
        eb = e.c_struct()
        planes = (abi.GainPlane * problem.N)(*[g.c_struct() for g in problem.gain])
        sd = problem.seed.c_struct() if problem.seed is not None else None
        rays = np.ascontiguousarray(rays, abi.ray_dtype)
        image = np.zeros(e.nx * e.ny * e.nv) if image is None else image
        I_ang = np.zeros(e.na * e.nb) if I_ang is None else I_ang
        _check_out(image, e.nx * e.ny * e.nv, "image")
        _check_out(I_ang, e.na * e.nb, "I_ang")
        fc, nf = C.c_uint(0), C.c_int(0)
        failed = np.zeros(abi.N_FAILED_MAX, abi.ray_dtype)
        self._check(self.L.rtb200_trace_rays(
            self.h, problem.N, C.byref(eb), planes, C.byref(sd) if sd else None, method,
            rays.ctypes.data_as(C.POINTER(abi.Ray)), rays.size, scale, _addr(image), _addr(I_ang),
            C.byref(fc), failed.ctypes.data_as(C.POINTER(abi.Ray)), abi.N_FAILED_MAX, C.byref(nf)))
        self.failure_code, self.n_failed = fc.value, nf.value
        self.failed = failed[:min(nf.value, abi.N_FAILED_MAX)]
        return image, I_ang

    def calc_rays(self, problem, rays, method=None, K=None):
        """RayTrace::calc_ray for a batch, plus the march intermediates."""
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
        self._check(self.L.rtb200_calc_rays(
            self.h, N, problem.euv_beam.dz, planes, C.byref(sd) if sd else None, K, method,
            rays.ctypes.data_as(C.POINTER(abi.Ray)), n, _addr(Iv),
            ray2.ctypes.data_as(C.POINTER(abi.Ray)), err.ctypes.data_as(C.POINTER(C.c_int)),
            _addr(gvl), _addr(evl), _addr(ivl)))
        return dict(Iv=Iv, ray2=ray2, error=err, gvl=gvl, evl=evl, ivl=ivl)

    def calc_ray_paths(self, problem, rays, method=None, c=0.5):
        """RayTrace::calc_ray_path for a ray list: x, y, I [n_rays, 3*(N-1)+1] and error codes."""
        rays = np.ascontiguousarray(rays, abi.ray_dtype)
        n, N = rays.size, problem.N
        method = problem.method if method is None else method
        N2 = abi.N_SUB * (N - 1) + 1
        planes = (abi.GainPlane * N)(*[g.c_struct() for g in problem.gain])
        sd = problem.seed.c_struct() if problem.seed is not None else None
        dv = np.ascontiguousarray(problem.euv_beam.dv, np.float64)
        x, y, I = (np.zeros((n, N2), np.float32) for _ in range(3))
        err = np.zeros(n, np.int32)
        self._check(self.L.rtb200_calc_ray_paths(
            self.h, N, problem.euv_beam.dz, planes, C.byref(sd) if sd else None, dv.size,
            dv.ctypes.data_as(abi.c_double_p), method, c, rays.ctypes.data_as(C.POINTER(abi.Ray)), n,
            _addr(x), _addr(y), _addr(I), err.ctypes.data_as(C.POINTER(C.c_int))))
        return dict(x=x, y=y, I=I, error=err)

    # ---- device-resident calls -----------------------------------------------------------------
    def stage(self, problem, flags=0):
        cp, keep = problem.c_struct()
        self._check(self.L.rtb200_stage(self.h, C.byref(cp), flags))
        self._keep = (cp, keep)  # with FLAG_LAZY_TABLES the library reads them again at the launch
        return self.L.rtb200_staged_pixels(self.h)

    @property
    def staged_pixels(self):
        return self.L.rtb200_staged_pixels(self.h)

    @property
    def staged_rays(self):
        return self.L.rtb200_staged_rays(self.h)

    def launch(self, pix_begin, pix_end, d_image, d_I_ang, stream=None):
        """d_image / d_I_ang: device buffers (torch CUDA tensors or raw addresses).  stream:
        None = the context's own stream; else a cudaStream_t handle (0, torch's handle of the
        legacy default stream, is passed as cudaStreamLegacy)."""
        if stream is not None and stream == 0:
            stream = 1  # cudaStreamLegacy
        self._check(self.L.rtb200_launch(self.h, pix_begin, pix_end, _addr(d_image),
                                         _addr(d_I_ang), stream))

    def launch_rows(self, row_offset, row_stride, d_image, d_I_ang, stream=None):
        """Trace the image rows row_offset, row_offset + row_stride, ... (multi-GPU sharding)."""
        if stream is not None and stream == 0:
            stream = 1  # cudaStreamLegacy
        self._check(self.L.rtb200_launch_rows(self.h, row_offset, row_stride, _addr(d_image),
                                              _addr(d_I_ang), stream))

    def launch_rows_compact(self, row_offset, row_stride, d_rows, d_I_ang, stream=None):
        """The same share written compactly (its rows back to back): what a gather moves."""
        if stream is not None and stream == 0:
            stream = 1  # cudaStreamLegacy
        self._check(self.L.rtb200_launch_rows_compact(self.h, row_offset, row_stride, _addr(d_rows),
                                                      _addr(d_I_ang), stream))

    def unpermute_rows(self, d_gathered, world, rows_per_dev, d_image, stream=None):
        """Gathered compact rows of `world` devices -> full image (zeroed by the caller)."""
        if stream is not None and stream == 0:
            stream = 1
        self._check(self.L.rtb200_unpermute_rows(self.h, _addr(d_gathered), world, rows_per_dev,
                                                 _addr(d_image), stream))

    def staged_info(self):
        s = abi.Staged()
        self._check(self.L.rtb200_staged_info(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in abi.Staged._fields_ if k != "reserved"}

    def sync(self, raise_on_failed=True):
        fc, nf = C.c_uint(0), C.c_int(0)
        failed = np.zeros(abi.N_FAILED_MAX, abi.ray_dtype)
        rc = self._check(self.L.rtb200_sync(self.h, C.byref(fc),
                                            failed.ctypes.data_as(C.POINTER(abi.Ray)),
                                            abi.N_FAILED_MAX, C.byref(nf)))
        self.failure_code, self.n_failed = fc.value, nf.value
        self.failed = failed[:min(nf.value, abi.N_FAILED_MAX)]
        if rc == abi.RAYS_FAILED and raise_on_failed:
            raise RaysFailed(fc.value, self.failed)
        return rc

    def reset_timings(self):
        self._check(self.L.rtb200_reset_timings(self.h))

    def timings(self):
        t = abi.Timings()
        self._check(self.L.rtb200_get_timings(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in abi.Timings._fields_ if k != "reserved"}

    def check_fdiv(self, b_first, b_count, exp_a=0, exp_b=0, variant=0):
        """Mismatches of the branch-free FP32 division against IEEE over b_count divisor
        significands x all 2^23 numerator significands; returns (count, a, b)."""
        n, a, b = C.c_ulonglong(0), C.c_float(0), C.c_float(0)
        self._check(self.L.rtb200_check_fdiv(self.h, b_first, b_count, exp_a, exp_b, variant, C.byref(n),
                                             C.byref(a), C.byref(b)))
        return n.value, a.value, b.value

    def measure_fp64_peak(self):
        r = C.c_double(0)
        self._check(self.L.rtb200_measure_fp64_peak(self.h, C.byref(r)))
        return r.value


class MultiContext:
    """Several devices of one box behind one call (rtb200_multi): the reference's `cuda-multigpu`
    method (src/RayTraceImage.cpp:396-405) with the partial results exchanged over NCCL."""

    def __init__(self, devices):
        self.L = load()
        self.h = C.c_void_p()
        devs = list(range(devices)) if isinstance(devices, int) else list(devices)
        arr = (C.c_int * len(devs))(*devs)
        rc = self.L.rtb200_multi_create(arr, len(devs), C.byref(self.h))
        if rc != abi.OK:
            msg = self.L.rtb200_multi_last_error(self.h).decode() if self.h else "no usable CUDA devices"
            self.close()
            raise RTB200Error(rc, msg)
        self.n = len(devs)

    def close(self):
        if self.h:
            self.L.rtb200_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def create_image(self, problem, flags=0, image=None, I_ang=None, raise_on_failed=True):
        e = problem.euv_beam
        cp, keep = problem.c_struct()
        image = np.empty(e.nx * e.ny * e.nv) if image is None else image
        I_ang = np.empty(e.na * e.nb) if I_ang is None else I_ang
        _check_out(image, e.nx * e.ny * e.nv, "image")
        _check_out(I_ang, e.na * e.nb, "I_ang")
        fc, nf = C.c_uint(0), C.c_int(0)
        failed = np.zeros(abi.N_FAILED_MAX, abi.ray_dtype)
        rc = self.L.rtb200_multi_create_image(
            self.h, C.byref(cp), flags, _addr(image), _addr(I_ang), C.byref(fc),
            failed.ctypes.data_as(C.POINTER(abi.Ray)), abi.N_FAILED_MAX, C.byref(nf))
        if rc < 0:
            raise RTB200Error(rc, self.L.rtb200_multi_last_error(self.h).decode())
        self.failure_code, self.n_failed = fc.value, nf.value
        self.failed = failed[:min(nf.value, abi.N_FAILED_MAX)]
        if rc == abi.RAYS_FAILED and raise_on_failed:
            raise RaysFailed(fc.value, self.failed)
        return image, I_ang

    def timings(self):
        t = (abi.Timings * self.n)()
        ex, tot = C.c_float(0), C.c_float(0)
        self.L.rtb200_multi_get_timings(self.h, t, C.byref(ex), C.byref(tot))
        per = [{k: getattr(x, k) for k, _ in abi.Timings._fields_ if k != "reserved"} for x in t]
        return {"per_device": per, "exchange_ms": ex.value, "total_ms": tot.value}
