// RayTraceImageB200.cpp — the reference-side back-end loop of the B200 path.
//
// This is the file a maintainer of the reference adds next to src/RayTraceImageCPU.cpp and
// src/RayTraceImageCuda.cu: it has the exact `RayTraceImage<Backend>Loop` signature that
// RayTrace::create_image dispatches to (extern declarations at src/RayTraceImage.cpp:47-75,
// call sites :350-423) and forwards to the C ABI of include/rtb200.h (librtb200.so).  It is
// compiled against the reference's own headers, so it is built only where the reference
// checkout exists (oracle/Makefile, target `b200`); nothing in librtb200.so depends on it.
//
// Behaviour mirrored from the reference's loops:
//  * image / I_ang are accumulated into (src/RayTraceImageCPU.cpp:56-68);
//  * failed rays are appended to failed_rays and their codes OR-ed into failure_code
//    (:32-36); the caller then aborts with "Some rays failed";
//  * device / runtime errors print a message and exit(-1), like CUDA_CHECK
//    (src/RayTraceImageCuda.cu:8-18);
//  * re-entrant: one rtb200 context per host thread (the legacy CUDA loop keeps
//    non-thread-safe static state, src/RayTraceImageCuda.cu:153-160).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <vector>

#include "RayTrace.h"
#include "common/RayTraceImageHelper.h"

#include "rtb200.h"

namespace {

struct ThreadContext {
    rtb200_ctx *ctx = nullptr;
    int device = -1;
    ~ThreadContext() { rtb200_destroy(ctx); }
};
thread_local ThreadContext tls;
std::mutex pending_mutex;
std::deque<int> pending_devices; // devices announced by the parent for workers not yet started

rtb200_ctx *context()
{
    int device = -1;
    if (!tls.ctx) { // a fresh worker thread takes the next announced device, if any
        std::lock_guard<std::mutex> lock(pending_mutex);
        if (!pending_devices.empty()) {
            device = pending_devices.front();
            pending_devices.pop_front();
        }
    } else {
        device = tls.device;
    }
    if (device < 0) {
        const char *e = getenv("RTB200_DEVICE");
        device = e ? atoi(e) : 0;
    }
    if (tls.ctx && tls.device == device)
        return tls.ctx;
    rtb200_destroy(tls.ctx);
    tls.ctx = nullptr;
    if (rtb200_create(device, &tls.ctx) != RTB200_OK) {
        fprintf(stderr, "rtb200: no usable CUDA device %d (the b200 method has no CPU fallback)\n", device);
        exit(-1);
    }
    tls.device = device;
    return tls.ctx;
}

// True when `rays` is the complete enumeration of the euv grid in create_image's order
// (x slowest ... b fastest, coordinates rounded to float).  The size must match exactly; the
// coordinates are checked on the first and last 64 rays and on 4096 evenly spaced ones.
bool is_full_grid(const RayTrace::EUV_beam_struct &beam, const std::vector<ray_struct> &rays)
{
    const size_t Nt = (size_t) beam.nx * beam.ny * beam.na * beam.nb;
    if (rays.size() != Nt || Nt == 0)
        return false;
    auto matches = [&](size_t ijkm) {
        const int m = (int) (ijkm % beam.nb);
        const int k = (int) ((ijkm / beam.nb) % beam.na);
        const int j = (int) ((ijkm / ((size_t) beam.na * beam.nb)) % beam.ny);
        const int i = (int) (ijkm / ((size_t) beam.ny * beam.na * beam.nb));
        const ray_struct &r = rays[ijkm];
        return r.x == (float) beam.x[i] && r.y == (float) beam.y[j] && r.a == (float) beam.a[k] &&
               r.b == (float) beam.b[m];
    };
    for (size_t q = 0; q < 64 && q < Nt; q++)
        if (!matches(q) || !matches(Nt - 1 - q))
            return false;
    const size_t step = Nt / 4096 + 1;
    for (size_t q = 0; q < Nt; q += step)
        if (!matches(q))
            return false;
    return true;
}

} // namespace

// setID hook for RayTraceImageThreadLoop ("b200-multigpu").  The reference calls setID(i) in the
// PARENT thread just before it starts worker i (src/RayTraceImage.cpp:116-119), so its setGPU
// (cudaSetDevice, :82-88) never reaches the worker and every worker lands on device 0.  Here the
// parent only announces the device; each freshly started worker thread picks one announcement up
// when it creates its context, so the N workers run on N distinct devices.
void RayTraceImageB200SetDevice(int device)
{
    std::lock_guard<std::mutex> lock(pending_mutex);
    pending_devices.push_back(device);
}

void RayTraceImageB200Loop(int N, const RayTrace::EUV_beam_struct &beam,
    const RayTrace::ray_gain_struct *gain, const RayTrace::ray_seed_struct *seed, int method,
    const std::vector<ray_struct> &rays, double scale, double *image, double *I_ang,
    unsigned int &failure_code, std::vector<ray_struct> &failed_rays)
{
    static_assert(sizeof(ray_struct) == sizeof(rtb200_ray), "ray_struct layout");
    failure_code = 0;
    rtb200_beam b;
    memset(&b, 0, sizeof(b));
    b.nx = beam.nx;
    b.ny = beam.ny;
    b.na = beam.na;
    b.nb = beam.nb;
    b.nv = beam.nv;
    b.dx = beam.dx;
    b.dy = beam.dy;
    b.da = beam.da;
    b.db = beam.db;
    b.dz = beam.dz;
    b.x = beam.x;
    b.y = beam.y;
    b.a = beam.a;
    b.b = beam.b;
    b.dv = beam.dv;
    std::vector<rtb200_gain_plane> planes(N);
    for (int i = 0; i < N; i++) {
        planes[i].Nx = gain[i].Nx;
        planes[i].Ny = gain[i].Ny;
        planes[i].Nv = gain[i].Nv;
        planes[i].x = gain[i].x;
        planes[i].y = gain[i].y;
        planes[i].n = gain[i].n;
        planes[i].g0 = gain[i].g0;
        planes[i].E0 = gain[i].E0;
        planes[i].gv = gain[i].gv;
    }
    rtb200_seed sd;
    if (seed) {
        for (int i = 0; i < 5; i++) {
            sd.dim[i] = seed->dim[i];
            sd.x[i] = seed->x[i];
            sd.f[i] = seed->f[i];
        }
        sd.f0 = seed->f0;
    }
    rtb200_ctx *ctx = context();
    rtb200_ray failed[RTB200_N_FAILED_MAX];
    int n_failed = 0;
    int rc;
    if (method == 1 && !seed && scale == 1.0 && is_full_grid(beam, rays)) {
        // create_image passed the complete ASE ray enumeration of the euv grid
        // (src/RayTraceImage.cpp:300-328 with N_start = 0, N_parallel = 1): let the device
        // enumerate the rays itself.  No ray upload, pixels owned by CTAs, no atomics on the image.
        rtb200_problem prob;
        memset(&prob, 0, sizeof(prob));
        prob.N = N;
        prob.N_start = 0;
        prob.N_parallel = 1;
        prob.euv_beam = &b;
        prob.gain = planes.data();
        const size_t n_img = (size_t) beam.nx * beam.ny * beam.nv, n_ang = (size_t) beam.na * beam.nb;
        std::vector<double> tmp(n_img + n_ang);
        rc = rtb200_create_image(ctx, &prob, RTB200_FLAG_NO_LIMITS, tmp.data(), tmp.data() + n_img,
            &failure_code, failed, RTB200_N_FAILED_MAX, &n_failed);
        if (rc >= 0) { // accumulate, like every *Loop (src/RayTraceImageCPU.cpp:56-68)
            for (size_t i = 0; i < n_img; i++)
                image[i] += tmp[i];
            for (size_t i = 0; i < n_ang; i++)
                I_ang[i] += tmp[n_img + i];
        }
    } else {
        rc = rtb200_trace_rays(ctx, N, &b, planes.data(), seed ? &sd : nullptr, method,
            reinterpret_cast<const rtb200_ray *>(rays.data()), rays.size(), scale, image, I_ang,
            &failure_code, failed, RTB200_N_FAILED_MAX, &n_failed);
    }
    if (rc < 0) {
        fprintf(stderr, "rtb200 error %d: %s\n", rc, rtb200_last_error(ctx));
        exit(-1);
    }
    for (int i = 0; i < n_failed && i < RTB200_N_FAILED_MAX; i++) {
        ray_struct r;
        r.x = failed[i].x;
        r.y = failed[i].y;
        r.a = failed[i].a;
        r.b = failed[i].b;
        failed_rays.push_back(r);
    }
}


// ---- the full-speed entries: called by create_image BEFORE it builds the host ray list --------
// (src/RayTraceImage.cpp:277), method names "b200-direct" and "b200-multigpu".  The problem
// descriptor carries the grids and N_start / N_parallel, so the rays are enumerated on the
// device, ASE pixels are owned by thread blocks and the seed is tabulated per grid index
// (rtb200_create_image).  image / I_ang are the zeroed buffers create_image allocated
// (:271-274); failures are reported the way create_image does (:427-430): messages on stderr,
// then exit(-1).
namespace {

struct ProblemView { // POD descriptors of include/rtb200.h filled from the reference's structs (pointers only)
    rtb200_beam euv, sbeam;
    std::vector<rtb200_gain_plane> planes;
    rtb200_seed sd;
    rtb200_problem prob;
    explicit ProblemView(const RayTrace::create_image_struct *info)
    {
        auto fill = [](rtb200_beam &b, int nx, int ny, int na, int nb, double dx, double dy, double da,
                       double db, const double *x, const double *y, const double *a, const double *bb) {
            memset(&b, 0, sizeof(b));
            b.nx = nx, b.ny = ny, b.na = na, b.nb = nb;
            b.dx = dx, b.dy = dy, b.da = da, b.db = db;
            b.x = x, b.y = y, b.a = a, b.b = bb;
        };
        const RayTrace::EUV_beam_struct &e = *info->euv_beam;
        fill(euv, e.nx, e.ny, e.na, e.nb, e.dx, e.dy, e.da, e.db, e.x, e.y, e.a, e.b);
        euv.nv = e.nv;
        euv.dz = e.dz;
        euv.dv = e.dv;
        if (info->seed_beam) {
            const RayTrace::seed_beam_struct &s = *info->seed_beam;
            fill(sbeam, s.nx, s.ny, s.na, s.nb, s.dx, s.dy, s.da, s.db, s.x, s.y, s.a, s.b);
        }
        planes.resize(info->N);
        for (int i = 0; i < info->N; i++) {
            const RayTrace::ray_gain_struct &g = info->gain[i];
            planes[i].Nx = g.Nx, planes[i].Ny = g.Ny, planes[i].Nv = g.Nv;
            planes[i].x = g.x, planes[i].y = g.y, planes[i].n = g.n;
            planes[i].g0 = g.g0, planes[i].E0 = g.E0, planes[i].gv = g.gv;
        }
        if (info->seed) {
            for (int i = 0; i < 5; i++) {
                sd.dim[i] = info->seed->dim[i];
                sd.x[i] = info->seed->x[i];
                sd.f[i] = info->seed->f[i];
            }
            sd.f0 = info->seed->f0;
        }
        memset(&prob, 0, sizeof(prob));
        prob.N = info->N;
        prob.N_start = info->N_start;
        prob.N_parallel = info->N_parallel;
        prob.euv_beam = &euv;
        prob.seed_beam = info->seed_beam ? &sbeam : nullptr;
        prob.gain = planes.data();
        prob.seed = info->seed ? &sd : nullptr;
    }
};

void report(int rc, const char *msg, unsigned failure_code)
{
    if (rc < 0) {
        fprintf(stderr, "rtb200 error %d: %s\n", rc, msg);
        exit(-1);
    }
    if (failure_code != 0) {
        if (failure_code & 2u)
            fprintf(stderr, "Invalid ray detected\n");
        if (failure_code & 4u)
            fprintf(stderr, "Negitive intensity detected\n");
        if (failure_code & 8u)
            fprintf(stderr, "NaNs detected in intensity\n");
        fprintf(stderr, "Some rays failed\n");
        exit(-1);
    }
}

} // namespace

void RayTraceImageB200Direct(const RayTrace::create_image_struct *info, double *image, double *I_ang)
{
    ProblemView v(info);
    rtb200_ctx *ctx = context();
    unsigned failure_code = 0;
    int n_failed = 0;
    const int rc = rtb200_create_image(ctx, &v.prob, 0, image, I_ang, &failure_code, nullptr, 0, &n_failed);
    report(rc, rtb200_last_error(ctx), failure_code);
}

// "b200-multigpu": every device of the box (or RTB200_NGPU of them) in one call.  Replaces the
// reference's cuda-multigpu branch (src/RayTraceImage.cpp:396-405: ThreadLoop over contiguous
// chunks of a host ray list, host-side sum, and - through the cudaSetDevice in the parent thread,
// :116-119 - every worker on device 0) by rtb200_multi_create_image: rows sharded row-cyclically,
// partial results exchanged over NVLink by NCCL, one download.  The communicator is created at the
// first call and kept (ncclCommInitAll takes about a second).
void RayTraceImageB200MultiDirect(const RayTrace::create_image_struct *info, double *image, double *I_ang)
{
    static std::mutex mtx;
    static rtb200_multi *multi = nullptr;
    std::lock_guard<std::mutex> lock(mtx);
    if (!multi) {
        int n = rtb200_device_count();
        if (const char *e = getenv("RTB200_NGPU"))
            n = atoi(e) < n ? atoi(e) : n;
        if (n < 1 || rtb200_multi_create(nullptr, n, &multi) != RTB200_OK) {
            fprintf(stderr, "rtb200: cannot set up %d device(s): %s\n", n, rtb200_multi_last_error(multi));
            exit(-1);
        }
    }
    ProblemView v(info);
    unsigned failure_code = 0;
    int n_failed = 0;
    const int rc = rtb200_multi_create_image(multi, &v.prob, 0, image, I_ang, &failure_code, nullptr, 0, &n_failed);
    report(rc, rtb200_multi_last_error(multi), failure_code);
}
