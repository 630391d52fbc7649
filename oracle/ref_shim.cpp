// ref_shim.cpp — TEST INFRASTRUCTURE ONLY.  A thin extern "C" door into the UNMODIFIED reference
// (compiled from /root/reference/src where it lies, see oracle/Makefile) so that Python tests
// and bench.py's CPU-baseline legs can run the reference's own RayTrace::create_image("cpu" /
// "threads") and RayTrace_calc_ray on a .dat file.  Nothing here is used by the product library.
//
// Built only where /root/reference exists; the resulting oracle/_ref/libref_oracle.so travels to
// the GPU box as a prebuilt file.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "RayTrace.h"
#include "common/RayTraceImageHelper.h"

namespace {
struct Handle {
    RayTrace::create_image_struct *info = nullptr;
    double *golden_image = nullptr;
    double *golden_I_ang = nullptr;
};
} // namespace

extern "C" {

// Mirrors loadInput (src/CreateImage.cpp:26-58): uint64 length, payload, unpack; the embedded
// golden arrays are detached from `info`.
void *ref_load(const char *filename)
{
    FILE *fid = fopen(filename, "rb");
    if (!fid)
        return nullptr;
    uint64_t n = 0;
    if (fread(&n, sizeof(n), 1, fid) != 1) {
        fclose(fid);
        return nullptr;
    }
    std::vector<char> data(n);
    if (fread(data.data(), 1, n, fid) != n) {
        fclose(fid);
        return nullptr;
    }
    fclose(fid);
    auto *h = new Handle;
    h->info = new RayTrace::create_image_struct();
    h->info->unpack(std::pair<const char *, size_t>(data.data(), n));
    h->golden_image = h->info->image;
    h->golden_I_ang = h->info->I_ang;
    h->info->image = nullptr;
    h->info->I_ang = nullptr;
    return h;
}

void ref_free(void *hv)
{
    auto *h = static_cast<Handle *>(hv);
    if (!h)
        return;
    free(h->golden_image);
    free(h->golden_I_ang);
    free(h->info->image);
    free(h->info->I_ang);
    h->info->image = h->info->I_ang = nullptr;
    delete h->info->euv_beam;
    delete h->info->seed_beam;
    delete[] h->info->gain;
    delete h->info->seed;
    delete h->info;
    delete h;
}

// dims: N, N_start, N_parallel, nx, ny, na, nb, nv, has_seed, has_golden_image, has_golden_I_ang
void ref_info(void *hv, int *dims)
{
    auto *h = static_cast<Handle *>(hv);
    const auto *e = h->info->euv_beam;
    int v[11] = { h->info->N, h->info->N_start, h->info->N_parallel, e->nx, e->ny, e->na, e->nb,
        e->nv, h->info->seed != nullptr, h->golden_image != nullptr, h->golden_I_ang != nullptr };
    memcpy(dims, v, sizeof(v));
}

void ref_golden(void *hv, double *image, double *I_ang)
{
    auto *h = static_cast<Handle *>(hv);
    const auto *e = h->info->euv_beam;
    if (image && h->golden_image)
        memcpy(image, h->golden_image, sizeof(double) * e->nx * e->ny * e->nv);
    if (I_ang && h->golden_I_ang)
        memcpy(I_ang, h->golden_I_ang, sizeof(double) * e->na * e->nb);
}

// Runs RayTrace::create_image(info, method) exactly as run_tests does (src/CreateImage.cpp:147-152)
// and returns the wall time of that call in seconds.  Aborts the process on failed rays, like
// the reference.
double ref_create_image(void *hv, const char *method, double *image, double *I_ang)
{
    auto *h = static_cast<Handle *>(hv);
    const auto *e = h->info->euv_beam;
    auto t0 = std::chrono::steady_clock::now();
    RayTrace::create_image(h->info, method);
    auto t1 = std::chrono::steady_clock::now();
    if (image)
        memcpy(image, h->info->image, sizeof(double) * e->nx * e->ny * e->nv);
    if (I_ang)
        memcpy(I_ang, h->info->I_ang, sizeof(double) * e->na * e->nb);
    free(h->info->image);
    free(h->info->I_ang);
    h->info->image = h->info->I_ang = nullptr;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Per-ray door: RayTrace_calc_ray (src/common/RayTraceImageHelper.h:379-595) on the loaded
// problem.  rays/ray2 are n x 4 floats, Iv n x K doubles, error n ints; debug (optional) is
// n x 3*(N_SUB*(N-1)+1) floats in the reference's RAY_DEBUG layout (x, y, I per sub-segment).
void ref_calc_rays(void *hv, int method, const float *rays, int n, double *Iv, float *ray2,
    int *error, float *debug)
{
    auto *h = static_cast<Handle *>(hv);
    const auto *info = h->info;
    const auto *e = info->euv_beam;
    const int K = e->nv;
    const int nd = 3 * (N_SUB * (info->N - 1) + 1);
    for (int i = 0; i < n; i++) {
        ray_struct r, r2;
        r.x = rays[4 * i + 0];
        r.y = rays[4 * i + 1];
        r.a = rays[4 * i + 2];
        r.b = rays[4 * i + 3];
        r2.x = r2.y = r2.a = r2.b = 0;
        double Iv_tmp[K_MAX];
        error[i] = RayTrace_calc_ray(r, info->N, (float) e->dz, info->gain, info->seed, K, method,
            Iv_tmp, r2, 0.5f, debug ? e->dv : nullptr, debug ? &debug[(size_t) i * nd] : nullptr);
        memcpy(&Iv[(size_t) i * K], Iv_tmp, sizeof(double) * K);
        ray2[4 * i + 0] = r2.x;
        ray2[4 * i + 1] = r2.y;
        ray2[4 * i + 2] = r2.a;
        ray2[4 * i + 3] = r2.b;
    }
}

// Unit doors for the helper functions (src/common/RayTraceImageHelper.h:101-220).
unsigned ref_findindex(const double *X, unsigned n, double Y) { return findindex(X, n, Y); }
size_t ref_findfirstsingle(const double *X, size_t n, double Y) { return findfirstsingle(X, n, Y); }
float ref_bilinear(float dx, float dy, float f1, float f2, float f3, float f4)
{
    return bilinear(dx, dy, f1, f2, f3, f4);
}
double ref_interp_pchip(size_t N, const double *xi, const double *yi, double x)
{
    return interp_pchip(N, xi, yi, x);
}
void ref_calc_seed(void *hv, double x, double y, double a, double b, double *Iv)
{
    auto *h = static_cast<Handle *>(hv);
    RayTrace::calc_seed(*h->info->seed, x, y, a, b, Iv);
}
int ref_hardware_threads() { return (int) std::thread::hardware_concurrency(); }

} // extern "C"
