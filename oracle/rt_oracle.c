/*
 * rt_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement, in plain C, of the reference's
 * image-formation hot path.  It is the parity checker for the CUDA path; it is never linked
 * into, called by or shipped with the product library (librtb200.so).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file bit-for-bit
 * against the reference's own RayTraceImageCPULoop compiled from /root/reference (oracle/_ref,
 * see oracle/Makefile) on ASE_small.dat and seed_small.dat, and tests/golden/ holds those
 * reference outputs so the check also runs where /root/reference does not exist.
 *
 * Every function cites the reference lines it follows (paths relative to the reference
 * checkout).  The arithmetic is spelled out operation by operation because the reference is
 * mixed float/double C++ and a 1e-10 image match needs the float march to be bit-exact:
 * compile with  gcc -O2 -ffp-contract=off  (x86-64 SSE2, FLT_EVAL_METHOD == 0).
 */
#include "rt_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define N_SUB RTB200_N_SUB

typedef struct {
    float x, y, z;
} vec3f;

/* src/common/RayTraceImageHelper.h:73-89 (normalize_s, the #else branch).
 * `1.0 / sqrt(tmp)`: float sqrt (C++ overload), double divide, rounded to float on store. */
static void normalize_s(vec3f *s)
{
    float tmp = s->x * s->x + s->y * s->y + s->z * s->z;
    tmp = (float) (1.0 / (double) sqrtf(tmp));
    s->x *= tmp;
    s->y *= tmp;
    s->z *= tmp;
}

/* src/common/RayTraceImageHelper.h:101-117 */
size_t rt_oracle_findfirstsingle(const double *X, size_t size_X, double Y)
{
    if (Y < X[0])
        return 0;
    if (Y > X[size_X - 1])
        return size_X;
    size_t lower = 0, upper = size_X - 1;
    while ((upper - lower) != 1) {
        size_t value = (upper + lower) / 2;
        if (X[value] >= Y)
            upper = value;
        else
            lower = value;
    }
    return upper;
}

/* src/common/RayTraceImageHelper.h:131-143 */
uint32_t rt_oracle_findindex(const double *X, uint32_t size_X, double Y)
{
    uint32_t lower = 0, upper = size_X - 1;
    while ((upper - lower) != 1) {
        uint32_t value = (upper + lower) / 2;
        if (X[value] >= Y)
            upper = value;
        else
            lower = value;
    }
    return upper;
}

/* src/common/RayTraceImageHelper.h:153-158 */
float rt_oracle_bilinear(float dx, float dy, float f1, float f2, float f3, float f4)
{
    float dx2 = 1.0f - dx;
    float dy2 = 1.0f - dy;
    return (dx * f2 + dx2 * f1) * dy2 + (dx * f4 + dx2 * f3) * dy;
}

/* src/common/RayTraceImageHelper.h:168-220 */
double rt_oracle_interp_pchip(size_t N, const double *xi, const double *yi, double x)
{
    if (x <= xi[0] || N <= 2) {
        double dx = (x - xi[0]) / (xi[1] - xi[0]);
        return (1.0 - dx) * yi[0] + dx * yi[1];
    } else if (x >= xi[N - 1]) {
        double dx = (x - xi[N - 2]) / (xi[N - 1] - xi[N - 2]);
        return (1.0 - dx) * yi[N - 2] + dx * yi[N - 1];
    }
    size_t i = rt_oracle_findfirstsingle(xi, N, x);
    double f1 = yi[i - 1];
    double f2 = yi[i];
    double dx = (x - xi[i - 1]) / (xi[i] - xi[i - 1]);
    double g1 = 0, g2 = 0;
    if (i <= 1) {
        g1 = f2 - f1;
    } else if ((f1 < f2 && f1 > yi[i - 2]) || (f1 > f2 && f1 < yi[i - 2])) {
        double f0 = yi[i - 2];
        double dx1 = xi[i - 1] - xi[i - 2];
        double dx2 = xi[i] - xi[i - 1];
        double a1 = (dx2 - dx1) / dx1;
        double a2 = dx1 / (dx1 + dx2);
        g1 = a1 * (f1 - f0) + a2 * (f2 - f0);
        double fx1 = fabs(f1 - f0) / dx1;
        double fx2 = fabs(f2 - f1) / dx2;
        double g_max = 2 * dx2 * (fx1 < fx2 ? fx1 : fx2);
        g1 = ((g1 >= 0) ? 1 : -1) * (fabs(g1) < g_max ? fabs(g1) : g_max);
    }
    if (i >= N - 1) {
        g2 = f2 - f1;
    } else if ((f2 < f1 && f2 > yi[i + 1]) || (f2 > f1 && f2 < yi[i + 1])) {
        double f0 = yi[i + 1];
        double dx1 = xi[i] - xi[i - 1];
        double dx2 = xi[i + 1] - xi[i];
        double a1 = -dx2 / (dx1 + dx2);
        double a2 = (dx2 - dx1) / dx2;
        g2 = a1 * (f1 - f0) + a2 * (f2 - f0);
        double fx1 = fabs(f2 - f1) / dx1;
        double fx2 = fabs(f0 - f2) / dx2;
        double g_max = 2 * dx1 * (fx1 < fx2 ? fx1 : fx2);
        g2 = ((g2 >= 0) ? 1 : -1) * (fabs(g2) < g_max ? fabs(g2) : g_max);
    }
    double dx2 = dx * dx;
    return f1 + dx2 * (2 * dx - 3) * (f1 - f2) + dx * g1 - dx2 * (g1 + (1 - dx) * (g1 + g2));
}

/* src/common/RayTraceImageHelper.h:230-247 */
void rt_oracle_calc_seed(const rtb200_seed *seed, double x, double y, double a, double b,
                         double *Iv)
{
    double f = 0.0;
    if (x >= seed->x[0][0] && x <= seed->x[0][seed->dim[0] - 1] && y >= seed->x[1][0] &&
        y <= seed->x[1][seed->dim[1] - 1] && a >= seed->x[2][0] &&
        a <= seed->x[2][seed->dim[2] - 1] && b >= seed->x[3][0] &&
        b <= seed->x[3][seed->dim[3] - 1]) {
        double fx = rt_oracle_interp_pchip(seed->dim[0], seed->x[0], seed->f[0], x);
        double fy = rt_oracle_interp_pchip(seed->dim[1], seed->x[1], seed->f[1], y);
        double fa = rt_oracle_interp_pchip(seed->dim[2], seed->x[2], seed->f[2], a);
        double fb = rt_oracle_interp_pchip(seed->dim[3], seed->x[3], seed->f[3], b);
        f = seed->f0 * fx * fy * fa * fb;
        f = f < 0.0 ? 0.0 : f;
    }
    for (int i = 0; i < seed->dim[4]; i++)
        Iv[i] = f * seed->f[4][i];
}

/* src/common/RayTraceImageHelper.h:270-313 (propagate).  All arithmetic float except the
 * `fabs(n - n0) < 0.05` comparison, which the reference evaluates against a double literal. */
static float propagate(vec3f *r, vec3f *s, float n0, float dn_dx, float dn_dy,
                       const float dx[3], float c, uint64_t *steps)
{
    float sum = 0.0f;
    float dz_max = c * 1.00001f * dx[2];
    r->x = 0;
    r->y = 0;
    r->z = 0;
    float n = n0;
    while (fabsf(r->x) < dx[0] && fabsf(r->y) < dx[1] && fabsf(r->z) < dx[2] &&
           (double) fabsf(n - n0) < 0.05) {
        n = n0 + r->x * dn_dx + r->y * dn_dy;
        float t = (s->x * dn_dx + s->y * dn_dy + 1e-12f) / n;
        float f[3] = { dn_dx / n - s->x * t, dn_dy / n - s->y * t, -s->z * t };
        float step = c * 0.1f / fabsf(t);
        step = step < dz_max ? step : dz_max;
        float step2 = 1.0001f * (dx[2] - fabsf(r->z)) / fabsf(s->z);
        float step3 = c * 0.05f * (fabsf(s->x) + 5e-4f) / (fabsf(f[0]) + 1e-8f);
        float step4 = c * 0.05f * (fabsf(s->y) + 5e-4f) / (fabsf(f[1]) + 1e-8f);
        step = step < step2 ? step : step2;
        step = step < step3 ? step : step3;
        step = step < step4 ? step : step4;
        float st = step * t;
        float c1 = 0.5f * step * step * (1.0f - st / 3.0f + st * st / 12.0f);
        r->x += s->x * step + c1 * f[0];
        r->y += s->y * step + c1 * f[1];
        r->z += s->z * step + c1 * f[2];
        float c2 = step * (1.0f - 0.5f * st + st * st / 6.0f);
        s->x += c2 * f[0];
        s->y += c2 * f[1];
        s->z += c2 * f[2];
        normalize_s(s);
        sum += step;
        if (steps)
            ++*steps;
    }
    return sum;
}

/* src/common/RayTraceImageHelper.h:318-351 (propagate2) */
static float propagate2(vec3f *pos, vec3f *s, float dz, const double x[2], const double y[2],
                        const float range[4], const double n[4], int abs_y, float c,
                        uint64_t *steps)
{
    float z = 0.0f;
    float ds_sum = 0.0f;
    const float dx = (float) (x[1] - x[0]);
    const float dy = (float) (y[1] - y[0]);
    float y2 = abs_y ? fabsf(pos->y) : pos->y;
    while (pos->x > range[0] && pos->x < range[1] && y2 > range[2] && y2 < range[3] &&
           (double) z < 0.999 * (double) dz) {
        y2 = abs_y ? fabsf(pos->y) : pos->y;
        float dxi = (float) (((double) pos->x - x[0]) / (double) dx);
        float dyi = (float) (((double) y2 - y[0]) / (double) dy);
        float n0 = rt_oracle_bilinear(dxi, dyi, (float) n[0], (float) n[1], (float) n[2],
                                      (float) n[3]);
        float dn_dx = (float) ((1.0 - (double) dyi) * (n[1] - n[0]) / (double) dx +
                               (double) dyi * (n[3] - n[2]) / (double) dx);
        float dn_dy = (float) ((1.0 - (double) dxi) * (n[2] - n[0]) / (double) dy +
                               (double) dxi * (n[3] - n[1]) / (double) dy);
        if (abs_y && pos->y < 0)
            dn_dy = -dn_dy;
        vec3f r = { 0, 0, 0 };
        float dx_max[3] = { 0.1f * dx, 0.1f * dy, dz - z };
        ds_sum += propagate(&r, s, n0, dn_dx, dn_dy, dx_max, c, steps);
        pos->x += r.x;
        pos->y += r.y;
        pos->z += r.z;
        z += fabsf(r.z);
        y2 = abs_y ? fabsf(pos->y) : pos->y;
    }
    return ds_sum;
}

/* src/common/RayTraceImageHelper.h:379-595 (RayTrace_calc_ray, non-debug path).
 * gvl/evl/ivl are [(N-1)*N_SUB] in the reference's [i][is] order (caller scratch, may be
 * inspected afterwards).  Returns 0, -1, -2 or -3 like the reference. */
static int calc_ray_impl(const rtb200_ray *ray, int N, float dz0, const rtb200_gain_plane *gain,
                         const rtb200_seed *seed, int K, int method, float c, double *Iv,
                         rtb200_ray *ray2, float *gvl, float *evl, int32_t *ivl,
                         int *escaped_out, uint64_t *steps, const double *dv, float *debug)
{
    const int S = (N - 1) * N_SUB;
    for (int i = 0; i < S; i++) {
        gvl[i] = 0.0f;
        evl[i] = 0.0f;
        ivl[i] = 0;
    }
    for (int k = 0; k < K; ++k)
        Iv[k] = 0.0;

    int use_emis = gain->E0 != NULL && seed == NULL; /* :402 */

    vec3f s, pos;
    pos.x = ray->x;
    pos.y = ray->y;
    pos.z = 0.0f;
    s.x = tanf(1e-3f * ray->a); /* :409-411, C++ tan(float) == tanf */
    s.y = tanf(1e-3f * ray->b);
    s.z = 1.0f;
    if (method == 1) {
        s.x = -s.x;
        s.y = -s.y;
        s.z = -s.z;
    }
    normalize_s(&s);
    if (dv != NULL && debug != NULL) { /* RAY_DEBUG, :419-426 */
        int ii = method == 1 ? (N - 1) * N_SUB : 0;
        memset(debug, 0, 3 * (size_t) (N_SUB * (N - 1) + 1) * sizeof(float));
        debug[3 * ii + 0] = pos.x;
        debug[3 * ii + 1] = pos.y;
    }

    int escaped = 0;
    for (int i = 0; i < N - 1 && !escaped; i++) { /* :430 */
        int ii = method == 1 ? N - i - 1 : i + 1;
        const rtb200_gain_plane *g = &gain[ii];
        uint32_t Nx = (uint32_t) g->Nx, Ny = (uint32_t) g->Ny;
        float range[4];
        range[0] = (float) g->x[0];
        range[1] = (float) g->x[Nx - 1];
        range[2] = (float) g->y[0];
        range[3] = (float) g->y[Ny - 1];
        int abs_y = 0;
        if (range[2] >= 0) {
            range[2] = -range[3];
            abs_y = 1;
        }
        const double *ptr_x = g->x, *ptr_y = g->y, *ptr_n = g->n;
        const float *ptr_g0 = g->g0, *ptr_E0 = g->E0;
        float z = 0.0f;
        for (int iz = 0; iz < N_SUB; iz++) { /* :460 */
            int is = method == 1 ? N_SUB - iz - 1 : iz;
            float z_stop = (dz0 * ((float) iz + 1.0f) / (float) N_SUB);
            while (z < 0.995f * z_stop) {
                if (pos.x < range[0] || pos.x > range[1] || pos.y < range[2] ||
                    pos.y > range[3] || (double) (s.z * s.z) < 0.01) {
                    escaped = 1;
                    break;
                }
                float y2 = abs_y ? fabsf(pos.y) : pos.y;
                uint32_t k1 = rt_oracle_findindex(ptr_x, Nx, (double) pos.x);
                uint32_t k2 = rt_oracle_findindex(ptr_y, Ny, (double) y2);
                uint32_t i1 = (k1 - 1) + (k2 - 1) * Nx;
                uint32_t i2 = k1 + (k2 - 1) * Nx;
                uint32_t i3 = (k1 - 1) + k2 * Nx;
                uint32_t i4 = k1 + k2 * Nx;
                double x[2] = { ptr_x[k1 - 1], ptr_x[k1] };
                double y[2] = { ptr_y[k2 - 1], ptr_y[k2] };
                double n[4] = { ptr_n[i1], ptr_n[i2], ptr_n[i3], ptr_n[i4] };
                float dxi = (float) (((double) pos.x - ptr_x[k1 - 1]) /
                                     (ptr_x[k1] - ptr_x[k1 - 1]));
                float dyi = (float) (((double) y2 - ptr_y[k2 - 1]) /
                                     (ptr_y[k2] - ptr_y[k2 - 1]));
                float g0 = rt_oracle_bilinear(dxi, dyi, ptr_g0[i1], ptr_g0[i2], ptr_g0[i3],
                                              ptr_g0[i4]);
                float E0 = 0.0f;
                if (use_emis) {
                    E0 = rt_oracle_bilinear(dxi, dyi, ptr_E0[i1], ptr_E0[i2], ptr_E0[i3],
                                            ptr_E0[i4]);
                    E0 = E0 >= 0 ? E0 : 0.0f;
                }
                pos.z = 0.0f;
                float cell[4] = { (float) (x[0] - 0.1 * (ptr_x[k1] - ptr_x[k1 - 1])),
                                  (float) (x[1] + 0.1 * (ptr_x[k1] - ptr_x[k1 - 1])),
                                  (float) (y[0] - 0.1 * (ptr_y[k2] - ptr_y[k2 - 1])),
                                  (float) (y[1] + 0.1 * (ptr_y[k2] - ptr_y[k2 - 1])) };
                if (abs_y && k2 <= 1)
                    cell[2] = -cell[3];
                float ds_sum = propagate2(&pos, &s, z_stop - z, x, y, cell, n, abs_y, c, steps);
                z += fabsf(pos.z);
                int idx = (ii - 1) * N_SUB + is;
                gvl[idx] += g0 * ds_sum;
                evl[idx] += E0 * ds_sum;
                ivl[idx] = (int32_t) i1;
            }
            if (dv != NULL && debug != NULL) { /* :505-511 */
                int index = N_SUB * (ii - 1) + is + (method == 1 ? 0 : 1);
                debug[3 * index + 0] = pos.x;
                debug[3 * index + 1] = pos.y;
            }
        }
    }
    if (escaped_out)
        *escaped_out = escaped;
    if ((double) (s.z * s.z) < 0.01) /* :515 */
        return -1;
    ray2->x = pos.x;
    ray2->y = pos.y;
    ray2->a = atanf(s.x / s.z) * 1e3f; /* :520-521, C++ atan(float) == atanf */
    ray2->b = atanf(s.y / s.z) * 1e3f;
    if (seed == NULL || escaped) {
        /* nothing */
    } else if (method == 1) {
        rt_oracle_calc_seed(seed, pos.x, pos.y, (double) ray2->a, (double) ray2->b, Iv);
    } else if (method == 2) {
        rt_oracle_calc_seed(seed, ray->x, ray->y, ray->a, ray->b, Iv);
    }
    if (dv != NULL && debug != NULL) { /* :536-542 */
        debug[2] = 0.0f;
        for (int k = 0; k < K; k++)
            debug[2] += (float) (2 * Iv[k] * dv[k]);
    }
    if (use_emis || debug != NULL) { /* :543-568 */
        for (int i = 0; i < N - 1; i++) {
            for (int is = 0; is < N_SUB; is++) {
                const float *gv = &gain[i + 1].gv[(size_t) ivl[i * N_SUB + is] * (size_t) K];
                float gvl_ = gvl[i * N_SUB + is], evl_ = evl[i * N_SUB + is];
                for (int k = 0; k < K; k++) {
                    double gl = (double) (gvl_ * gv[k]); /* float product, then widened */
                    double el = (double) (evl_ * gv[k]);
                    if (fabs(gl) < 1e-3) {
                        Iv[k] = el * (1.0 + 0.5 * gl * (1.0 + 0.3333333333 * gl)) +
                                Iv[k] * (1.0 + gl * (1.0 + 0.5 * gl));
                    } else {
                        double exp_gl = exp(gl);
                        Iv[k] = el / gl * (exp_gl - 1.0) + Iv[k] * exp_gl;
                    }
                }
                if (dv != NULL && debug != NULL) { /* :559-566 */
                    int index = 3 * (N_SUB * i + is + 1) + 2;
                    debug[index] = 0.0f;
                    for (int k = 0; k < K; k++)
                        debug[index] += (float) (2 * Iv[k] * dv[k]);
                }
            }
        }
    } else { /* :569-581 */
        for (int k = 0; k < K; k++) {
            double gl = 0;
            for (int i = 0; i < N - 1; i++) {
                for (int is = 0; is < N_SUB; is++) {
                    double gv = (double) gain[i + 1].gv[(size_t) k + (size_t) ivl[i * N_SUB + is] * (size_t) K];
                    gl += (double) gvl[i * N_SUB + is] * gv;
                }
            }
            Iv[k] *= exp(gl);
        }
    }
    int neg = 0, nans = 0;
    for (int jj = 0; jj < K; jj++) {
        neg = neg || Iv[jj] < 0.0;
        nans = nans || Iv[jj] != Iv[jj];
    }
    return neg ? -2 : (nans ? -3 : 0);
}

int rt_oracle_calc_ray(const rtb200_ray *ray, int N, float dz0, const rtb200_gain_plane *gain,
                       const rtb200_seed *seed, int K, int method, float c, double *Iv,
                       rtb200_ray *ray2, float *gvl, float *evl, int32_t *ivl,
                       int *escaped_out, uint64_t *steps)
{
    return calc_ray_impl(ray, N, dz0, gain, seed, K, method, c, Iv, ray2, gvl, evl, ivl,
                         escaped_out, steps, NULL, NULL);
}

/* RayTrace_calc_ray with dv and debug given (the RAY_DEBUG trajectory used by
 * RayTrace::calc_ray_path, src/RayTraceImage.cpp:440-477): debug[3*n + 0..2] = x, y, I at the
 * n-th sub-segment boundary, n = 0 .. N_SUB*(N-1).  Note that with debug != NULL the reference
 * always takes the emission-style integration (:543). */
int rt_oracle_calc_ray_debug(const rtb200_ray *ray, int N, float dz0,
                             const rtb200_gain_plane *gain, const rtb200_seed *seed, int K,
                             int method, float c, const double *dv, double *Iv, rtb200_ray *ray2,
                             float *debug)
{
    const int S = (N - 1) * N_SUB;
    float *gvl = (float *) malloc(sizeof(float) * (size_t) (S > 0 ? S : 1));
    float *evl = (float *) malloc(sizeof(float) * (size_t) (S > 0 ? S : 1));
    int32_t *ivl = (int32_t *) malloc(sizeof(int32_t) * (size_t) (S > 0 ? S : 1));
    int rc = calc_ray_impl(ray, N, dz0, gain, seed, K, method, c, Iv, ray2, gvl, evl, ivl, NULL,
                           NULL, dv, debug);
    free(gvl);
    free(evl);
    free(ivl);
    return rc;
}

/* src/RayTraceImageCPU.cpp:11-16 */
static int get_index(int n, const double *x, double dx, double y)
{
    if (y < x[0] - 0.5 * dx || y > x[n - 1] + 0.5 * dx)
        return -1;
    return (int) rt_oracle_findfirstsingle(x, (size_t) n, y - 0.5 * dx);
}

/* src/RayTraceImageCPU.cpp:19-70 (RayTraceImageCPULoop).  Accumulates into image / I_ang. */
void rt_oracle_trace_rays(int N, const rtb200_beam *beam, const rtb200_gain_plane *gain,
                          const rtb200_seed *seed, int method, const rtb200_ray *rays,
                          size_t n_rays, double scale, double *image, double *I_ang,
                          unsigned *failure_code, rtb200_ray *failed, int max_failed,
                          int *n_failed, uint64_t *steps)
{
    const int K = beam->nv;
    const int S = (N - 1) * N_SUB;
    double *Iv = (double *) malloc(sizeof(double) * (size_t) (K > 0 ? K : 1));
    float *gvl = (float *) malloc(sizeof(float) * (size_t) (S > 0 ? S : 1));
    float *evl = (float *) malloc(sizeof(float) * (size_t) (S > 0 ? S : 1));
    int32_t *ivl = (int32_t *) malloc(sizeof(int32_t) * (size_t) (S > 0 ? S : 1));
    *failure_code = 0;
    int nf = 0;
    for (size_t it = 0; it < n_rays; ++it) {
        const rtb200_ray ray = rays[it];
        rtb200_ray ray2;
        int error = rt_oracle_calc_ray(&ray, N, (float) beam->dz, gain, seed, K, method, 0.5f, Iv,
                                       &ray2, gvl, evl, ivl, NULL, steps);
        if (error != 0) {
            if (failed && nf < max_failed)
                failed[nf] = ray;
            nf++;
            *failure_code |= 1u << (unsigned) (-error);
            continue;
        }
        if (method == 1) {
            ray2 = ray;
        } else {
            ray2.a = -ray2.a;
            ray2.b = -ray2.b;
            if (ray2.y < 0.0 && beam->y[0] >= 0.0)
                ray2.y = -ray2.y;
        }
        int i1 = get_index(beam->nx, beam->x, beam->dx, ray2.x);
        int i2 = get_index(beam->ny, beam->y, beam->dy, ray2.y);
        int i3 = get_index(beam->na, beam->a, beam->da, ray2.a);
        int i4 = get_index(beam->nb, beam->b, beam->db, ray2.b);
        if (i1 >= 0 && i2 >= 0) {
            double *Iv2 = &image[(size_t) beam->nv * (size_t) (i1 + i2 * beam->nx)];
            for (int iv = 0; iv < beam->nv; iv++)
                Iv2[iv] += Iv[iv] * scale;
        }
        if (i3 >= 0 && i4 >= 0) {
            double tmp = 0.0;
            for (int iv = 0; iv < beam->nv; iv++)
                tmp += 2.0 * beam->dv[iv] * Iv[iv];
            I_ang[i3 + i4 * beam->na] += tmp;
        }
    }
    if (n_failed)
        *n_failed = nf;
    free(Iv);
    free(gvl);
    free(evl);
    free(ivl);
}

/* src/RayTraceImage.cpp:237-242 */
static int check_grid(int N, double dx, const double *x)
{
    int error = 0;
    for (int i = 1; i < N; i++)
        error = error || (fabs((x[i] - x[i - 1]) - dx) > 1e-12 * dx);
    return error;
}

/* src/RayTraceImage.cpp:277-328: method / scale / ray enumeration. */
size_t rt_oracle_ray_count(const rtb200_problem *p)
{
    const rtb200_beam *g = p->seed != NULL ? p->seed_beam : p->euv_beam;
    long Nt = (long) g->nx * g->ny * g->na * g->nb;
    size_t n = 0;
    for (long it = 0; it < (Nt / p->N_parallel) + 1; ++it)
        if (p->N_start + it * p->N_parallel < Nt)
            n++;
    return n;
}

void rt_oracle_build_rays(const rtb200_problem *p, rtb200_ray *rays, int *method, double *scale)
{
    const rtb200_beam *e = p->euv_beam;
    const rtb200_beam *g = e;
    if (p->seed != NULL) { /* sizes and scale follow `seed` (:283-294) ... */
        g = p->seed_beam;
        *method = 2;
        *scale = (g->dx * g->dy * g->da * g->db) / (e->dx * e->dy);
    } else {
        *method = 1;
        *scale = 1.0;
    }
    const int N2[4] = { g->nx, g->ny, g->na, g->nb };
    if (p->seed_beam != NULL) /* ... while the coordinates follow `seed_beam` (:316-326) */
        g = p->seed_beam;
    int Nt = N2[0] * N2[1] * N2[2] * N2[3];
    size_t n = 0;
    for (int it = 0; it < (Nt / p->N_parallel) + 1; ++it) {
        int ijkm = p->N_start + it * p->N_parallel;
        if (ijkm >= Nt)
            continue;
        int m = ijkm % N2[3];
        int k = (ijkm / N2[3]) % N2[2];
        int j = (ijkm / (N2[2] * N2[3])) % N2[1];
        int i = ijkm / (N2[1] * N2[2] * N2[3]);
        rays[n].x = (float) g->x[i];
        rays[n].y = (float) g->y[j];
        rays[n].a = (float) g->a[k];
        rays[n].b = (float) g->b[m];
        n++;
    }
}

/* src/RayTraceImage.cpp:227-434 (create_image, method "cpu").  image / I_ang must be zeroed
 * by the caller (the reference callocs them). */
int rt_oracle_create_image(const rtb200_problem *p, unsigned flags, double *image, double *I_ang,
                           unsigned *failure_code, rtb200_ray *failed, int max_failed,
                           int *n_failed, uint64_t *steps)
{
    const rtb200_beam *e = p->euv_beam;
    if (!(flags & RTB200_FLAG_NO_LIMITS)) {
        if (p->N > RTB200_N_MAX || e->nv >= RTB200_K_MAX)
            return RTB200_ERR_LIMITS;
    }
    if (check_grid(e->nx, e->dx, e->x) || check_grid(e->ny, e->dy, e->y) ||
        check_grid(e->na, e->da, e->a) || check_grid(e->nb, e->db, e->b))
        return RTB200_ERR_GRID;
    if (p->seed_beam != NULL) {
        const rtb200_beam *sb = p->seed_beam;
        if (check_grid(sb->nx, sb->dx, sb->x) || check_grid(sb->ny, sb->dy, sb->y) ||
            check_grid(sb->na, sb->da, sb->a) || check_grid(sb->nb, sb->db, sb->b))
            return RTB200_ERR_GRID;
        if ((e->y[0] >= 0.0) != (sb->y[0] >= 0.0))
            return RTB200_ERR_GRID;
    }
    size_t n = rt_oracle_ray_count(p);
    rtb200_ray *rays = (rtb200_ray *) malloc(sizeof(rtb200_ray) * (n > 0 ? n : 1));
    int method;
    double scale;
    rt_oracle_build_rays(p, rays, &method, &scale);
    rt_oracle_trace_rays(p->N, e, p->gain, p->seed, method, rays, n, scale, image, I_ang,
                         failure_code, failed, max_failed, n_failed, steps);
    free(rays);
    return *failure_code ? RTB200_RAYS_FAILED : RTB200_OK;
}

/* src/RayTraceImage.cpp:89-134 (RayTraceImageThreadLoop): contiguous chunks of
 * size/N_threads + 1 rays, one thread each with private zeroed partial images, summed in
 * thread order after the joins. */
typedef struct {
    const rtb200_problem *p;
    const rtb200_ray *rays;
    size_t n;
    int method;
    double scale;
    double *image, *I_ang;
    unsigned failure_code;
} thread_arg;

static void *thread_main(void *v)
{
    thread_arg *a = (thread_arg *) v;
    int nf = 0;
    rt_oracle_trace_rays(a->p->N, a->p->euv_beam, a->p->gain, a->p->seed, a->method, a->rays,
                         a->n, a->scale, a->image, a->I_ang, &a->failure_code, NULL, 0, &nf,
                         NULL);
    return NULL;
}

int rt_oracle_create_image_threads(const rtb200_problem *p, unsigned flags, int n_threads,
                                   double *image, double *I_ang, unsigned *failure_code,
                                   size_t *rays_done)
{
    const rtb200_beam *e = p->euv_beam;
    if (!(flags & RTB200_FLAG_NO_LIMITS)) {
        if (p->N > RTB200_N_MAX || e->nv >= RTB200_K_MAX)
            return RTB200_ERR_LIMITS;
    }
    if (n_threads < 1)
        n_threads = 1;
    size_t n = rt_oracle_ray_count(p);
    rtb200_ray *rays = (rtb200_ray *) malloc(sizeof(rtb200_ray) * (n > 0 ? n : 1));
    int method;
    double scale;
    rt_oracle_build_rays(p, rays, &method, &scale);
    const size_t n_img = (size_t) e->nx * e->ny * e->nv, n_ang = (size_t) e->na * e->nb;
    thread_arg *args = (thread_arg *) calloc((size_t) n_threads, sizeof(thread_arg));
    pthread_t *tid = (pthread_t *) calloc((size_t) n_threads, sizeof(pthread_t));
    size_t chunk = n / (size_t) n_threads + 1, j = 0;
    for (int t = 0; t < n_threads; t++) {
        size_t m = chunk < n - j ? chunk : n - j;
        args[t].p = p;
        args[t].rays = rays + j;
        args[t].n = m;
        args[t].method = method;
        args[t].scale = scale;
        args[t].image = (double *) calloc(n_img, sizeof(double));
        args[t].I_ang = (double *) calloc(n_ang, sizeof(double));
        j += m;
        pthread_create(&tid[t], NULL, thread_main, &args[t]);
    }
    *failure_code = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(tid[t], NULL);
        for (size_t q = 0; q < n_img; q++)
            image[q] += args[t].image[q];
        for (size_t q = 0; q < n_ang; q++)
            I_ang[q] += args[t].I_ang[q];
        *failure_code |= args[t].failure_code;
        free(args[t].image);
        free(args[t].I_ang);
    }
    if (rays_done)
        *rays_done = n;
    free(args);
    free(tid);
    free(rays);
    return *failure_code ? RTB200_RAYS_FAILED : RTB200_OK;
}
