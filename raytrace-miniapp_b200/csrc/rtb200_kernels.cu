// rtb200_kernels.cu — hand-written sm_100a kernels of the image-formation path.
//
//   march_kernel               one thread per ray: refractive march through the gain planes
//                              (FP32, bit-exact with the reference), hands gvl/evl/ivl per
//                              (segment, sub-segment) to the integration through L2.
//   integrate_ase_owner_kernel one CTA per source pixel, one warp per ray, lanes = frequency
//                              bins: ASE gain + emission integration (FP64), per-pixel spectrum
//                              accumulated in registers (no atomics), I_ang by warp-shuffle
//                              reduction + one FP64 atomic per ray.
//   integrate_scatter_kernel   one warp per ray, lanes = frequency bins: seeded (gain-only) and
//                              list-mode rays, binning by the exit ray with FP64 atomics;
//                              also the per-ray dump used by rtb200_calc_rays.
//
// There are no tensor-core instructions here on purpose: the path is not a dense contraction
// (SURVEY.md §8d); the binding unit is the FP64 pipe.
#include <cuda_runtime.h>
#include <math.h>

#include "rtb200_kernels.cuh"

namespace rtb {

// ------------------------------------------------------------------------------------------
// slot <-> ray decoding (src/RayTraceImage.cpp:300-328: ijkm = N_start + it*N_parallel,
// m = b fastest, k = a, j = y, i = x slowest).  Work is organised by SOURCE PIXEL p = i + j*snx;
// the rays of one pixel are ab = ab0 + t*n_parallel < sna*snb.
// ------------------------------------------------------------------------------------------
struct PixelRays {
    int i, j;
    int ab0;
    int cnt;
};

__device__ __forceinline__ PixelRays pixel_rays(const DevProblem &P, long long p)
{
    PixelRays r;
    r.i = (int) (p % P.snx);
    r.j = (int) (p / P.snx);
    const long long AB = (long long) P.sna * P.snb;
    const long long base = ((long long) r.i * P.sny + r.j) * AB;
    const long long d = base - P.n_start;
    long long ab0;
    if (d >= 0) {
        const long long rem = d % P.n_parallel;
        ab0 = rem == 0 ? 0 : P.n_parallel - rem;
    } else {
        ab0 = -d;
    }
    r.ab0 = (int) (ab0 < AB ? ab0 : AB);
    r.cnt = ab0 < AB ? (int) ((AB - 1 - ab0) / P.n_parallel + 1) : 0;
    return r;
}

__device__ __forceinline__ void report_failure(FailState *fail, int code, float x, float y,
                                               float a, float b)
{
    atomicOr(&fail->failure_code, 1u << code);
    const unsigned idx = atomicAdd(&fail->n_failed, 1u);
    if (idx < 32u) {
        fail->failed[4 * idx + 0] = x;
        fail->failed[4 * idx + 1] = y;
        fail->failed[4 * idx + 2] = a;
        fail->failed[4 * idx + 3] = b;
    }
}

// atanf as glibc 2.39 evaluates it (sysdeps/ieee754/flt-32/s_atanf.c, the fdlibm algorithm in
// float arithmetic).  ray2.a = atan(s.x/s.z)*1e3f (RayTraceImageHelper.h:520-521) feeds a
// discrete bin index in seeded mode, so it is reproduced operation by operation.
__device__ float atanf_fdlibm(float x)
{
    const float atanhi[4] = { 4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f,
                              1.5707962513e+00f };
    const float atanlo[4] = { 5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f,
                              7.5497894159e-08f };
    const float aT[11] = { 3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f,
                           -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                           6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f,
                           -3.6531571299e-02f, 1.6285819933e-02f };
    const int hx = __float_as_int(x);
    const int ix = hx & 0x7fffffff;
    int id;
    if (ix >= 0x4c800000) { // |x| >= 2^26
        if (ix > 0x7f800000)
            return fadd(x, x); // NaN
        return hx > 0 ? fadd(atanhi[3], atanlo[3]) : -fadd(atanhi[3], atanlo[3]);
    }
    if (ix < 0x3ee00000) {     // |x| < 0.4375
        if (ix < 0x31000000) { // |x| < 2^-29
            return x;
        }
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {     // |x| < 1.1875
            if (ix < 0x3f300000) { // 7/16 <= |x| < 11/16
                id = 0;
                x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x));
            } else { // 11/16 <= |x| < 19/16
                id = 1;
                x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f));
            }
        } else {
            if (ix < 0x401c0000) { // |x| < 2.4375
                id = 2;
                x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x)));
            } else { // 2.4375 <= |x| < 2^26
                id = 3;
                x = fdiv(-1.0f, x);
            }
        }
    }
    const float z = fmul(x, x);
    const float w = fmul(z, z);
    // s1 = z*(aT[0]+w*(aT[2]+w*(aT[4]+w*(aT[6]+w*(aT[8]+w*aT[10])))))
    float s1 = fadd(aT[8], fmul(w, aT[10]));
    s1 = fadd(aT[6], fmul(w, s1));
    s1 = fadd(aT[4], fmul(w, s1));
    s1 = fadd(aT[2], fmul(w, s1));
    s1 = fmul(z, fadd(aT[0], fmul(w, s1)));
    // s2 = w*(aT[1]+w*(aT[3]+w*(aT[5]+w*(aT[7]+w*aT[9]))))
    float s2 = fadd(aT[7], fmul(w, aT[9]));
    s2 = fadd(aT[5], fmul(w, s2));
    s2 = fadd(aT[3], fmul(w, s2));
    s2 = fmul(w, fadd(aT[1], fmul(w, s2)));
    if (id < 0)
        return fsub(x, fmul(x, fadd(s1, s2)));
    const float zz = fsub(atanhi[id], fsub(fsub(fmul(x, fadd(s1, s2)), atanlo[id]), x));
    return hx < 0 ? -zz : zz;
}

// ------------------------------------------------------------------------------------------
// march
// ------------------------------------------------------------------------------------------
struct GlobalSink {
    SegRec *seg;
    __device__ __forceinline__ void operator()(int idx, float gvl, float evl, int cell) const
    {
        int4 v;
        v.x = __float_as_int(gvl);
        v.y = __float_as_int(evl);
        v.z = cell;
        v.w = 0;
        *reinterpret_cast<int4 *>(&seg[idx]) = v;
    }
};

template <bool LIST, bool COUNT>
__global__ void __launch_bounds__(128) march_kernel(const DevProblem P, const Chunk c,
                                                    const Handoff h, FailState *fail)
{
    const long long L = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int S = (P.N - 1) * RTB_N_SUB;
    float rx, ry, ra, rb, ta, tb;
    if (LIST) {
        if (L >= c.ray1 - c.ray0)
            return;
        const float4 r = __ldg(&c.rays[c.ray0 + L]);
        const float2 t = __ldg(&c.tans[c.ray0 + L]);
        rx = r.x, ry = r.y, ra = r.z, rb = r.w, ta = t.x, tb = t.y;
    } else {
        const long long npix = c.pix1 - c.pix0;
        if (L >= npix * P.ab_max)
            return;
        const long long p = c.pix0 + L / P.ab_max;
        const int t = (int) (L % P.ab_max);
        const PixelRays pr = pixel_rays(P, p);
        if (t >= pr.cnt) {
            h.meta[L] = RTB_META_INACTIVE;
            return;
        }
        const int ab = pr.ab0 + t * (int) P.n_parallel;
        const int k = ab / P.snb, m = ab % P.snb;
        rx = __ldg(&P.sxf[pr.i]);
        ry = __ldg(&P.syf[pr.j]);
        ra = __ldg(&P.saf[k]);
        rb = __ldg(&P.sbf[m]);
        ta = __ldg(&P.tanA[k]);
        tb = __ldg(&P.tanB[m]);
    }
    GlobalSink sink{ h.seg + L * S };
    MarchResult res;
    unsigned steps = 0;
    march_ray(P.planes, P.N, P.method, P.dz0, P.c, P.use_emis != 0, rx, ry, ta, tb, sink, res,
              steps);
    unsigned meta = (unsigned) res.seg_lo | ((unsigned) res.seg_hi << 12);
    if (res.escaped)
        meta |= RTB_META_ESCAPED;
    if (lt_0p01(fmul(res.s.z, res.s.z))) { // error -1 (RayTraceImageHelper.h:515-516)
        meta |= RTB_META_INVALID;
        report_failure(fail, 1, rx, ry, ra, rb);
    } else if (h.exit_ray) {
        float4 e;
        e.x = res.pos.x;
        e.y = res.pos.y;
        e.z = fmul(atanf_fdlibm(fdiv(res.s.x, res.s.z)), 1e3f);
        e.w = fmul(atanf_fdlibm(fdiv(res.s.y, res.s.z)), 1e3f);
        h.exit_ray[L] = e;
    }
    h.meta[L] = meta;
    if (COUNT) {
        unsigned tot = steps;
        for (int o = 16; o > 0; o >>= 1)
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if ((threadIdx.x & 31) == 0)
            atomicAdd(&fail->march_steps, (unsigned long long) tot);
    }
}

void launch_march(const DevProblem &P, const Chunk &c, bool list_mode, const Handoff &h,
                  FailState *fail, bool count_steps, cudaStream_t st)
{
    const long long n = list_mode ? (c.ray1 - c.ray0) : (c.pix1 - c.pix0) * P.ab_max;
    if (n <= 0)
        return;
    const int threads = 128;
    const unsigned blocks = (unsigned) ((n + threads - 1) / threads);
    if (list_mode) {
        if (count_steps)
            march_kernel<true, true><<<blocks, threads, 0, st>>>(P, c, h, fail);
        else
            march_kernel<true, false><<<blocks, threads, 0, st>>>(P, c, h, fail);
    } else {
        if (count_steps)
            march_kernel<false, true><<<blocks, threads, 0, st>>>(P, c, h, fail);
        else
            march_kernel<false, false><<<blocks, threads, 0, st>>>(P, c, h, fail);
    }
}

// ------------------------------------------------------------------------------------------
// frequency integration
// ------------------------------------------------------------------------------------------

// One (segment, sub-segment) update of one frequency bin with gain and spontaneous emission
// (RayTraceImageHelper.h:549-557).  gl, el are float products widened to double.
__device__ __forceinline__ double ase_update(double Iv, double gl, double el)
{
    if (fabs(gl) < 1e-3) {
        return el * (1.0 + 0.5 * gl * (1.0 + 0.3333333333 * gl)) +
               Iv * (1.0 + gl * (1.0 + 0.5 * gl));
    }
    const double e = exp(gl);
    return el / gl * (e - 1.0) + Iv * e;
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Integrates one ray's frequency bins k = lane + 32*q, q < KS, over its visited records.
// Returns the failure code of the ray (0, 2 = negative, 3 = NaN), warp-uniform.
template <int KS>
__device__ __forceinline__ int integrate_ray(const DevProblem &P, const SegRec *seg, unsigned meta,
                                             int lane, int kbase, double (&Iv)[KS])
{
    const int lo = meta & 0xfff, hi = (meta >> 12) & 0xfff;
    const int K = P.K;
    if (P.use_emis) {
        for (int pl = lo / RTB_N_SUB; pl < P.N - 1; pl++) {
            const float *gvp = P.planes[pl + 1].gv;
            for (int is = 0; is < RTB_N_SUB; is++) {
                const int s = pl * RTB_N_SUB + is;
                if (s < lo || s >= hi)
                    continue;
                const int4 rv = __ldg(reinterpret_cast<const int4 *>(&seg[s]));
                const float gvl = __int_as_float(rv.x), evl = __int_as_float(rv.y);
                if (gvl == 0.0f && evl == 0.0f)
                    continue; // gl = el = 0: the update is the identity
                const float *row = gvp + (size_t) rv.z * K + kbase;
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    const int k = lane + 32 * q;
                    if (kbase + k < K) {
                        const float g = __ldg(row + k);
                        const double gl = (double) __fmul_rn(gvl, g);
                        const double el = (double) __fmul_rn(evl, g);
                        Iv[q] = ase_update(Iv[q], gl, el);
                    }
                }
            }
        }
    } else {
        // gain only (:569-581): Iv[k] *= exp(sum_s (double)gvl_s * (double)gv_s[k])
        double gl[KS];
#pragma unroll
        for (int q = 0; q < KS; q++)
            gl[q] = 0.0;
        for (int pl = lo / RTB_N_SUB; pl < P.N - 1; pl++) {
            const float *gvp = P.planes[pl + 1].gv;
            for (int is = 0; is < RTB_N_SUB; is++) {
                const int s = pl * RTB_N_SUB + is;
                if (s < lo || s >= hi)
                    continue;
                const int4 rv = __ldg(reinterpret_cast<const int4 *>(&seg[s]));
                const double gvl = (double) __int_as_float(rv.x);
                const float *row = gvp + (size_t) rv.z * K + kbase;
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    const int k = lane + 32 * q;
                    if (kbase + k < K)
                        gl[q] = __dadd_rn(gl[q], __dmul_rn(gvl, (double) __ldg(row + k)));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < KS; q++)
            Iv[q] *= exp(gl[q]);
    }
    bool neg = false, nan = false;
#pragma unroll
    for (int q = 0; q < KS; q++) {
        neg = neg || Iv[q] < 0.0;
        nan = nan || Iv[q] != Iv[q];
    }
    const bool any_neg = __any_sync(0xffffffffu, neg);
    const bool any_nan = __any_sync(0xffffffffu, nan);
    return any_neg ? 2 : (any_nan ? 3 : 0);
}

#define RTB_OWNER_WARPS 8

template <int KS>
__global__ void __launch_bounds__(RTB_OWNER_WARPS * 32)
    integrate_ase_owner_kernel(const DevProblem P, const Chunk c, const Handoff h, const Outputs o)
{
    __shared__ double part[RTB_OWNER_WARPS][KS * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p = c.pix0 + blockIdx.x;
    const PixelRays pr = pixel_rays(P, p);
    const int S = (P.N - 1) * RTB_N_SUB;
    const int K = P.K;
    const long long slot0 = (long long) blockIdx.x * P.ab_max;
    double pix[KS], dv2[KS];
#pragma unroll
    for (int q = 0; q < KS; q++) {
        pix[q] = 0.0;
        const int k = lane + 32 * q;
        dv2[q] = k < K ? __ldg(&P.dv2[k]) : 0.0;
    }
    for (int t = warp; t < pr.cnt; t += RTB_OWNER_WARPS) {
        const long long slot = slot0 + t;
        const unsigned meta = __ldg(&h.meta[slot]);
        if (meta & RTB_META_INVALID)
            continue; // error -1, reported by the march
        double Iv[KS];
#pragma unroll
        for (int q = 0; q < KS; q++)
            Iv[q] = 0.0;
        const int code = integrate_ray<KS>(P, h.seg + slot * S, meta, lane, 0, Iv);
        const int ab = pr.ab0 + t * (int) P.n_parallel;
        const int ka = ab / P.snb, m = ab % P.snb;
        if (code != 0) {
            if (lane == 0)
                report_failure(o.fail, code, P.sxf[pr.i], P.syf[pr.j], P.saf[ka], P.sbf[m]);
            continue;
        }
        double w = 0.0;
#pragma unroll
        for (int q = 0; q < KS; q++) {
            w += dv2[q] * Iv[q];
            pix[q] += Iv[q] * P.scale;
        }
        w = warp_sum(w);
        const int ba = __ldg(&P.binA[ka]), bb = __ldg(&P.binB[m]);
        if (lane == 0 && ba >= 0 && bb >= 0)
            atomicAdd(&o.I_ang[ba + bb * P.na], w);
    }
#pragma unroll
    for (int q = 0; q < KS; q++)
        part[warp][q * 32 + lane] = pix[q];
    __syncthreads();
    const int pi = __ldg(&P.pixI[pr.i]), pj = __ldg(&P.pixJ[pr.j]);
    if (pi < 0 || pj < 0)
        return;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < RTB_OWNER_WARPS; w++)
            sum += part[w][k];
        o.image[(size_t) K * ((size_t) pi + (size_t) pj * P.nx) + k] = sum;
    }
}

void launch_integrate_ase_owner(const DevProblem &P, const Chunk &c, const Handoff &h,
                                const Outputs &o, cudaStream_t st)
{
    const long long npix = c.pix1 - c.pix0;
    if (npix <= 0)
        return;
    const unsigned blocks = (unsigned) npix;
    const int threads = RTB_OWNER_WARPS * 32;
    const int ks = (P.K + 31) / 32;
    switch (ks) {
    case 1: integrate_ase_owner_kernel<1><<<blocks, threads, 0, st>>>(P, c, h, o); break;
    case 2: integrate_ase_owner_kernel<2><<<blocks, threads, 0, st>>>(P, c, h, o); break;
    case 3: integrate_ase_owner_kernel<3><<<blocks, threads, 0, st>>>(P, c, h, o); break;
    default: integrate_ase_owner_kernel<4><<<blocks, threads, 0, st>>>(P, c, h, o); break;
    }
}

// getIndex (RayTraceImageCPU.cpp:11-16) on the device, for the exit ray.
__device__ __forceinline__ int dev_get_index(int n, const double *x, double dx, double y)
{
    if (y < __ldg(&x[0]) - 0.5 * dx || y > __ldg(&x[n - 1]) + 0.5 * dx)
        return -1;
    const double Y = y - 0.5 * dx;
    if (Y < __ldg(&x[0]))
        return 0;
    if (Y > __ldg(&x[n - 1]))
        return n;
    int lo = 0, hi = n - 1;
    while (hi - lo != 1) {
        const int mid = (hi + lo) / 2;
        if (__ldg(&x[mid]) >= Y)
            hi = mid;
        else
            lo = mid;
    }
    return hi;
}

// One warp per ray slot; K is covered in passes of 64 bins.  Handles both integration modes,
// both ray sources, scatter binning and the per-ray dumps.
template <bool LIST>
__global__ void __launch_bounds__(256)
    integrate_scatter_kernel(const DevProblem P, const Chunk c, const Handoff h, const Outputs o)
{
    constexpr int KS = 2;
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long) gridDim.x * blockDim.x) >> 5;
    const long long n_slots = LIST ? (c.ray1 - c.ray0) : (c.pix1 - c.pix0) * P.ab_max;
    const int S = (P.N - 1) * RTB_N_SUB;
    const int K = P.K;
    for (long long slot = warp_id; slot < n_slots; slot += n_warps) {
        const unsigned meta = __ldg(&h.meta[slot]);
        if (meta & RTB_META_INACTIVE)
            continue;
        float rx, ry, ra, rb;
        double f = 0.0; // seed amplitude
        if (LIST) {
            const float4 r = __ldg(&c.rays[c.ray0 + slot]);
            rx = r.x, ry = r.y, ra = r.z, rb = r.w;
        } else {
            const long long p = c.pix0 + slot / P.ab_max;
            const int t = (int) (slot % P.ab_max);
            const PixelRays pr = pixel_rays(P, p);
            const int ab = pr.ab0 + t * (int) P.n_parallel;
            const int ka = ab / P.snb, m = ab % P.snb;
            rx = __ldg(&P.sxf[pr.i]);
            ry = __ldg(&P.syf[pr.j]);
            ra = __ldg(&P.saf[ka]);
            rb = __ldg(&P.sbf[m]);
            if (P.seed_fx && !(meta & RTB_META_ESCAPED)) {
                // calc_seed_inline (:230-247) from the per-index tables
                const double fx = __ldg(&P.seed_fx[pr.i]), fy = __ldg(&P.seed_fy[pr.j]);
                const double fa = __ldg(&P.seed_fa[ka]), fb = __ldg(&P.seed_fb[m]);
                if (fx == fx && fy == fy && fa == fa && fb == fb) {
                    f = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(P.seed_f0, fx), fy), fa), fb);
                    f = f < 0.0 ? 0.0 : f;
                }
            }
        }
        const bool invalid = (meta & RTB_META_INVALID) != 0;
        int code = invalid ? 1 : 0;
        // destination cells
        int i1 = -1, i2 = -1, i3 = -1, i4 = -1;
        if (!invalid && (o.image || o.I_ang)) {
            float bx = rx, by = ry, ba = ra, bb = rb;
            if (P.method != 1) { // forward: bin by the exit ray (RayTraceImageCPU.cpp:40-49)
                const float4 e = h.exit_ray[slot];
                bx = e.x;
                by = e.y;
                ba = -e.z;
                bb = -e.w;
                if (by < 0.0f && P.y_mirror)
                    by = -by;
            }
            i1 = dev_get_index(P.nx, P.ex, P.edx, (double) bx);
            i2 = dev_get_index(P.ny, P.ey, P.edy, (double) by);
            i3 = dev_get_index(P.na, P.ea, P.eda, (double) ba);
            i4 = dev_get_index(P.nb, P.eb, P.edb, (double) bb);
        }
        double w = 0.0;
        bool bad = false;
        for (int kbase = 0; kbase < K && !bad; kbase += 32 * KS) {
            double Iv[KS];
#pragma unroll
            for (int q = 0; q < KS; q++) {
                const int k = kbase + lane + 32 * q;
                Iv[q] = (f != 0.0 && k < K) ? __dmul_rn(f, __ldg(&P.seed_fv[k])) : 0.0;
            }
            if (!invalid) {
                const int cc = integrate_ray<KS>(P, h.seg + slot * S, meta, lane, kbase, Iv);
                if (cc != 0) {
                    code = code == 0 ? cc : (cc < code ? cc : code); // negative (2) wins over NaN (3)
                }
            }
            if (o.Iv) {
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    const int k = kbase + lane + 32 * q;
                    if (k < K)
                        o.Iv[(size_t) slot * K + k] = Iv[q];
                }
            }
            // Binning is deferred until the whole ray is known to be valid when K needs
            // several passes; with one pass (K <= 64) it happens right here.
            if (K <= 32 * KS) {
                if (code == 0 && !invalid) {
#pragma unroll
                    for (int q = 0; q < KS; q++) {
                        const int k = lane + 32 * q;
                        if (k < K) {
                            w += __ldg(&P.dv2[k]) * Iv[q];
                            if (o.image && i1 >= 0 && i2 >= 0)
                                atomicAdd(&o.image[(size_t) K * ((size_t) i1 + (size_t) i2 * P.nx) + k],
                                          Iv[q] * P.scale);
                        }
                    }
                }
            }
        }
        if (K > 32 * KS && code == 0 && !invalid && (o.image || o.I_ang)) {
            // second sweep: recompute and bin (rays are independent, so this is exact)
            for (int kbase = 0; kbase < K; kbase += 32 * KS) {
                double Iv[KS];
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    const int k = kbase + lane + 32 * q;
                    Iv[q] = (f != 0.0 && k < K) ? __dmul_rn(f, __ldg(&P.seed_fv[k])) : 0.0;
                }
                integrate_ray<KS>(P, h.seg + slot * S, meta, lane, kbase, Iv);
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    const int k = kbase + lane + 32 * q;
                    if (k < K) {
                        w += __ldg(&P.dv2[k]) * Iv[q];
                        if (o.image && i1 >= 0 && i2 >= 0)
                            atomicAdd(&o.image[(size_t) K * ((size_t) i1 + (size_t) i2 * P.nx) + k],
                                      Iv[q] * P.scale);
                    }
                }
            }
        }
        if (code == 0 && !invalid && o.I_ang) {
            w = warp_sum(w);
            if (lane == 0 && i3 >= 0 && i4 >= 0)
                atomicAdd(&o.I_ang[i3 + i4 * P.na], w);
        }
        if (lane == 0) {
            if (o.error)
                o.error[slot] = -code;
            if (code >= 2)
                report_failure(o.fail, code, rx, ry, ra, rb);
        }
    }
}

void launch_integrate_scatter(const DevProblem &P, const Chunk &c, bool list_mode,
                              const Handoff &h, const Outputs &o, cudaStream_t st)
{
    const long long n = list_mode ? (c.ray1 - c.ray0) : (c.pix1 - c.pix0) * P.ab_max;
    if (n <= 0)
        return;
    const int threads = 256;
    long long blocks = (n + 7) / 8;
    const long long cap = 148LL * 8 * 16;
    if (blocks > cap)
        blocks = cap;
    if (list_mode)
        integrate_scatter_kernel<true><<<(unsigned) blocks, threads, 0, st>>>(P, c, h, o);
    else
        integrate_scatter_kernel<false><<<(unsigned) blocks, threads, 0, st>>>(P, c, h, o);
}

// ------------------------------------------------------------------------------------------
// FP64 peak micro-benchmark: 8 independent DFMA chains per thread.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters)
{
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    double a4 = a0 + 4e-3, a5 = a0 + 5e-3, a6 = a0 + 6e-3, a7 = a0 + 7e-3;
    const double m = 0.999999999, b = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, b);
        a1 = fma(a1, m, b);
        a2 = fma(a2, m, b);
        a3 = fma(a3, m, b);
        a4 = fma(a4, m, b);
        a5 = fma(a5, m, b);
        a6 = fma(a6, m, b);
        a7 = fma(a7, m, b);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678)
        out[0] = r; // never true; keeps the chains alive
}

void launch_fp64_peak(double *out, int iters, cudaStream_t st, int *blocks, int *threads)
{
    *blocks = 148 * 8;
    *threads = 256;
    fp64_peak_kernel<<<*blocks, *threads, 0, st>>>(out, iters);
}

} // namespace rtb
