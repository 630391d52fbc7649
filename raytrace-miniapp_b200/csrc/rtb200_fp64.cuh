// rtb200_fp64.cuh — the FP64 frequency-bin update of the integration kernels.
//
// One update of one frequency bin over one (length segment, sub-segment)
// (src/common/RayTraceImageHelper.h:549-557):
//     gl = gvl*gv[k], el = evl*gv[k]                  (float products, widened afterwards)
//     |gl| < 1e-3 :  I = el*(1 + gl/2*(1 + 0.3333333333*gl)) + I*(1 + gl*(1 + gl/2))
//     otherwise   :  I = el/gl*(exp(gl) - 1) + I*exp(gl)
// The stated tolerance of the path is 1e-10 relative on the image, so exp and the division are
// evaluated to ~2e-16 / ~2e-14 relative with a fraction of the instructions of the IEEE library
// routines (which cost 18 + 10 FP64 instructions plus slow-path branches):
//   exp : x = (128 m + j) ln2/128 + r,  exp(x) = 2^m * 2^(j/128) * (1 + r*q(r)),  |r| <= ln2/256,
//         2^(j/128) from a 128-entry table, q of degree 3 (the degree-5 Taylor polynomial with
//         its last term economised, tools/gen_exptab.py), ONE correctly rounded ln2/128 in the
//         reduction (abs error of r: 1.8e-19 * |128 m + j|, i.e. 3e-17 * |x|)
//                                                               -> 8 FP64 instructions;
//   1/gl: single-precision reciprocal seed + one Newton step in double -> 2 FP64 instructions.
// Written __host__ __device__ so tests/hostsim can check it against libm without a GPU.
#pragma once
#include "rtb200_exptab.h"
#include "rtb200_math.cuh"

namespace rtb {

#define RTB_EXP_MAGIC 6755399441055744.0 /* 1.5 * 2^52: round-to-nearest-integer by addition */

// Constants of the update, as an array so that the kernels can keep them in the kernel
// PARAMETER bank (DevProblem::kfp): c[0x0][..] is directly addressable as an FP64 operand,
// which avoids materialising 64-bit immediates (2 UMOV per use) or LDC loads in the hot loop.
#define RTB_K_N_OVER_LN2 0
#define RTB_K_LN2_OVER_N 1 /* negated */
#define RTB_K_C1 2
#define RTB_K_C3 3
#define RTB_K_C4 4
#define RTB_K_THIRD 5
#define RTB_K_COUNT 8
#define RTB_K_VALUES                                                                             \
    RTB_EXP_N_OVER_LN2, -RTB_EXP_LN2_OVER_N, RTB_EXP_C1, RTB_EXP_C3, RTB_EXP_C4, 0.3333333333,   \
        0.0, 0.0
#define RTB_EXP_TABLE_SIZE (1 << RTB_EXP_TABLE_BITS)

RTB_HD double fma64(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

RTB_HD int lo32(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int) (unsigned) (b & 0xffffffffLL);
#endif
}

RTB_HD double scale_pow2(double e, int m) // e * 2^m for a normal e and a normal result
{
#if defined(__CUDA_ARCH__)
#ifndef RTB_NO_MAD_POW2 // one multiply-add on the high word (left alone the compiler emits shift + mask + add; -1.6 %)
    int hi;
    asm("mad.lo.s32 %0, %1, 0x100000, %2;" : "=r"(hi) : "r"(m), "r"(__double2hiint(e)));
    return __hiloint2double(hi, __double2loint(e));
#else
    return __hiloint2double(__double2hiint(e) + (m << 20), __double2loint(e));
#endif
#else
    long long b;
    memcpy(&b, &e, 8);
    b += (long long) m << 52;
    memcpy(&e, &b, 8);
    return e;
#endif
}

// The constants and the 2^(j/64) table are reached through a provider type so that the same
// arithmetic runs with plain arrays on the host (tests/hostsim) and, in the hot kernel, with
// the constants pinned in registers and the table read by explicit shared-memory loads.
struct ArrayConsts {
    const double *kc; // RTB_K_VALUES
    const double *T;  // 2^(j/128), j = 0..127
    RTB_HD double l2e() const { return kc[RTB_K_N_OVER_LN2]; }
    RTB_HD double nln2() const { return kc[RTB_K_LN2_OVER_N]; }
    RTB_HD double c1() const { return kc[RTB_K_C1]; }
    RTB_HD double c3() const { return kc[RTB_K_C3]; }
    RTB_HD double c4() const { return kc[RTB_K_C4]; }
    RTB_HD double third() const { return kc[RTB_K_THIRD]; }
    RTB_HD double tab(int j) const { return T[j]; }
};

// exp(x) for |x| < 700 (the caller routes everything else to the library exp): relative error
// <= 4e-16 + 3.4e-17*|x| (tests/test_fp64_update.py), against the path's tolerance of 1e-10.
template <class KC>
RTB_HD double exp_core(double x, const KC &C)
{
    const double t = fma64(x, C.l2e(), RTB_EXP_MAGIC);
    const int n = lo32(t);
    const double tn = t - RTB_EXP_MAGIC;
    const double r = fma64(tn, C.nln2(), x);
    double q = fma64(r, C.c4(), C.c3());
    q = fma64(r, q, 0.5);
    q = fma64(r, q, C.c1());
    const double Tj = C.tab(n & (RTB_EXP_TABLE_SIZE - 1));
    const double e = fma64(Tj, r * q, Tj);
    return scale_pow2(e, n >> RTB_EXP_TABLE_BITS);
}

template <class KC>
RTB_HD double exp_any(double x, const KC &C)
{
    if (fabs(x) < 700.0)
        return exp_core(x, C);
    return exp(x); // overflow / underflow / NaN: the library routine's semantics
}

// The |gl| < 1e-3 branch (7 FP64 instructions).
template <class KC>
RTB_HD double ase_update_small(double Iv, double gl, double el, const KC &C)
{
    const double a = fma64(gl, C.third(), 1.0);
    const double c = fma64(0.5 * gl, a, 1.0);
    const double d = fma64(gl, 0.5, 1.0);
    const double e = fma64(gl, d, 1.0);
    return fma64(el, c, Iv * e);
}

// The exp branch for |gl| < 700 (13 FP64 instructions); rcp_seed is a single-precision
// approximation of 1/gl.  el/gl*(e - 1) + Iv*e is evaluated as u*e + (Iv*e - u), u = el/gl.
template <class KC>
RTB_HD double ase_update_large(double Iv, double gl, double el, float rcp_seed, const KC &C)
{
    const double e = exp_core(gl, C);
    const double r0 = (double) rcp_seed;
    const double r1 = fma64(r0, fma64(-gl, r0, 1.0), r0); // 1/gl to ~2^-46
    const double u = el * r1;
    return fma64(u, e, fma64(Iv, e, -u));
}

} // namespace rtb
