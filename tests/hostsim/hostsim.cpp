// hostsim.cpp — TEST-ONLY build of the march's device functions for the host.
//
// rtb200_march.cuh / rtb200_pack.h are written as __host__ __device__ / plain C++ so that the
// exact same source the GPU runs can be unit-tested here, without a GPU, against the CPU oracle
// (tests/test_march_hostsim.py).  This library is NOT part of librtb200.so, is not reachable
// from the product API and is not a CPU fallback: it exists so that `pytest -m "not gpu"`
// can check the march logic (index search, sub-segment bookkeeping, packing) bit for bit.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../raytrace-miniapp_b200/csrc/rtb200_fp64.cuh"
#include "../../raytrace-miniapp_b200/csrc/rtb200_march_flat.cuh"
#include "../../raytrace-miniapp_b200/csrc/rtb200_pack.h"

using namespace rtb;

namespace {
struct ArraySink {
    float *gvl, *evl;
    int *ivl;
    void operator()(int idx, float g, float e, int cell) const
    {
        gvl[idx] = g;
        evl[idx] = e;
        ivl[idx] = cell;
    }
    void point(int, float, float) const {}
};
} // namespace

extern "C" {

// Packs the problem exactly as the product does (host pointers instead of device pointers) and
// marches rays [first, first+count) of the grid enumeration.  Outputs: gvl/evl/ivl
// [count][(N-1)*3] (zero outside the visited range), exit [count][6] = pos.x, pos.y, s.x, s.y,
// s.z, escaped; meta [count][2] = seg_lo, seg_hi.  Returns total march steps, or -1.
long long hostsim_march(const rtb200_problem *p, long long first, long long count, float *gvl,
                        float *evl, int *ivl, float *exit_state, int *meta, int flat)
{
    DevProblem P;
    const size_t bytes = pack_problem(*p, false, 0, 0.0, nullptr, nullptr, P);
    std::vector<char> blob(bytes + 256);
    char *base = (char *) (((uintptr_t) blob.data() + 255) & ~(uintptr_t) 255);
    pack_problem(*p, false, 0, 0.0, base, base, P);
    const int S = (P.N - 1) * RTB_N_SUB;
    long long steps_total = 0;
    const long long AB = (long long) P.sna * P.snb;
    for (long long r = 0; r < count; r++) {
        const long long ijkm = P.n_start + (first + r) * P.n_parallel;
        if (ijkm >= (long long) P.snx * P.sny * AB)
            return -1;
        const int m = (int) (ijkm % P.snb);
        const int k = (int) ((ijkm / P.snb) % P.sna);
        const int j = (int) ((ijkm / AB) % P.sny);
        const int i = (int) (ijkm / (AB * P.sny));
        std::memset(gvl + r * S, 0, sizeof(float) * S);
        std::memset(evl + r * S, 0, sizeof(float) * S);
        std::memset(ivl + r * S, 0, sizeof(int) * S);
        ArraySink sink{ gvl + r * S, evl + r * S, ivl + r * S };
        MarchResult res;
        unsigned steps = 0;
        if (flat) { // the flat state machine the GPU kernel runs
            float zt[RTB_N_SUB];
            for (int iz = 0; iz < RTB_N_SUB; iz++)
                zt[iz] = march_sub_limit(iz, P.dz0);
            MarchConsts K;
            march_consts(K, P.lite, zt, P.N, P.method, P.c, P.use_emis != 0);
            FlatMarch fm;
            flat_init(fm, K, P.sxf[i], P.syf[j], P.tanA[k], P.tanB[m]);
            while (flat_phase(fm) != PH_DONE && flat_iterate(fm, K, sink)) {
            }
            res.pos = fm.pos;
            res.s = fm.s;
            res.escaped = flat_escaped(fm) ? 1 : 0;
            flat_visited_range(fm, K, res.seg_lo, res.seg_hi);
            steps = fm.steps;
        } else {
            march_ray(P.planes, P.N, P.method, P.dz0, P.c, P.use_emis != 0, P.sxf[i], P.syf[j],
                      P.tanA[k], P.tanB[m], sink, res, steps);
        }
        steps_total += steps;
        float *e = exit_state + r * 6;
        e[0] = res.pos.x;
        e[1] = res.pos.y;
        e[2] = res.s.x;
        e[3] = res.s.y;
        e[4] = res.s.z;
        e[5] = (float) res.escaped;
        meta[2 * r] = res.seg_lo;
        meta[2 * r + 1] = res.seg_hi;
    }
    return steps_total;
}

// Table doors (host packing logic): owner tables and seed factors for the grid enumeration.
int hostsim_tables(const rtb200_problem *p, int *pixI, int *pixJ, int *binA, int *binB,
                   float *tanA, float *tanB, double *seed_f /* nx+ny+na+nb or null */)
{
    DevProblem P;
    const size_t bytes = pack_problem(*p, false, 0, 0.0, nullptr, nullptr, P);
    std::vector<char> blob(bytes + 256);
    char *base = (char *) (((uintptr_t) blob.data() + 255) & ~(uintptr_t) 255);
    pack_problem(*p, false, 0, 0.0, base, base, P);
    std::memcpy(pixI, P.pixI, sizeof(int) * P.snx);
    std::memcpy(pixJ, P.pixJ, sizeof(int) * P.sny);
    std::memcpy(binA, P.binA, sizeof(int) * P.sna);
    std::memcpy(binB, P.binB, sizeof(int) * P.snb);
    std::memcpy(tanA, P.tanA, sizeof(float) * P.sna);
    std::memcpy(tanB, P.tanB, sizeof(float) * P.snb);
    if (seed_f && P.seed_fx) {
        std::memcpy(seed_f, P.seed_fx, sizeof(double) * P.snx);
        std::memcpy(seed_f + P.snx, P.seed_fy, sizeof(double) * P.sny);
        std::memcpy(seed_f + P.snx + P.sny, P.seed_fa, sizeof(double) * P.sna);
        std::memcpy(seed_f + P.snx + P.sny + P.sna, P.seed_fb, sizeof(double) * P.snb);
    }
    return P.method;
}

double hostsim_pchip(size_t N, const double *xi, const double *yi, double x)
{
    return host_interp_pchip(N, xi, yi, x);
}
int hostsim_get_index(int n, const double *x, double dx, double y) { return host_get_index(n, x, dx, y); }
int hostsim_find_cell(const double *X, int n, double Y)
{
    double xl, xr;
    const double inv = n > 1 ? (double) (n - 1) / (X[n - 1] - X[0]) : 0.0;
    return find_cell(X, n, X[0], inv, Y, xl, xr);
}

int hostsim_find_cell_fast(const double *X, int n, double Yd)
{
    std::vector<AxisCell> t((size_t) n);
    fill_axis_cells(X, n, t.data());
    const double inv = n > 1 ? (double) (n - 1) / (X[n - 1] - X[0]) : 0.0;
    const float Yf = (float) Yd; // the march looks up float coordinates
    return find_cell_fast(t.data(), X, n, (float) X[0], (float) inv, X[0], inv, Yf, (double) Yf);
}

// FP64 update doors (rtb200_fp64.cuh): the fast exp and the two update branches.
static const double k_exp_table[RTB_EXP_TABLE_SIZE] = { RTB_EXP_TABLE_VALUES };
static const double k_fp[RTB_K_COUNT] = { RTB_K_VALUES };
static const ArrayConsts k_consts = { k_fp, k_exp_table };
void hostsim_exp(const double *x, double *y, int n)
{
    for (int i = 0; i < n; i++)
        y[i] = exp_any(x[i], k_consts);
}
// One update with float inputs gvl, evl, g as the kernel forms them; branch chosen like the kernel.
double hostsim_ase_update(double Iv, float gvl, float evl, float g)
{
    const float glf = gvl * g, elf = evl * g;
    const double gl = (double) glf, el = (double) elf;
    if (!(fabsf(glf) < 700.0f))
        return -1.0; // library path on the device
    if (fabsf(glf) < 1e-3f)
        return ase_update_small(Iv, gl, el, k_consts);
    return ase_update_large(Iv, gl, el, 1.0f / glf, k_consts);
}

// Exactness doors for the division shortcuts of rtb200_math.cuh.
// Random operand pairs for ddiv_by; returns the number of mismatches against `/`.
long long hostsim_check_ddiv_by(long long n, unsigned long long seed)
{
    unsigned long long st = seed * 2862933555777941757ULL + 3037000493ULL;
    auto next = [&]() {
        st ^= st << 13;
        st ^= st >> 7;
        st ^= st << 17;
        return st;
    };
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        // a: signed, magnitude 2^[-40, 10); b: positive, magnitude 2^[-30, 0); random mantissas
        const unsigned long long ma = next(), mb = next();
        const int ea = (int) (next() % 50) - 40, eb = (int) (next() % 30) - 30;
        double a = ldexp(1.0 + (double) (ma >> 12) / 4503599627370496.0, ea);
        if (ma & 1)
            a = -a;
        const double b = ldexp(1.0 + (double) (mb >> 12) / 4503599627370496.0, eb);
        const double rb = 1.0 / b;
        if (ddiv_by(a, b, rb) != a / b)
            bad++;
    }
    return bad;
}
// markstein_safe (rtb200_pack.h): the per-divisor proof obligation of ddiv_by.
int hostsim_markstein_safe(double b, double *witness, double rb_test)
{
    return markstein_safe(b, witness, rb_test) ? 1 : 0;
}
double hostsim_ddiv_by(double a, double b) { return ddiv_by(a, b, 1.0 / b); }
double hostsim_ddiv_by_rb(double a, double b, double rb) { return ddiv_by(a, b, rb); }
// Searches random divisors in [2^-20, 2^-19) for one that markstein_safe rejects; returns the
// number of rejected divisors among `n` and the last one with its witness numerator.
long long hostsim_find_unsafe_divisor(long long n, unsigned long long seed, double *b_out, double *a_out)
{
    unsigned long long st = seed * 2862933555777941757ULL + 3037000493ULL;
    long long found = 0;
    for (long long i = 0; i < n; i++) {
        st ^= st << 13;
        st ^= st >> 7;
        st ^= st << 17;
        const double b = ldexp(1.0 + (double) (st >> 12) / 4503599627370496.0, -20);
        double w = 0.0;
        if (!markstein_safe(b, &w)) {
            found++;
            *b_out = b;
            *a_out = w;
        }
    }
    return found;
}
// Every float in [lo_bits, hi_bits] (bit patterns, both signs) for fdiv_const with c.
long long hostsim_check_fdiv_const(float c, unsigned lo_bits, unsigned hi_bits)
{
    const float rc = 1.0f / c;
    long long bad = 0;
    for (unsigned long long b = lo_bits; b <= hi_bits; b++) {
        for (int sgn = 0; sgn < 2; sgn++) {
            const unsigned bits = (unsigned) b | (sgn ? 0x80000000u : 0u);
            float x;
            memcpy(&x, &bits, 4);
            const float q = fdiv_const(x, c, rc), want = x / c;
            unsigned qb, wb;
            memcpy(&qb, &q, 4);
            memcpy(&wb, &want, 4);
            if (qb != wb && !(q != q && want != want))
                bad++;
        }
    }
    return bad;
}
// 1.0 / sqrt(float) as a double division rounded to float (the reference's normalize_s) equals
// the float division 1.0f / sqrtf(x): checked over bit patterns [lo_bits, hi_bits].
long long hostsim_check_rsqrt_identity(unsigned lo_bits, unsigned hi_bits)
{
    long long bad = 0;
    for (unsigned long long b = lo_bits; b <= hi_bits; b++) {
        float x;
        const unsigned bits = (unsigned) b;
        memcpy(&x, &bits, 4);
        const float r = sqrtf(x);
        const float viad = (float) (1.0 / (double) r), viaf = 1.0f / r;
        if (viad != viaf && !(viad != viad && viaf != viaf))
            bad++;
    }
    return bad;
}

} // extern "C"

// Timing door (tools): seconds per packing pass of the main blob, lineshape tables deferred as
// rtb200_create_image does (they are packed while the march runs).
#include <chrono>
extern "C" double hostsim_pack_seconds(const rtb200_problem *p, int reps)
{
    DevProblem P;
    GvBlob gm{ nullptr, nullptr, false, 0 };
    const size_t bytes = pack_problem(*p, false, 0, 0.0, nullptr, nullptr, P, &gm);
    std::vector<char> blob(bytes + 256), gvh(gm.bytes + 256);
    char *base = (char *) (((uintptr_t) blob.data() + 255) & ~(uintptr_t) 255);
    char *gbase = (char *) (((uintptr_t) gvh.data() + 255) & ~(uintptr_t) 255);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) {
        GvBlob g{ gbase, gbase, false, 0 };
        pack_problem(*p, false, 0, 0.0, base, base, P, &g);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}
