// rtb200_pack.h — host-side re-layout of a create_image problem into ONE packed
// structure-of-arrays blob (pure host C++, no CUDA calls).
//
// Replaces the reference's deep copy of an array-of-structs-of-pointers (7 cudaMalloc + 7
// cudaMemcpy per gain plane, 10 + 10 for the seed: src/RayTraceImageCuda.cu:225-329) by one
// contiguous buffer that is filled in pinned memory and uploaded with a single H2D copy.
// Device pointers inside the blob are computed from the blob's device base address.
//
// Also tabulates everything that depends on a grid INDEX only, so that no libm call whose
// result could differ between host and device libms is ever evaluated on the device:
//   * tanf(1e-3f*a), tanf(1e-3f*b) for the start direction (RayTraceImageHelper.h:409-410);
//   * the owner cell of each source coordinate for method 1 (getIndex, RayTraceImageCPU.cpp:11-16);
//   * the separable seed factors for method 2 (calc_seed_inline / interp_pchip, :168-247).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/rtb200.h"
#include "rtb200_device.cuh"

namespace rtb {

// The large arrays of a problem (n, g0, E0, gv) may alias an unaligned byte stream
// (rtb200_create_image_from_dat): every read of them goes through memcpy.
template <class T>
inline T ld_u(const T *p)
{
    T v;
    std::memcpy(&v, p, sizeof(T));
    return v;
}

// Splits [0, n) into at most 8 contiguous pieces and runs fn(piece, begin, end) on host threads
// when the job is large enough to pay for them (re-layout of multi-megabyte gain planes); small
// problems - every shipped input - stay on the calling thread (one piece).
enum { RTB_MAX_PIECES = 8 };
template <class F>
inline void parallel_pieces(size_t n, size_t bytes, F fn)
{
    unsigned t = 1;
    if (bytes >= ((size_t) 8 << 20)) {
        const unsigned hw = std::thread::hardware_concurrency();
        t = hw > RTB_MAX_PIECES ? (unsigned) RTB_MAX_PIECES : (hw < 1 ? 1 : hw);
        if ((size_t) t > n)
            t = (unsigned) (n ? n : 1);
    }
    if (t <= 1) {
        fn(0u, (size_t) 0, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (n + t - 1) / t;
    for (unsigned i = 0; i < t; i++) {
        const size_t a = std::min(n, i * per), b = std::min(n, a + per);
        if (a < b)
            th.emplace_back([=]() { fn(i, a, b); });
    }
    for (auto &x : th)
        x.join();
}

// findfirstsingle (RayTraceImageHelper.h:101-117)
inline size_t host_findfirstsingle(const double *X, size_t n, double Y)
{
    if (Y < X[0])
        return 0;
    if (Y > X[n - 1])
        return n;
    size_t lo = 0, hi = n - 1;
    while (hi - lo != 1) {
        size_t mid = (hi + lo) / 2;
        if (X[mid] >= Y)
            hi = mid;
        else
            lo = mid;
    }
    return hi;
}

// getIndex (RayTraceImageCPU.cpp:11-16)
inline int host_get_index(int n, const double *x, double dx, double y)
{
    if (y < x[0] - 0.5 * dx || y > x[n - 1] + 0.5 * dx)
        return -1;
    return (int) host_findfirstsingle(x, (size_t) n, y - 0.5 * dx);
}

// interp_pchip (RayTraceImageHelper.h:168-220): monotone cubic Hermite interpolation.
inline double host_interp_pchip(size_t N, const double *xi, const double *yi, double x)
{
    if (x <= xi[0] || N <= 2) {
        double t = (x - xi[0]) / (xi[1] - xi[0]);
        return (1.0 - t) * yi[0] + t * yi[1];
    }
    if (x >= xi[N - 1]) {
        double t = (x - xi[N - 2]) / (xi[N - 1] - xi[N - 2]);
        return (1.0 - t) * yi[N - 2] + t * yi[N - 1];
    }
    const size_t i = host_findfirstsingle(xi, N, x);
    const double f1 = yi[i - 1], f2 = yi[i];
    const double t = (x - xi[i - 1]) / (xi[i] - xi[i - 1]);
    double g1 = 0, g2 = 0;
    if (i <= 1) {
        g1 = f2 - f1;
    } else if ((f1 < f2 && f1 > yi[i - 2]) || (f1 > f2 && f1 < yi[i - 2])) {
        const double f0 = yi[i - 2];
        const double h1 = xi[i - 1] - xi[i - 2], h2 = xi[i] - xi[i - 1];
        const double a1 = (h2 - h1) / h1, a2 = h1 / (h1 + h2);
        g1 = a1 * (f1 - f0) + a2 * (f2 - f0);
        const double s1 = std::fabs(f1 - f0) / h1, s2 = std::fabs(f2 - f1) / h2;
        const double g_max = 2 * h2 * (s1 < s2 ? s1 : s2);
        g1 = ((g1 >= 0) ? 1 : -1) * (std::fabs(g1) < g_max ? std::fabs(g1) : g_max);
    }
    if (i >= N - 1) {
        g2 = f2 - f1;
    } else if ((f2 < f1 && f2 > yi[i + 1]) || (f2 > f1 && f2 < yi[i + 1])) {
        const double f0 = yi[i + 1];
        const double h1 = xi[i] - xi[i - 1], h2 = xi[i + 1] - xi[i];
        const double a1 = -h2 / (h1 + h2), a2 = (h2 - h1) / h2;
        g2 = a1 * (f1 - f0) + a2 * (f2 - f0);
        const double s1 = std::fabs(f2 - f1) / h1, s2 = std::fabs(f0 - f2) / h2;
        const double g_max = 2 * h1 * (s1 < s2 ? s1 : s2);
        g2 = ((g2 >= 0) ? 1 : -1) * (std::fabs(g2) < g_max ? std::fabs(g2) : g_max);
    }
    const double t2 = t * t;
    return f1 + t2 * (2 * t - 3) * (f1 - f2) + t * g1 - t2 * (g1 + (1 - t) * (g1 + g2));
}

// Is ddiv_by(a, b, RN(1/b)) == RN(a / b) for EVERY a?  (rtb200_math.cuh)
//
// With y = RN(1/b), q0 = RN(a*y), r = a - b*q0 and q1 = RN(q0 + r*y), the value rounded last is
// a/b + (a/b - q0)*(b*y - 1): it misses a/b by less than 2^-104 relative (|a/b - q0| <= 2 ulp,
// |b*y - 1| < 2^-53, plus the rounding of r should it be inexact), so q1 can only differ from
// RN(a/b) when a/b lies that close to a midpoint of two doubles.  Scale a and b to integer
// significands A, B in [2^52, 2^53): a midpoint is M*2^-53 (A >= B) or M*2^-54 (A < B) with M odd
// in (2^53, 2^54), and |A/B - midpoint| = |D| / (B*2^s), D = A*2^s - B*M a non-zero integer, s = 53
// or 54.  Closeness therefore needs |D| < 8 (|D| <= 16 is searched), and for a given B every such A follows from the
// congruence B*M = -D (mod 2^s).  The handful of candidates is enumerated and the algorithm is
// run on each; a divisor that fails once is not used with ddiv_by (the plane falls back to IEEE
// divisions).  `witness` (optional) receives a failing numerator.  Floating point is scale
// invariant, so the verdict holds for every binade of a and b short of over/underflow.
// `rb_test` (tests only) replaces RN(1/B) for the integer-scaled divisor B: a reciprocal that is
// one ulp off must be rejected, which is how tests/test_math_identities.py checks the search.
inline bool markstein_safe(double b, double *witness = nullptr, double rb_test = 0.0)
{
    if (!(b > 0.0) || !std::isnormal(b))
        return false;
    int eb;
    const double mb = std::frexp(b, &eb); // [0.5, 1)
    const uint64_t B = (uint64_t) std::ldexp(mb, 53);
    const double bd = (double) B, rb = rb_test != 0.0 ? rb_test : 1.0 / bd;
    int t = 0;
    while (((B >> t) & 1u) == 0u)
        t++;
    const uint64_t Bodd = B >> t;
    uint64_t inv = Bodd; // inverse of Bodd modulo 2^64 (Newton; 3 correct bits to start with)
    for (int it = 0; it < 6; it++)
        inv *= 2u - Bodd * inv;
    const uint64_t two52 = 1ull << 52, two53 = 1ull << 53, two54 = 1ull << 54;
    for (int s = 53; s <= 54; s++) {
        for (int D = -16; D <= 16; D++) {
            if (D == 0 || t > 4 || (D % (1 << t)) != 0)
                continue; // B*M + D must be divisible by 2^s, hence D by 2^t
            const int bits = s - t; // modulus of the congruence for M
            const uint64_t mask = (1ull << bits) - 1u;
            const uint64_t rhs = (uint64_t) (int64_t) (-D / (1 << t));
            for (uint64_t M = (rhs * inv) & mask; M < two54; M += mask + 1u) {
                if (M <= two53 || (M & 1u) == 0u)
                    continue;
                const unsigned __int128 num = (unsigned __int128) B * M + (__int128) D;
                const uint64_t A = (uint64_t) (num >> s);
                if (A < two52 || A >= two53 || ((unsigned __int128) A << s) != num)
                    continue;
                const double a = (double) A;
                if (ddiv_by(a, bd, rb) != a / bd) {
                    if (witness)
                        *witness = rb_test != 0.0 ? a : std::ldexp(a, eb - 53); // same significands
                    return false;
                }
            }
        }
    }
    return true;
}

// Interval table of one axis (AxisCell, rtb200_march.cuh): entry k describes [c[k-1], c[k]].
inline void fill_axis_cells(const double *c, int n, AxisCell *t)
{
    std::memset(&t[0], 0, sizeof(AxisCell));
    for (int k = 1; k < n; k++) {
        AxisCell a;
        a.lo = c[k - 1];
        a.hi = c[k];
        a.w = c[k] - c[k - 1];
        a.rw = 1.0 / a.w;
        a.d = (float) a.w;
        a.dd = (double) a.d;
        a.rd = 1.0 / a.dd;
        a.dm = 0.1f * a.d;
        const double h = 0.1 * a.w;
        a.halo_lo = (float) (a.lo - h);
        a.halo_hi = (float) (a.hi + h);
        t[k] = a;
    }
}

// Cell records of one plane (CellRec, rtb200_march.cuh): record i1 = (k1-1) + (k2-1)*Nx holds
// the corners i1, i1+1, i1+Nx, i1+Nx+1 in the reference's order (:474-477).  The last row and the
// last column of the array are never addressed (cells end at Nx-2, Ny-2).
inline void fill_cell_records(const rtb200_gain_plane &g, const AxisCell *cx, const AxisCell *cy,
                              CellRec *t)
{
    const int Nx = g.Nx, Ny = g.Ny;
    parallel_pieces((size_t) Ny, sizeof(CellRec) * (size_t) Nx * Ny, [=](unsigned, size_t j0, size_t j1) {
    for (int j = (int) j0; j < (int) j1; j++)
        for (int i = 0; i < Nx; i++) {
            if (i + 1 >= Nx || j + 1 >= Ny) { // last row / column: never addressed
                std::memset(&t[(size_t) i + (size_t) j * Nx], 0, sizeof(CellRec));
                continue;
            }
            const size_t i1 = (size_t) i + (size_t) j * Nx;
            const size_t c[4] = { i1, i1 + 1, i1 + (size_t) Nx, i1 + (size_t) Nx + 1 };
            CellRec r;
            double nc[4];
            for (int q = 0; q < 4; q++) {
                nc[q] = ld_u(&g.n[c[q]]);
                r.nf[q] = (float) nc[q];
                r.g0[q] = ld_u(&g.g0[c[q]]);
                r.E0[q] = g.E0 ? ld_u(&g.E0[c[q]]) : 0.0f;
            }
            r.n10 = nc[1] - nc[0];
            r.n32 = nc[3] - nc[2];
            r.n20 = nc[2] - nc[0];
            r.n31 = nc[3] - nc[1];
            r.xl = cx[i + 1].lo;
            r.dxd = cx[i + 1].dd;
            r.rdx = cx[i + 1].rd;
            r.yl = cy[j + 1].lo;
            r.dyd = cy[j + 1].dd;
            r.rdy = cy[j + 1].rd;
            t[i1] = r;
        }
    });
}

// max over |v| as float bit patterns (sign cleared; NaN > inf > every finite value)
inline unsigned abs_max_bits(const float *v, size_t n, unsigned m = 0u)
{
    const unsigned char *b = reinterpret_cast<const unsigned char *>(v);
    for (size_t i = 0; i < n; i++) {
        uint32_t a;
        std::memcpy(&a, b + 4 * i, 4);
        a &= 0x7fffffffu;
        m = a > m ? a : m;
    }
    return m;
}

// Copies n floats and returns max |v| as float bits (sign cleared; NaN > inf > every finite
// value), on host threads for large tables.
inline unsigned copy_abs_max(float *dst, const float *src, size_t n, unsigned m0)
{
    unsigned part[RTB_MAX_PIECES] = { 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u };
    parallel_pieces(n, n * sizeof(float), [&part, dst, src](unsigned piece, size_t a, size_t b) {
        std::memcpy(dst + a, src + a, (b - a) * sizeof(float));
        part[piece] = abs_max_bits(src + a, b - a);
    });
    unsigned m = m0;
    for (unsigned v : part)
        m = v > m ? v : m;
    return m;
}

// float -> double, on host threads for large tables
inline void widen_floats(double *dst, const float *src, size_t n)
{
    parallel_pieces(n, n * sizeof(double), [dst, src](unsigned, size_t a, size_t b) {
        for (size_t i = a; i < b; i++)
            dst[i] = (double) ld_u(&src[i]);
    });
}

// RayTraceImageHelper.h:402: emission + gain (ASE) unless a seed is given
inline bool problem_use_emis(const rtb200_problem &p) { return p.gain[0].E0 != nullptr && p.seed == nullptr; }

// Bump allocator over the staging blob.  With host == nullptr it only measures.
class Blob {
public:
    Blob(char *host, const char *dev_base) : host_(host), dev_(dev_base), off_(0) {}
    template <class T>
    T *alloc(size_t n, const T **dev_ptr)
    {
        off_ = (off_ + 255) & ~size_t(255);
        T *h = host_ ? reinterpret_cast<T *>(host_ + off_) : nullptr;
        *dev_ptr = reinterpret_cast<const T *>(dev_ + off_);
        off_ += n * sizeof(T);
        return h;
    }
    size_t size() const { return (off_ + 255) & ~size_t(255); }
    bool filling() const { return host_ != nullptr; }

private:
    char *host_;
    const char *dev_;
    size_t off_;
};

// The lineshape tables gv[cell][k] are nine tenths of the blob and only the integration reads
// them, so they can live in a second blob that is filled and uploaded while the march is already
// running (rtb200_host.cu).  `bytes` is set by every pass; with `copy` false the arrays are only
// laid out and pack_gv() fills them later.
struct GvBlob {
    char *host;
    const char *dev;
    bool copy;
    size_t bytes;
};

// The per-cell records (CellRec) are derived data, a third of the main blob: the product lays them
// out in a device buffer of their own and lets a small kernel derive them from the uploaded nodes
// and interval tables (build_cell_records_kernel: the same IEEE operations as fill_cell_records),
// so the host neither computes nor uploads them.  Without a CellBlob (host builds of the march:
// tests/hostsim) they stay in the main blob and are filled here.
struct CellBlob {
    const char *dev;
    size_t bytes;
};

// Packs `p` into the blob.  Returns the number of bytes used; fills `out` (pointers relative
// to dev_base).  explicit_rays: the ray list comes separately (rtb200_trace_rays), so only the
// planes, the destination grid and dv are packed and method/scale are given by the caller.
// gvb == nullptr keeps the lineshape tables inside the main blob.
inline size_t pack_problem(const rtb200_problem &p, bool explicit_rays, int method_in,
                           double scale_in, char *host, const char *dev_base, DevProblem &out,
                           GvBlob *gvb = nullptr, CellBlob *cellb = nullptr, CellBlob *gvdb = nullptr)
{
    // gvdb: like cellb for the double copy of the lineshape tables of gain-only problems
    // (DevPlane::gvd): laid out in a device buffer of its own and widened there (widen_gv_kernel)
    Blob blob(host, dev_base);
    Blob gv_blob(gvb ? gvb->host : nullptr, gvb ? gvb->dev : nullptr);
    Blob cell_blob(nullptr, cellb ? cellb->dev : nullptr);
    Blob gvd_blob(nullptr, gvdb ? gvdb->dev : nullptr);
    const bool fill = blob.filling();
    const rtb200_beam &e = *p.euv_beam;
    const int N = p.N, K = e.nv;
    std::memset(&out, 0, sizeof(out));
    out.N = N;
    out.K = K;
    out.dz0 = (float) e.dz;
    out.c = 0.5f;
    out.use_emis = problem_use_emis(p) ? 1 : 0; // :402
    const double kfp[RTB_K_COUNT] = { RTB_K_VALUES };
    std::memcpy(out.kfp, kfp, sizeof(kfp));

    double *kfp_g = blob.alloc<double>(RTB_K_COUNT, &out.kfp_g);
    if (fill)
        std::memcpy(kfp_g, kfp, sizeof(kfp));

    // ---- gain planes ------------------------------------------------------------------------
    unsigned gv_absmax = 0u;
    DevPlane *planes = blob.alloc<DevPlane>((size_t) N, &out.planes);
    PlaneLite *lite = blob.alloc<PlaneLite>((size_t) N, &out.lite);
    for (int ii = 0; ii < N; ii++) {
        const rtb200_gain_plane &g = p.gain[ii];
        const size_t nn = (size_t) g.Nx * g.Ny;
        DevPlane P;
        std::memset(&P, 0, sizeof(P));
        double *x = blob.alloc<double>((size_t) g.Nx, &P.x);
        double *y = blob.alloc<double>((size_t) g.Ny, &P.y);
        Node *node = blob.alloc<Node>(nn, &P.node);
        float *gv = (gvb ? gv_blob : blob).alloc<float>(nn * (size_t) K, &P.gv);
        // gain-only problems: the table once more in double (DevPlane::gvd)
        double *gvd = nullptr;
        if (!out.use_emis) {
            if (gvdb) {
                gvd_blob.alloc<double>(nn * (size_t) K, &P.gvd);
                gvdb->bytes = gvd_blob.size();
            } else {
                gvd = blob.alloc<double>(nn * (size_t) K, &P.gvd);
            }
        }
        if (gvb)
            gvb->bytes = gv_blob.size();
        AxisCell *cx = blob.alloc<AxisCell>((size_t) g.Nx, &P.cx);
        AxisCell *cy = blob.alloc<AxisCell>((size_t) g.Ny, &P.cy);
        CellRec *cell = cellb ? cell_blob.alloc<CellRec>(nn, &P.cell) : blob.alloc<CellRec>(nn, &P.cell);
        if (cellb)
            cellb->bytes = cell_blob.size();
        if (fill) {
            bool ok = true;
            double last_w = 0.0;
            bool last_ok = false;
            auto recip = [&](const double *c, int n) {
                for (int k = 1; k < n; k++) {
                    const double w = c[k] - c[k - 1];
                    const double rw = 1.0 / w, rd = 1.0 / (double) (float) w;
                    ok = ok && std::isnormal(rw) && std::isnormal(rd) && std::isnormal(w);
                    // every numerator must divide exactly through the tabulated reciprocal
                    // (neighbouring cells of a grid mostly share their width: test it once)
                    if (w != last_w) {
                        last_ok = markstein_safe(w) && markstein_safe((double) (float) w);
                        last_w = w;
                    }
                    ok = ok && last_ok;
                }
            };
            recip(g.x, g.Nx);
            recip(g.y, g.Ny);
            fill_axis_cells(g.x, g.Nx, cx);
            fill_axis_cells(g.y, g.Ny, cy);
            if (cell) // (null: derived on the device)
                fill_cell_records(g, cx, cy, cell);
            P.fast_div = ok ? 1 : 0;
            std::memcpy(x, g.x, sizeof(double) * g.Nx);
            std::memcpy(y, g.y, sizeof(double) * g.Ny);
            parallel_pieces(nn, nn * sizeof(Node), [&](unsigned, size_t q0, size_t q1) {
                for (size_t q = q0; q < q1; q++) {
                    node[q].n = ld_u(&g.n[q]);
                    node[q].g0 = ld_u(&g.g0[q]);
                    node[q].E0 = g.E0 ? ld_u(&g.E0[q]) : 0.0f;
                }
            });
            if (gvd) // (null: widened on the device)
                widen_floats(gvd, g.gv, nn * (size_t) K);
            if (gv && (!gvb || gvb->copy)) {
                gv_absmax = copy_abs_max(gv, g.gv, nn * (size_t) K, gv_absmax);
            } else {
                gv_absmax = 0x7fffffffu; // tables filled later: pack_gv() returns the value
            }
            P.Nx = g.Nx;
            P.Ny = g.Ny;
            P.range[0] = (float) g.x[0]; // :445-453
            P.range[1] = (float) g.x[g.Nx - 1];
            P.range[2] = (float) g.y[0];
            P.range[3] = (float) g.y[g.Ny - 1];
            P.abs_y = 0;
            if (P.range[2] >= 0) {
                P.range[2] = -P.range[3];
                P.abs_y = 1;
            }
            P.x0 = g.x[0];
            P.y0 = g.y[0];
            P.inv_dx = g.Nx > 1 ? (double) (g.Nx - 1) / (g.x[g.Nx - 1] - g.x[0]) : 0.0;
            P.inv_dy = g.Ny > 1 ? (double) (g.Ny - 1) / (g.y[g.Ny - 1] - g.y[0]) : 0.0;
            P.x0f = (float) P.x0;
            P.y0f = (float) P.y0;
            P.inv_dxf = (float) P.inv_dx;
            P.inv_dyf = (float) P.inv_dy;
            planes[ii] = P;
            PlaneLite L;
            std::memset(&L, 0, sizeof(L));
            L.cx = P.cx;
            L.cy = P.cy;
            L.cell = P.cell;
            L.full = out.planes + ii;
            L.x0f = P.x0f;
            L.inv_dxf = P.inv_dxf;
            L.y0f = P.y0f;
            L.inv_dyf = P.inv_dyf;
            L.r0 = P.range[0];
            L.r1 = P.range[1];
            L.r2 = P.range[2];
            L.r3 = P.range[3];
            L.Nx = P.Nx;
            L.Ny = P.Ny;
            L.flags = (P.abs_y ? 1 : 0) | (P.fast_div ? 2 : 0);
            lite[ii] = L;
        }
    }

    out.gv_absmax_bits = fill ? gv_absmax : 0x7fffffffu;

    // ---- destination grid (euv_beam) --------------------------------------------------------
    out.nx = e.nx;
    out.ny = e.ny;
    out.na = e.na;
    out.nb = e.nb;
    out.edx = e.dx;
    out.edy = e.dy;
    out.eda = e.da;
    out.edb = e.db;
    out.y_mirror = (e.ny > 0 && e.y && e.y[0] >= 0.0) ? 1 : 0;
    double *ex = blob.alloc<double>((size_t) e.nx, &out.ex);
    double *ey = blob.alloc<double>((size_t) e.ny, &out.ey);
    double *ea = blob.alloc<double>((size_t) e.na, &out.ea);
    double *eb = blob.alloc<double>((size_t) e.nb, &out.eb);
    double *dv2 = blob.alloc<double>((size_t) K, &out.dv2);
    if (fill) {
        if (e.nx > 0)
            std::memcpy(ex, e.x, sizeof(double) * e.nx);
        if (e.ny > 0)
            std::memcpy(ey, e.y, sizeof(double) * e.ny);
        if (e.na > 0)
            std::memcpy(ea, e.a, sizeof(double) * e.na);
        if (e.nb > 0)
            std::memcpy(eb, e.b, sizeof(double) * e.nb);
        for (int k = 0; k < K; k++)
            dv2[k] = e.dv ? 2.0 * e.dv[k] : 0.0; // RayTraceImageCPU.cpp:66: (2.0*dv[iv])*Iv[iv]
    }
    if (p.seed) {
        double *fv = blob.alloc<double>((size_t) p.seed->dim[4], &out.seed_fv);
        if (fill)
            std::memcpy(fv, p.seed->f[4], sizeof(double) * p.seed->dim[4]);
        out.seed_f0 = p.seed->f0;
        for (int d = 0; d < 4; d++) {
            out.sd_dim[d] = p.seed->dim[d];
            double *x = blob.alloc<double>((size_t) p.seed->dim[d], &out.sd_x[d]);
            double *f = blob.alloc<double>((size_t) p.seed->dim[d], &out.sd_f[d]);
            if (fill) {
                std::memcpy(x, p.seed->x[d], sizeof(double) * p.seed->dim[d]);
                std::memcpy(f, p.seed->f[d], sizeof(double) * p.seed->dim[d]);
            }
        }
    }
    if (explicit_rays) {
        out.method = method_in;
        out.scale = scale_in;
        return blob.size();
    }

    // ---- ray source grid (src/RayTraceImage.cpp:277-328) ------------------------------------
    const rtb200_beam &sz = p.seed ? *p.seed_beam : e;                 // sizes follow `seed`
    const rtb200_beam &sc = p.seed_beam ? *p.seed_beam : *p.euv_beam; // coordinates follow `seed_beam`
    if (p.seed) {
        out.method = 2;
        out.scale = (sz.dx * sz.dy * sz.da * sz.db) / (e.dx * e.dy);
    } else {
        out.method = 1;
        out.scale = 1.0;
    }
    out.snx = sz.nx;
    out.sny = sz.ny;
    out.sna = sz.na;
    out.snb = sz.nb;
    out.n_start = p.N_start;
    out.n_parallel = p.N_parallel;
    const long long AB = (long long) sz.na * sz.nb;
    out.ab_max = (int) ((AB + p.N_parallel - 1) / p.N_parallel);
    float *sxf = blob.alloc<float>((size_t) sz.nx, &out.sxf);
    float *syf = blob.alloc<float>((size_t) sz.ny, &out.syf);
    float *saf = blob.alloc<float>((size_t) sz.na, &out.saf);
    float *sbf = blob.alloc<float>((size_t) sz.nb, &out.sbf);
    float *tanA = blob.alloc<float>((size_t) sz.na, &out.tanA);
    float *tanB = blob.alloc<float>((size_t) sz.nb, &out.tanB);
    int *pixI = blob.alloc<int>((size_t) sz.nx, &out.pixI);
    int *pixJ = blob.alloc<int>((size_t) sz.ny, &out.pixJ);
    int *binA = blob.alloc<int>((size_t) sz.na, &out.binA);
    int *binB = blob.alloc<int>((size_t) sz.nb, &out.binB);
    if (fill) {
        for (int i = 0; i < sz.nx; i++) {
            sxf[i] = (float) sc.x[i];
            pixI[i] = host_get_index(e.nx, e.x, e.dx, (double) sxf[i]);
        }
        for (int i = 0; i < sz.ny; i++) {
            syf[i] = (float) sc.y[i];
            pixJ[i] = host_get_index(e.ny, e.y, e.dy, (double) syf[i]);
        }
        for (int i = 0; i < sz.na; i++) {
            saf[i] = (float) sc.a[i];
            tanA[i] = tanf(1e-3f * saf[i]);
            binA[i] = host_get_index(e.na, e.a, e.da, (double) saf[i]);
        }
        for (int i = 0; i < sz.nb; i++) {
            sbf[i] = (float) sc.b[i];
            tanB[i] = tanf(1e-3f * sbf[i]);
            binB[i] = host_get_index(e.nb, e.b, e.db, (double) sbf[i]);
        }
    }
    if (p.seed) {
        // method 2: calc_seed_inline(seed, ray.x, ray.y, ray.a, ray.b) depends on the source
        // indices only.  NaN marks "outside the seed grid" (whole product becomes f = 0).
        const rtb200_seed &s = *p.seed;
        const double nan = std::numeric_limits<double>::quiet_NaN();
        double *f[4];
        f[0] = blob.alloc<double>((size_t) sz.nx, &out.seed_fx);
        f[1] = blob.alloc<double>((size_t) sz.ny, &out.seed_fy);
        f[2] = blob.alloc<double>((size_t) sz.na, &out.seed_fa);
        f[3] = blob.alloc<double>((size_t) sz.nb, &out.seed_fb);
        if (fill) {
            const int n[4] = { sz.nx, sz.ny, sz.na, sz.nb };
            const float *src[4] = { sxf, syf, saf, sbf };
            for (int d = 0; d < 4; d++)
                for (int i = 0; i < n[d]; i++) {
                    const double v = (double) src[d][i];
                    const bool inside = v >= s.x[d][0] && v <= s.x[d][s.dim[d] - 1];
                    f[d][i] = inside ? host_interp_pchip((size_t) s.dim[d], s.x[d], s.f[d], v) : nan;
                }
        }
    }
    return blob.size();
}

// Fills a GvBlob that pack_problem laid out with copy == false (same allocation sequence).
// Returns DevProblem::gv_absmax_bits.
inline unsigned pack_gv(const rtb200_problem &p, char *host)
{
    Blob gv_blob(host, host);
    const int K = p.euv_beam->nv;
    unsigned m = 0u;
    for (int ii = 0; ii < p.N; ii++) {
        const rtb200_gain_plane &g = p.gain[ii];
        const size_t n = (size_t) g.Nx * g.Ny * (size_t) K;
        const float *unused;
        float *gv = gv_blob.alloc<float>(n, &unused);
        m = copy_abs_max(gv, g.gv, n, m);
    }
    return m;
}

} // namespace rtb
