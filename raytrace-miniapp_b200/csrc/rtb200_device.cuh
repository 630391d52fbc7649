// rtb200_device.cuh — device-resident problem description shared by the host packer and the
// kernels.  Everything the kernels read lives in ONE packed blob (rtb200_pack.h) uploaded with a
// single H2D copy; the structs below hold pointers into it.
#pragma once
#include "rtb200_fp64.cuh"
#include "rtb200_march.cuh"

namespace rtb {

// Hand-off record between the march and the frequency integration: one per
// (ray, length segment, sub-segment); the reference's gvl/evl/ivl[ii-1][is]
// (src/common/RayTraceImageHelper.h:386-388, :501-503).
struct __attribute__((aligned(16))) SegRec {
    float gvl;
    float evl;
    int cell;
    int pad;
};

// Per-ray status word written by the march.
//   bits 0..11  seg_lo      (first visited record)
//   bits 12..23 seg_hi      (one past the last visited record)
//   bit 24      escaped     (left the plasma column)
//   bit 25      invalid     (s.z^2 < 0.01 at exit: error -1, :515-516)
//   bit 26      inactive    (no ray in this slot)
#define RTB_META_ESCAPED (1u << 24)
#define RTB_META_INVALID (1u << 25)
#define RTB_META_INACTIVE (1u << 26)
#define RTB_MAX_SEGS 4095

struct FailState {
    unsigned failure_code; // bit n set <=> some ray returned error -n (set_bit, :47-51)
    unsigned n_failed;
    float failed[32 * 4]; // first N_FAILED_MAX failed rays (x, y, a, b)
    unsigned long long march_steps;
};

struct DevProblem {
    const DevPlane *planes; // [N]
    const PlaneLite *lite;  // [N] what the march's cell look-up starts from (staged in shared memory)
    int N, K, method, use_emis;
    float dz0, c;
    double scale;
    // --- ray source grid (euv_beam for ASE, seed_beam for seeded), src/RayTraceImage.cpp:300-328
    int snx, sny, sna, snb;
    long long n_start, n_parallel; // ijkm = n_start + it*n_parallel
    int ab_max;                    // max rays of one source pixel in this worker = ceil(sna*snb/n_parallel)
    const float *sxf, *syf, *saf, *sbf; // source coordinates rounded to float (the ray_struct fields)
    const float *tanA, *tanB;           // tanf(1e-3f*a), tanf(1e-3f*b) from the host libm
    // --- destination (euv_beam) grid
    int nx, ny, na, nb;
    const double *ex, *ey, *ea, *eb; // cell centres, for getIndex on the exit ray (seeded)
    double edx, edy, eda, edb;
    int y_mirror; // beam.y[0] >= 0 (src/RayTraceImageCPU.cpp:45)
    // owner tables for method 1 (ray2 = ray): getIndex of the source coordinate, -1 = outside
    const int *pixI, *pixJ, *binA, *binB;
    const double *dv2; // 2.0*dv[k]
    // largest |gv| of all planes as float bits (NaN sorts above everything); 0x7fffffff = not
    // known.  Lets the ASE integration skip its per-record exp-range test for a whole ray when
    // max|gvl| * max|gv| < 700 (rtb200_kernels.cu, integrate_ray_ase_fast).
    unsigned gv_absmax_bits;
    // --- separable seed, tabulated per source index (method 2): factor, or NaN when the
    //     coordinate is outside the seed grid (calc_seed_inline's range test, :235-237)
    const double *seed_fx, *seed_fy, *seed_fa, *seed_fb, *seed_fv;
    double seed_f0;
    // the raw seed tables (ray_seed_struct x[d], f[d], d = x, y, a, b) for explicit ray lists,
    // where the seed is interpolated per ray on the device
    const double *sd_x[4], *sd_f[4];
    int sd_dim[4];
    // --- constants of the FP64 update (rtb200_fp64.cuh), kept in the kernel parameter bank
    double kfp[RTB_K_COUNT];
    const double *kfp_g; // the same constants in the staged blob (read once with volatile loads)
};

} // namespace rtb
