/*
 * rtb200.h — C ABI of the B200-native image-formation path of the XRayTrace
 * `CreateImage` miniapp (reference: Nikhil-Kulkarni/RayTrace-miniapp).
 *
 * This header is the drop-in boundary.  Every entry point cites the reference
 * interface it replaces (paths relative to the reference checkout).  Only plain
 * pointers and sizes cross the boundary: no C++ types, no torch types.
 *
 * The library behind it (librtb200.so) contains hand-written sm_100a CUDA
 * kernels only.  There is no CPU fallback: every compute entry point returns
 * RTB200_ERR_CUDA when no usable device is present.
 *
 * Data model (mirrors src/RayTraceStructures.h, flattened to PODs):
 *   rtb200_ray         <- ray_struct              src/common/RayTraceImageHelper.h:36-41
 *   rtb200_beam        <- EUV_beam_struct fields used on the path (src/RayTraceStructures.h:26-56)
 *                         and the grid part of seed_beam_struct (src/RayTraceStructures.h:150-180)
 *   rtb200_gain_plane  <- ray_gain_struct         src/RayTraceStructures.h:218-230
 *   rtb200_seed        <- ray_seed_struct         src/RayTraceStructures.h:276-281
 *   rtb200_problem     <- create_image_struct     src/RayTraceStructures.h:321-338
 */
#ifndef RTB200_H
#define RTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Limits of the reference (src/common/RayTraceImageHelper.h:29-32).  The CUDA path itself
 * has no N or K limit; rtb200_create_image enforces the reference's limits for drop-in
 * behaviour unless RTB200_FLAG_NO_LIMITS is given. */
#define RTB200_N_MAX 20
#define RTB200_K_MAX 100
#define RTB200_N_SUB 3
#define RTB200_N_FAILED_MAX 32

/* Return codes.  The reference aborts the process (RAY_ERROR -> exit(-1),
 * src/utilities/RayUtilityMacros.h:16-24); this ABI never exits, it returns. */
#define RTB200_OK 0
#define RTB200_RAYS_FAILED 1   /* failure_code != 0: reference prints "Some rays failed" (src/RayTraceImage.cpp:427-430) */
#define RTB200_ERR_LIMITS (-1) /* "Exceeded maximum number of length segments / frequencies" (src/RayTraceImage.cpp:229-232) */
#define RTB200_ERR_GRID (-2)   /* "Only uniform grid spacings are currently supported" (src/RayTraceImage.cpp:243-264) */
#define RTB200_ERR_CUDA (-3)   /* CUDA runtime error / no device (reference: CUDA_CHECK -> exit, src/RayTraceImageCuda.cu:8-18) */
#define RTB200_ERR_ARG (-4)    /* NULL / inconsistent argument */
#define RTB200_ERR_FORMAT (-5) /* malformed .dat byte stream */

/* Flags for rtb200_create_image / rtb200_trace_rays */
#define RTB200_FLAG_NO_LIMITS 0x1u /* do not enforce N <= N_MAX, nv < K_MAX */
/* rtb200_stage only: the lineshape tables (nine tenths of the bytes, read by the integration only)
 * are packed and uploaded by the first launch, next to its running march kernel, as
 * rtb200_create_image does.  The rtb200_problem and its arrays must stay valid until that launch
 * returns. */
#define RTB200_FLAG_LAZY_TABLES 0x2u


typedef struct rtb200_ray {
    float x, y, a, b;
} rtb200_ray;

typedef struct rtb200_beam {
    int32_t nx, ny, na, nb, nv; /* nv / dv / dz are unused for a seed_beam grid */
    double dx, dy, da, db, dz;
    const double *x, *y, *a, *b; /* cell centres, uniform */
    const double *dv;            /* [nv] frequency bin widths */
} rtb200_beam;

typedef struct rtb200_gain_plane {
    int32_t Nx, Ny, Nv;
    const double *x;  /* [Nx] */
    const double *y;  /* [Ny] */
    const double *n;  /* [Nx*Ny]  index of refraction, ix + iy*Nx */
    const float *g0;  /* [Nx*Ny]  line-centre gain */
    const float *E0;  /* [Nx*Ny]  line-centre emissivity (may be NULL) */
    const float *gv;  /* [Nx*Ny*Nv] lineshape, (ix + iy*Nx)*Nv + k */
} rtb200_gain_plane;

typedef struct rtb200_seed {
    int32_t dim[5];     /* x, y, a, b, v */
    const double *x[5]; /* grids */
    const double *f[5]; /* separable factors */
    double f0;
} rtb200_seed;

typedef struct rtb200_problem {
    int32_t N;          /* number of length planes (gain[0..N-1]) */
    int32_t N_start;    /* first ray of this worker (strided decomposition) */
    int32_t N_parallel; /* ray stride (>= 1) */
    const rtb200_beam *euv_beam;
    const rtb200_beam *seed_beam; /* NULL for ASE */
    const rtb200_gain_plane *gain; /* [N] */
    const rtb200_seed *seed;       /* NULL for ASE */
} rtb200_problem;

/* Per-call device timings (CUDA events on the context's stream), milliseconds. */
typedef struct rtb200_timings {
    float h2d_ms;       /* staging upload */
    float march_ms;     /* refractive march kernel(s); ASE grid launches overlap the march's tail with the
                           integration (programmatic dependent launch): there this is the time of BOTH
                           kernels and integrate_ms is 0 - RTB200_OVERLAP=0 runs them one after the other */
    float integrate_ms; /* frequency integration + binning kernel(s) */
    float d2h_ms;       /* image / I_ang download */
    float total_ms;     /* first event to last event */
    int32_t kernel_launches;
    int32_t reserved;
    uint64_t n_rays;
    uint64_t march_steps; /* device counter: inner `propagate` steps taken (0 unless RTB200_COUNT_STEPS) */
} rtb200_timings;

typedef struct rtb200_ctx rtb200_ctx; /* opaque; one per host thread / device; not shared between threads */

/* ---- context ------------------------------------------------------------------------------ */

/* Create a context bound to CUDA device `device` (own stream, pinned staging, device arena).
 * Replaces the reference's per-call cudaMalloc/cudaFree and its non-thread-safe static
 * state (src/RayTraceImageCuda.cu:145-221, :153-160) and setGPU (src/RayTraceImage.cpp:82-88). */
int rtb200_create(int device, rtb200_ctx **ctx);
void rtb200_destroy(rtb200_ctx *ctx);
const char *rtb200_last_error(const rtb200_ctx *ctx); /* never NULL */
int rtb200_device_count(void);                        /* 0 when no CUDA device is usable */
const char *rtb200_version(void);

/* ---- the path, host buffers (H2D + kernels + D2H inside the call) --------------------------- */

/* Replaces RayTrace::create_image (src/RayTraceImage.cpp:227-434, declared src/RayTrace.h:93):
 * limits + uniform-grid validation, method/scale selection, ray enumeration
 * (ijkm = N_start + it*N_parallel, b fastest), trace, binning, failure report.
 * image[nx*ny*nv] and I_ang[na*nb] are caller-allocated and are OVERWRITTEN (the reference
 * callocs them itself, src/RayTraceImage.cpp:271-274).  failed may be NULL. */
int rtb200_create_image(rtb200_ctx *ctx, const rtb200_problem *problem, unsigned flags,
                        double *image, double *I_ang, unsigned *failure_code,
                        rtb200_ray *failed, int max_failed, int *n_failed);

/* Replaces the RayTraceImage<Backend>Loop back-end signature (extern at
 * src/RayTraceImage.cpp:47-75; CPU body src/RayTraceImageCPU.cpp:19-70): trace an explicit
 * ray list.  `beam` is the euv_beam (output grids, dv, dz).  method 1 = backward (ASE),
 * 2 = forward (seeded).  image / I_ang are ACCUMULATED into (+=), like the CPU loop. */
int rtb200_trace_rays(rtb200_ctx *ctx, int N, const rtb200_beam *beam,
                      const rtb200_gain_plane *gain, const rtb200_seed *seed, int method,
                      const rtb200_ray *rays, size_t n_rays, double scale, double *image,
                      double *I_ang, unsigned *failure_code, rtb200_ray *failed,
                      int max_failed, int *n_failed);

/* Replaces RayTrace::calc_ray for a batch (src/RayTraceImage.cpp:189-204 -> RayTrace_calc_ray,
 * src/common/RayTraceImageHelper.h:379-595): per-ray outputs without binning.
 * Iv[n_rays*K], ray2[n_rays], error[n_rays] (0, -1, -2, -3); optional march intermediates
 * gvl/evl[n_rays*(N-1)*3] and ivl[n_rays*(N-1)*3] in the reference's [i][is] order. */
int rtb200_calc_rays(rtb200_ctx *ctx, int N, double dz, const rtb200_gain_plane *gain,
                     const rtb200_seed *seed, int K, int method, const rtb200_ray *rays,
                     size_t n_rays, double *Iv, rtb200_ray *ray2, int *error, float *gvl,
                     float *evl, int32_t *ivl);

/* Replaces RayTrace::calc_ray_path (src/RayTrace.h:69-72, src/RayTraceImage.cpp:440-477; the
 * RAY_DEBUG path of RayTrace_calc_ray, src/common/RayTraceImageHelper.h:419-426, :505-511,
 * :536-566): the trajectory of every ray.  xr, yr, Ir are [n_rays][N_SUB*(N-1)+1] floats in the
 * order of `rays` (position at every sub-segment boundary, intensity sum_k 2*Iv[k]*dv[k] after
 * every sub-segment; emission-style integration as the reference does whenever it records a
 * trajectory); error[n_rays] = 0, -1, -2, -3.  c = step safety factor (0.5 in create_image).
 * K <= 128.  Returns RTB200_RAYS_FAILED when any ray failed (the reference returns the count). */
int rtb200_calc_ray_paths(rtb200_ctx *ctx, int N, double dz, const rtb200_gain_plane *gain,
                          const rtb200_seed *seed, int K, const double *dv, int method, double c,
                          const rtb200_ray *rays, size_t n_rays, float *xr, float *yr, float *Ir,
                          int *error);

/* ---- the path, device-resident (for throughput measurement and multi-GPU tiling) ----------- */

/* Upload + re-layout the problem into the context's device arena (one packed SoA blob, one
 * H2D copy).  Replaces ray_gain_struct::copy_device / ray_seed_struct::copy_device
 * (src/RayTraceStructures.cpp:1432-1489, 2049-2136; src/RayTraceImageCuda.cu:225-329). */
int rtb200_stage(rtb200_ctx *ctx, const rtb200_problem *problem, unsigned flags);

/* Number of source pixels (ASE: euv nx*ny; seeded: seed_beam nx*ny) of the staged problem:
 * the unit the path is sharded by. */
int64_t rtb200_staged_pixels(const rtb200_ctx *ctx);
int64_t rtb200_staged_rays(const rtb200_ctx *ctx);

/* Shape of the staged problem, for callers that shard it (rtb200_multi, bench.py).
 * owner = 1: ASE with one source pixel per destination pixel, traced by the pixel-owner kernel:
 * image rows of different devices are disjoint and can be gathered (rtb200_launch_rows_compact +
 * rtb200_unpermute_rows); owner = 0 (seeded, or a non-injective pixel map): partial images
 * overlap and must be summed. */
typedef struct rtb200_staged {
    int32_t method, owner;
    int32_t snx, sny;    /* source grid (euv_beam for ASE, seed_beam for seeded) */
    int32_t nx, ny, na, nb, nv; /* destination grid */
    int32_t reserved;
} rtb200_staged;
int rtb200_staged_info(const rtb200_ctx *ctx, rtb200_staged *out);

/* Trace the staged problem's source pixels [pix_begin, pix_end) on the context's stream.
 * d_image / d_I_ang are DEVICE pointers to full-size buffers (nx*ny*nv, na*nb doubles).
 * ASE: writes (overwrites) image rows of the owned pixels only and adds into d_I_ang;
 * seeded: adds into both.  The caller zeroes the buffers.  Asynchronous; pair with
 * rtb200_sync.  `cuda_stream` = 0 uses the context's own stream, else a cudaStream_t. */
int rtb200_launch(rtb200_ctx *ctx, int64_t pix_begin, int64_t pix_end, double *d_image,
                  double *d_I_ang, void *cuda_stream);
/* Same, for the image rows j = row_offset, row_offset + row_stride, ... (source-grid rows).
 * This is the multi-GPU decomposition: rank r of W traces rows r, r + W, ... so that the
 * ranks' work is balanced; results land at their true positions in the full-size buffers,
 * and the exchange is a sum (the rows of different ranks are disjoint).  It plays the role of
 * the reference's strided N_start / N_parallel decomposition (src/RayTraceImage.cpp:300-308)
 * at pixel-row granularity, which keeps every pixel's spectrum on one device. */
int rtb200_launch_rows(rtb200_ctx *ctx, int row_offset, int row_stride, double *d_image,
                       double *d_I_ang, void *cuda_stream);
/* The same share, written COMPACTLY: d_rows receives the device's rows back to back
 * (ceil((sny - row_offset) / row_stride) rows of snx*nv doubles, in source-grid order), which is
 * what a gather over the devices moves: 1/row_stride of the image instead of a full-size buffer
 * per device.  Only for owner-traced problems (rtb200_staged_info).  d_I_ang: as above. */
int rtb200_launch_rows_compact(rtb200_ctx *ctx, int row_offset, int row_stride, double *d_rows,
                               double *d_I_ang, void *cuda_stream);
/* Puts gathered compact rows where they belong: d_gathered holds `world` blocks of
 * rows_per_dev*snx*nv doubles (block r = the rows of device r, rows_per_dev >= ceil(sny/world));
 * d_image is the full image (nx*ny*nv doubles, zeroed by the caller: destination pixels that no
 * source pixel owns are not written).  The exchange step of the multi-device ASE path; replaces
 * the sum of full-size partial images of intensity_step_struct::sum_reduce
 * (src/RayTraceStructures.cpp:1603-1646) by a gather of owned rows. */
int rtb200_unpermute_rows(rtb200_ctx *ctx, const double *d_gathered, int world, int64_t rows_per_dev,
                          double *d_image, void *cuda_stream);
int rtb200_sync(rtb200_ctx *ctx, unsigned *failure_code, rtb200_ray *failed, int max_failed,
                int *n_failed);
int rtb200_get_timings(const rtb200_ctx *ctx, rtb200_timings *out);
/* Forget the launches recorded so far: the timings returned after the next rtb200_sync then
 * cover exactly the launches issued in between (used to time a series of steps). */
int rtb200_reset_timings(rtb200_ctx *ctx);

/* ---- several devices of one box (single process) --------------------------------------------- */

/* Replaces the reference's `cuda-multigpu` method (src/RayTraceImage.cpp:396-405:
 * RayTraceImageThreadLoop with one worker per GPU and a host-side sum of partial images; its
 * cudaSetDevice in the parent thread, :116-119, puts every worker on device 0) and the
 * application's cross-rank intensity_step_struct::sum_reduce (src/RayTraceStructures.cpp:
 * 1603-1646).  One context per device, the image sharded by rows (device r of W traces rows r,
 * r + W, ...), the partial results gathered / reduced to the first device over NVLink by NCCL
 * (ncclCommInitAll; libnccl.so.2 is bound at run time, n_dev == 1 needs none), one download.
 * devices == NULL means 0 .. n_dev-1.  On RTB200_ERR_CUDA *out may still be non-NULL so that
 * rtb200_multi_last_error() can tell why; destroy it either way. */
typedef struct rtb200_multi rtb200_multi;
int rtb200_multi_create(const int *devices, int n_dev, rtb200_multi **out);
void rtb200_multi_destroy(rtb200_multi *m);
const char *rtb200_multi_last_error(const rtb200_multi *m); /* never NULL */
int rtb200_multi_device_count(const rtb200_multi *m);
/* Same contract as rtb200_create_image (host buffers in and out, image / I_ang overwritten).
 * ASE images are bit-identical to the single-device result (rows are owned, never summed);
 * I_ang and seeded images are sums of per-device partials (~1e-16 relative). */
int rtb200_multi_create_image(rtb200_multi *m, const rtb200_problem *problem, unsigned flags,
                              double *image, double *I_ang, unsigned *failure_code,
                              rtb200_ray *failed, int max_failed, int *n_failed);
/* Timings of the last call: per_device[n_dev] (may be NULL), the exchange step on the first
 * device (NCCL + un-permute) and first event to last event on the first device. */
int rtb200_multi_get_timings(const rtb200_multi *m, rtb200_timings *per_device, float *exchange_ms,
                             float *total_ms);

/* ---- wire format ---------------------------------------------------------------------------- */

/* Parse a serialized create_image_struct (the payload of a .dat file after its uint64 length,
 * src/CreateImage.cpp:26-58; format src/RayTraceStructures.cpp:2159-2292 and the nested
 * pack() functions) into a problem that owns copies of the arrays on the path (the byte
 * stream is not aligned).  golden_image / golden_I_ang receive pointers to the embedded
 * golden arrays or NULL.  Free everything with rtb200_free_problem. */
int rtb200_parse_dat(const void *bytes, size_t n_bytes, rtb200_problem **problem,
                     const double **golden_image, const double **golden_I_ang);
void rtb200_free_problem(rtb200_problem *problem);

/* RayTrace::create_image straight from the serialized form (what the reference's driver does
 * with loadInput + create_image, src/CreateImage.cpp:26-58, :147-152): the large arrays of every
 * gain plane are read ONCE, from the (unaligned) byte stream into the pinned staging blob, and
 * uploaded - file buffer -> pinned blob -> device, no intermediate AoS-of-pointers copy
 * (create_image_struct::unpack, src/RayTraceStructures.cpp:2224-2292).  `bytes` is the payload
 * after the file's uint64 length and must stay valid during the call. */
int rtb200_create_image_from_dat(rtb200_ctx *ctx, const void *bytes, size_t n_bytes, unsigned flags,
                                 double *image, double *I_ang, unsigned *failure_code,
                                 rtb200_ray *failed, int max_failed, int *n_failed);

/* create_image_struct::pack (src/RayTraceStructures.cpp:2159-2223): serializes a problem (plus
 * optional golden arrays) into the payload of a .dat file that the reference's own CreateImage
 * loads.  Fields off the image-formation path get neutral values (one z plane, v = 0, no seed
 * shapes, gv0 = 0).  out == NULL only measures; *n_bytes always receives the size needed. */
int rtb200_write_dat(const rtb200_problem *problem, const double *golden_image,
                     const double *golden_I_ang, void *out, size_t capacity, size_t *n_bytes);

/* ---- measurement helpers ---------------------------------------------------------------------- */

/* DFMA micro-benchmark: returns the measured FP64 instruction rate (warp-level FMA
 * instructions * 32 lanes per second, i.e. FP64 lane-instr/s) of the context's device. */
int rtb200_measure_fp64_peak(rtb200_ctx *ctx, double *fp64_lane_instr_per_s);

/* Proof by exhaustion for the march's branch-free FP32 division (csrc/rtb200_math.cuh,
 * fdiv_refined): compares it with the IEEE division on the device for the divisor significands
 * [b_first, b_first + b_count) (of 2^23) times ALL 2^23 numerator significands, operands scaled
 * by 2^exp_a / 2^exp_b (-60 <= exp < 60).  Returns the number of differing quotients and one
 * offending pair.  The whole significand space takes about 75 s on a B200.  variant 0 is the
 * division the march uses; variant 1 omits its correction step (a self-test of the detector:
 * it must report mismatches); variant 2 checks the branch-free square root and the reciprocal
 * that follows it in normalize_s (fsqrt_refined) for the significands [b_first, b_first +
 * b_count) at exponents exp_b and exp_b + 1 against the IEEE results (exp_a is ignored);
 * variant 3 is variant 0 through the packed two-quotient form the step uses (FFMA2, both lanes). */
int rtb200_check_fdiv(rtb200_ctx *ctx, unsigned b_first, unsigned b_count, int exp_a, int exp_b,
                      int variant, unsigned long long *mismatches, float *witness_a, float *witness_b);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
