// rtb200_kernels.cuh — launch interface between the host layer (rtb200_host.cu) and the
// sm_100a kernels (rtb200_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "rtb200_device.cuh"

namespace rtb {

// A contiguous range of source pixels [pix0, pix1) in x-fastest order (p = i + j*snx), or, in
// list mode, a contiguous range of explicit rays.
struct Chunk {
    long long pix0, pix1;   // grid mode: LOGICAL pixel range; physical pixel = phys_pixel(q)
    int row_off, row_stride; // row-cyclic sharding: physical row = row_off + (q / snx)*row_stride
    long long ray0, ray1;   // list mode (explicit rays)
    const float4 *rays;     // list mode: (x, y, a, b) per ray
    const float2 *tans;     // list mode: host tanf(1e-3f*a), tanf(1e-3f*b) per ray
};

struct Handoff {
    SegRec *seg;        // [slots * S]
    unsigned *meta;     // [slots]
    float4 *exit_ray;   // [slots] ray2 = (x, y, a, b) at exit; may be null (ASE binning never reads it)
    float2 *path;       // [slots * ((N-1)*3 + 1)] RAY_DEBUG trajectory (x, y); null except for calc_ray_paths
    // Grid mode, optional: rays of each logical pixel of the chunk whose hand-off is complete
    // (counted by the march with release semantics).  Lets the integration of a pixel begin while
    // the march is still finishing other pixels (programmatic dependent launch); null = not counted.
    unsigned *pix_done;
};

// Output selection of the integration kernel.
struct Outputs {
    double *image;   // [nx*ny*nv]
    double *I_ang;   // [na*nb]
    double *Iv;      // per-ray dump [slots*K] (rtb200_calc_rays) or null
    int *error;      // per-ray error code dump or null
    FailState *fail; // never null
    // Owner kernel only: image holds the chunk's LOGICAL pixels back to back (pixel q of the
    // launch at image[q*K]) instead of the full image in destination order.  This is what one
    // device of a row-cyclic multi-device launch produces: its rows compacted, ready to be
    // gathered; rtb200_unpermute_rows puts the gathered rows where they belong.
    int compact;
    // Owner kernel only: when not null, the CTA of logical pixel q waits until pix_done[q] has
    // reached ab_max (see Handoff::pix_done) before it reads the pixel's hand-off.
    const unsigned *pix_done;
    // Overlapped launch while the lineshape tables are still being uploaded on the copy stream
    // (rtb200_create_image): when not null, the CTA also waits until *gv_flag == gv_epoch - the word
    // is written by a 4-byte copy that follows the tables' copy on the same stream.
    const unsigned *gv_flag;
    unsigned gv_epoch;
};

// Persistent flat-state-machine march with refill (work = device counter, reset by the
// launcher).  persistent_blocks <= 0: resident CTAs per SM x SMs of the current device.
void launch_march(const DevProblem &P, const Chunk &c, bool list_mode, const Handoff &h,
                  FailState *fail, bool count_steps, cudaStream_t st, unsigned long long *work,
                  int persistent_blocks);
// ASE (method 1, emission + gain), grid mode: one CTA per source pixel, the pixel's spectrum is
// owned by the CTA (plain stores), I_ang by atomics.
// overlap: launch with programmatic stream serialization (the kernel may start as soon as every
// CTA of the march that precedes it in the stream has signalled or exited; o.pix_done must be set).
void launch_integrate_ase_owner(const DevProblem &P, const Chunk &c, const Handoff &h,
                                const Outputs &o, cudaStream_t st, bool overlap = false);
// Generic: one warp per ray slot, scatter binning with FP64 atomics (list mode, seeded mode,
// non-identity owner maps) and/or per-ray dumps.
void launch_integrate_scatter(const DevProblem &P, const Chunk &c, bool list_mode,
                              const Handoff &h, const Outputs &o, cudaStream_t st);
// Gathered compact rows of a row-cyclic launch over `world` devices -> full image in
// destination order (skips source pixels that own no destination pixel).
void launch_unpermute_rows(const DevProblem &P, const double *gathered, int world, long long rows_per_dev,
                           double *image, cudaStream_t st);
void launch_path_intensity(const DevProblem &P, const Chunk &c, const Handoff &h, float *path_I,
                           int *error, cudaStream_t st);
// Derives the per-cell records of every plane from the uploaded nodes and interval tables
// (what rtb200_pack.h's fill_cell_records does on the host, operation for operation).
void launch_build_cell_records(const DevPlane *planes, int N, long long max_nodes, cudaStream_t st);
// Gain-only problems: DevPlane::gvd[i] = (double) DevPlane::gv[i] for every plane (K bins per node).
void launch_widen_gv(const DevPlane *planes, int N, int K, cudaStream_t st);
void launch_fp64_peak(double *out, int iters, cudaStream_t st, int *blocks, int *threads);
void launch_fdiv_check(unsigned b_first, unsigned b_count, int ea, int eb, int variant,
                       unsigned long long *out, cudaStream_t st);

} // namespace rtb
