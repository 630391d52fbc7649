// rtb200_kernels.cu — hand-written sm_100a kernels of the image-formation path.
//
//   march_flat_kernel          persistent grid, lane = ray: refractive march through the gain
//                              planes as a flat state machine with batched refills (FP32 +
//                              mixed FP64, bit-exact with the reference); hands gvl/evl/ivl per
//                              (segment, sub-segment) to the integration through L2 / HBM.
//   integrate_ase_owner_kernel one CTA per source pixel, one warp per ray, lanes = frequency
//                              bins: ASE gain + emission integration (FP64), per-pixel spectrum
//                              accumulated in registers (no atomics), I_ang by warp-shuffle
//                              reduction + one FP64 atomic per ray.
//   integrate_seeded_kernel    the seeded image of create_image (grid mode, gain-only): one warp
//                              per 64 ray slots, lanes = frequency bins, per-ray scalars by one lane
//                              per ray, binning by the exit ray with run-length combined atomics.
//   integrate_scatter_kernel   one warp per ray, lanes = frequency bins: list-mode rays, emission
//                              along listed rays, K > 128 seeded; binning by the exit ray with FP64
//                              atomics; also the per-ray dump used by rtb200_calc_rays.
//
// ASE grid launches overlap the two kernels: the march counts the closed ray slots of every pixel
// (Handoff::pix_done) and releases its dependents when its work queue is empty; the owner kernel
// is launched with programmatic stream serialization and each of its CTAs waits for its pixel.
//
// There are no tensor-core instructions here on purpose: the path is not a dense contraction
// (SURVEY.md §8d); the binding unit is the FP64 pipe.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <cstring>
#include <type_traits>

#include "rtb200_fp64.cuh"
#include "rtb200_kernels.cuh"
#include "rtb200_march_flat.cuh"

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
    const unsigned pu = (unsigned) p; // pixel indices fit 32 bits (validated on the host)
    r.j = (int) (pu / (unsigned) P.snx);
    r.i = (int) (pu - (unsigned) r.j * (unsigned) P.snx);
    const int AB32 = P.sna * P.snb;
    if (P.n_parallel == 1 && P.n_start == 0) { // the common case: every ray of the pixel
        r.ab0 = 0;
        r.cnt = AB32;
        return r;
    }
    const long long AB = (long long) AB32;
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

// Logical -> physical source pixel.  Contiguous tiles use row_stride = 1 (identity); the
// multi-GPU sharding gives rank r the image rows r, r + world, r + 2*world, ... so that the
// ranks' work is balanced (rows near the target surface escape early, rows far from it do not).
__device__ __forceinline__ long long phys_pixel(const DevProblem &P, const Chunk &c, long long q)
{
    if (c.row_stride == 1)
        return q;
    const long long jr = q / P.snx;
    return ((long long) c.row_off + jr * c.row_stride) * P.snx + (q - jr * P.snx);
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
// Writes the hand-off record of one (segment, sub-segment) of ray slot L; with PATH also the
// RAY_DEBUG trajectory points (rtb200_calc_ray_paths).  Holds no state of its own: the record
// address is formed from the slot number when a record is emitted (once per ~25 trips).
template <bool PATH>
struct GlobalSinkT {
    SegRec *seg;  // hand-off arena
    float2 *path; // trajectories [slot][S + 1]; unused unless PATH
    unsigned L;
    int S;
    unsigned rec0; // L*S: record numbers of one chunk fit 32 bits (the host sizes chunks that way)
    __device__ __forceinline__ void operator()(int idx, float gvl, float evl, int cell) const
    {
        int4 v;
        v.x = __float_as_int(gvl);
        v.y = __float_as_int(evl);
        v.z = cell;
        v.w = 0;
        *reinterpret_cast<int4 *>(&seg[rec0 + (unsigned) idx]) = v;
    }
    __device__ __forceinline__ void point(int idx, float x, float y) const
    {
        if (PATH)
            path[(size_t) L * (size_t) (S + 1) + (size_t) idx] = make_float2(x, y);
    }
};

// Source coordinates of ray slot L (grid mode: decoded from the slot number; list mode: the
// uploaded ray) and the host-evaluated tangents of its angles.  `active` is false for the
// padding slots of a strided worker's last pixel.
template <bool LIST>
__device__ __forceinline__ bool slot_source(const DevProblem &P, const Chunk &c, unsigned L, float &rx,
                                            float &ry, float &ra, float &rb, float &ta, float &tb)
{
    if (LIST) {
        const float4 r = __ldg(&c.rays[c.ray0 + L]);
        const float2 t = __ldg(&c.tans[c.ray0 + L]);
        rx = r.x, ry = r.y, ra = r.z, rb = r.w, ta = t.x, tb = t.y;
        return true;
    }
    const unsigned lq = L / (unsigned) P.ab_max; // slots fit 32 bits
    const long long p = phys_pixel(P, c, c.pix0 + lq);
    const int t = (int) (L - lq * (unsigned) P.ab_max);
    const PixelRays pr = pixel_rays(P, p);
    const bool active = t < pr.cnt;
    const int ab = pr.ab0 + t * (int) P.n_parallel;
    const int k = active ? ab / P.snb : 0, mm = active ? ab % P.snb : 0;
    rx = __ldg(&P.sxf[pr.i]);
    ry = __ldg(&P.syf[pr.j]);
    ra = __ldg(&P.saf[k]);
    rb = __ldg(&P.sbf[mm]);
    ta = __ldg(&P.tanA[k]);
    tb = __ldg(&P.tanB[mm]);
    return active;
}

// Persistent march: a fixed grid of warps pulls ray slots from a global counter.  Lane = ray,
// flat state machine (rtb200_march_flat.cuh); a lane whose ray has finished claims the next
// unprocessed slot (one warp-aggregated atomic per run of slots), so lanes stay busy although
// the number of steps per ray spans 1..~600 and more than half of the rays of ASE_medium leave
// the plasma early.  The plane descriptors every cell look-up starts from and the sub-segment
// limits are staged in shared memory once per CTA.
#ifndef RTB_MARCH_MINBLOCKS
#define RTB_MARCH_MINBLOCKS 6
#endif
#ifndef RTB_MARCH_CHUNK
#define RTB_MARCH_CHUNK 32
#endif
#ifndef RTB_REFILL_MIN
#define RTB_REFILL_MIN 6
#endif
#define RTB_MARCH_THREADS 128
#define RTB_ST_HUNG 64u  // gave up on the ray (hang guard): reported as invalid
#define RTB_ST_DEAD 128u // the lane has no ray and there is none left to claim
// Overlapped integration (Handoff::pix_done): one more closed ray slot of a pixel.  The slot's
// records and meta word were stored by THIS thread, so a release reduction orders them before the
// count; no fence of wider scope (__threadfence() would also invalidate the SM's L1 - CCTL.IVALL -
// and with it the march's tables, once per ray).
__device__ __forceinline__ void count_closed_slot(unsigned *counter)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}

template <bool LIST, bool PATH, bool COUNT, bool OVL = false>
__global__ void __launch_bounds__(RTB_MARCH_THREADS, RTB_MARCH_MINBLOCKS)
    march_flat_kernel(const DevProblem P, const Chunk c, const Handoff h, FailState *fail,
                      unsigned long long *work)
{
    extern __shared__ __align__(16) unsigned char march_smem[]; // PlaneLite[N]
    __shared__ float s_zt[4];
    {
        const int4 *src = reinterpret_cast<const int4 *>(P.lite);
        int4 *dst = reinterpret_cast<int4 *>(march_smem);
        const int n16 = P.N * (int) (sizeof(PlaneLite) / 16);
        for (int i = threadIdx.x; i < n16; i += blockDim.x)
            dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < RTB_N_SUB)
        s_zt[threadIdx.x] = march_sub_limit(threadIdx.x, P.dz0);
    __syncthreads();
    // the shuffles make the shared-memory addresses opaque: otherwise the compiler rebuilds them
    // from the CTA's shared window (S2R + LEA) at every use
    MarchConsts K;
    march_consts(K, __shfl_sync(0xffffffffu, (unsigned) __cvta_generic_to_shared(march_smem), 0),
                 __shfl_sync(0xffffffffu, (unsigned) __cvta_generic_to_shared(s_zt), 0), P.N, P.method, P.c,
                 P.use_emis != 0);
    const int lane = threadIdx.x & 31;
    const int S = (P.N - 1) * RTB_N_SUB;
    // slot counts of one chunk fit 31 bits (the host sizes chunks that way)
    const unsigned n_slots = (unsigned) (LIST ? (c.ray1 - c.ray0) : (c.pix1 - c.pix0) * P.ab_max);
    FlatMarch m;
    m.st = PH_DONE;
    m.steps = 0;
    bool exhausted = false;
    unsigned L = 0, run_next = 0, run_end = 0;
    unsigned total_steps = 0;
    bool pending = false; // a marched ray whose hand-off entry has not been closed yet
    // Closes the lane's finished ray: meta word, failure report, exit ray (the two atanf of the
    // seeded path).  Runs at the lane's next refill, i.e. for RTB_REFILL_MIN or more lanes at a
    // time, instead of right after the trip in which a single lane happened to finish.
    auto finalize = [&]() {
        int lo, hi;
        flat_visited_range(m, K, lo, hi);
        unsigned meta = (unsigned) lo | ((unsigned) hi << 12);
        if (flat_escaped(m))
            meta |= RTB_META_ESCAPED;
        if (lt_0p01(fmul(m.s.z, m.s.z)) || (m.st & RTB_ST_HUNG)) { // error -1 (:515-516)
            meta |= RTB_META_INVALID;
            float rx, ry, ra, rb, ta, tb;
            slot_source<LIST>(P, c, L, rx, ry, ra, rb, ta, tb);
            report_failure(fail, 1, rx, ry, ra, rb);
        } else if (h.exit_ray) {
            float4 e;
            e.x = m.pos.x;
            e.y = m.pos.y;
            e.z = fmul(atanf_fdlibm(fdiv(m.s.x, m.s.z)), 1e3f);
            e.w = fmul(atanf_fdlibm(fdiv(m.s.y, m.s.z)), 1e3f);
            h.exit_ray[L] = e;
        }
        h.meta[L] = meta;
        if (OVL) // this ray's records and meta word first, then the count
            count_closed_slot(h.pix_done + L / (unsigned) P.ab_max);
        if (COUNT)
            total_steps += m.steps;
        pending = false;
    };
    bool signalled = false; // (warp-uniform) dependents may be launched: this warp sees no more work
    for (;;) {
        const bool need = (m.st & (RTB_ST_PHASE | RTB_ST_DEAD)) == (unsigned) PH_DONE;
        const unsigned want = __ballot_sync(0xffffffffu, need);
        // Refills run for at least RTB_REFILL_MIN lanes at a time (or when the warp has nothing
        // else to do): the ~200 instructions of a ray start are issued for the whole warp.
        if (want != 0u && (RTB_REFILL_MIN <= 1 || __popc(want) >= RTB_REFILL_MIN ||
                           __ballot_sync(0xffffffffu, flat_phase(m) != PH_DONE) == 0u)) {
            // The warp owns a run of RTB_MARCH_CHUNK consecutive slots and hands them to its
            // lanes; one atomic per run instead of one per refill, and the rays a warp marches
            // together are neighbours (same source pixel, adjacent angles), so they cross the
            // same cells at about the same time.
            if (run_next >= run_end && !exhausted) { // warp-uniform
                unsigned long long first = 0;
                if (lane == 0)
                    first = atomicAdd(work, (unsigned long long) RTB_MARCH_CHUNK);
                first = __shfl_sync(0xffffffffu, first, 0);
                run_next = first < (unsigned long long) n_slots ? (unsigned) first : n_slots;
                run_end = n_slots - run_next > (unsigned) RTB_MARCH_CHUNK ? run_next + RTB_MARCH_CHUNK : n_slots;
                if (run_next >= n_slots)
                    exhausted = true;
            }
            if (OVL && exhausted && !signalled) {
                // the work queue is empty: the integration kernel may start filling the SMs that
                // the march's last CTAs leave (it waits per pixel on Handoff::pix_done)
                asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
                signalled = true;
            }
            const unsigned mine = run_next + __popc(want & ((1u << lane) - 1u));
            const unsigned after = run_next + __popc(want);
            if (need && pending)
                finalize();
            if (need) {
                if (mine >= run_end) {
                    if (exhausted) // otherwise: served from the next run on the next trip
                        m.st |= RTB_ST_DEAD;
                } else {
                    L = mine;
                    float rx, ry, ra, rb, ta, tb;
                    const bool active = slot_source<LIST>(P, c, L, rx, ry, ra, rb, ta, tb);
                    if (!active) {
                        h.meta[L] = RTB_META_INACTIVE;
                        if (OVL)
                            count_closed_slot(h.pix_done + L / (unsigned) P.ab_max);
                    } else {
                        if (PATH) // trajectory start point (:419-426)
                            h.path[(size_t) L * (size_t) (S + 1) + (size_t) (P.method == 1 ? S : 0)] =
                                make_float2(rx, ry);
                        flat_init(m, K, rx, ry, ta, tb);
                        // (N == 1: nothing to march, but the ray is still closed by finalize():
                        // exit ray = start ray, error -1 test, RayTraceImageHelper.h:515-521)
                        pending = true;
                    }
                }
            }
            run_next = after < run_end ? after : run_end;
        }
        if (__ballot_sync(0xffffffffu, flat_phase(m) != PH_DONE) == 0u) {
            if (__all_sync(0xffffffffu, (m.st & RTB_ST_DEAD) != 0u))
                break;
            continue;
        }
        // Inner loop: trips until enough lanes have finished for a batched refill (or the
        // warp has run dry).  Keeping the refill code out of this loop keeps its live ranges
        // out of the hot path.
        GlobalSinkT<PATH> sink{ h.seg, h.path, L, S, L * (unsigned) S };
        // lanes without a ray and with none left to claim (set by the refill above only)
        const unsigned dead = __ballot_sync(0xffffffffu, (m.st & RTB_ST_DEAD) != 0u);
        unsigned marching = __ballot_sync(0xffffffffu, flat_phase(m) != PH_DONE);
        unsigned trips = 0; // (warp-uniform counter)
        do {
            // every lane takes the trip (finished lanes fall through): see flat_trip
            flat_trip(m, K, sink, marching, !exhausted);
            // ONE vote per trip: the lanes that are marching; the others can take a new ray
            // unless they are dead.  The loop ends when enough lanes wait for a refill, when the
            // warp has run dry, or at the hang guard (no ray takes 2^22 trips).
            marching = __ballot_sync(0xffffffffu, flat_phase(m) != PH_DONE);
        } while (marching != 0u && __popc(~(marching | dead)) < RTB_REFILL_MIN && ++trips <= (1u << 22));
        if (trips > (1u << 22) && flat_phase(m) != PH_DONE) // whatever is still marching is given up
            m.st = (m.st | RTB_ST_HUNG | RTB_ST_PHASE);     // PH_DONE == all phase bits
    }
    if (COUNT) {
        unsigned tot = total_steps;
        for (int o = 16; o > 0; o >>= 1)
            tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0)
            atomicAdd(&fail->march_steps, (unsigned long long) tot);
    }
}

template <bool LIST, bool PATH, bool COUNT, bool OVL = false>
static void launch_march_t(const DevProblem &P, const Chunk &c, const Handoff &h, FailState *fail,
                           cudaStream_t st, unsigned long long *work, int persistent_blocks, long long n)
{
    const size_t smem = sizeof(PlaneLite) * (size_t) P.N;
    auto kern = march_flat_kernel<LIST, PATH, COUNT, OVL>;
    if (smem > 40 * 1024) // deep stacks of planes (hundreds): opt in to the large carve-out
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    long long blocks = persistent_blocks;
    if (blocks <= 0) { // resident CTAs per SM (occupancy) x SMs of the current device
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RTB_MARCH_THREADS, smem);
        blocks = (long long) std::max(1, per_sm) * std::max(1, sms);
    }
    blocks = std::min(blocks, (n + RTB_MARCH_THREADS - 1) / RTB_MARCH_THREADS);
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    kern<<<(unsigned) blocks, RTB_MARCH_THREADS, smem, st>>>(P, c, h, fail, work);
}

void launch_march(const DevProblem &P, const Chunk &c, bool list_mode, const Handoff &h,
                  FailState *fail, bool count_steps, cudaStream_t st, unsigned long long *work,
                  int persistent_blocks)
{
    const long long n = list_mode ? (c.ray1 - c.ray0) : (c.pix1 - c.pix0) * P.ab_max;
    if (n <= 0)
        return;
    if (list_mode && h.path) // trajectories (rtb200_calc_ray_paths)
        launch_march_t<true, true, false>(P, c, h, fail, st, work, persistent_blocks, n);
    else if (list_mode && count_steps)
        launch_march_t<true, false, true>(P, c, h, fail, st, work, persistent_blocks, n);
    else if (list_mode)
        launch_march_t<true, false, false>(P, c, h, fail, st, work, persistent_blocks, n);
    else if (count_steps)
        launch_march_t<false, false, true>(P, c, h, fail, st, work, persistent_blocks, n);
    else if (h.pix_done) // overlapped integration: closed slots counted per pixel, early trigger
        launch_march_t<false, false, false, true>(P, c, h, fail, st, work, persistent_blocks, n);
    else
        launch_march_t<false, false, false>(P, c, h, fail, st, work, persistent_blocks, n);
}

// ------------------------------------------------------------------------------------------
// frequency integration
// ------------------------------------------------------------------------------------------

// Library-precision form of one update (RayTraceImageHelper.h:549-557): only used for the
// rare out-of-range arguments (|gl| >= 700, inf, NaN), where the library exp's overflow /
// underflow / NaN semantics are wanted.
__device__ __noinline__ double ase_update_library(double Iv, double gl, double el)
{
    if (fabs(gl) < 1e-3) {
        return el * (1.0 + 0.5 * gl * (1.0 + 0.3333333333 * gl)) +
               Iv * (1.0 + gl * (1.0 + 0.5 * gl));
    }
    const double e = exp(gl);
    return el / gl * (e - 1.0) + Iv * e;
}

__constant__ double c_exp_table[RTB_EXP_TABLE_SIZE] = { RTB_EXP_TABLE_VALUES };

__device__ __forceinline__ void load_exp_table(double *T)
{
    for (int i = threadIdx.x; i < RTB_EXP_TABLE_SIZE; i += blockDim.x)
        T[i] = c_exp_table[i];
    __syncthreads();
}

__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// A value every lane of the warp holds, re-issued through a warp reduction: the result lives in
// a uniform register, so the compiler KNOWS that branches and loop bounds derived from it are
// warp-uniform (no divergence check in front of the votes, no reconvergence points).
__device__ __forceinline__ unsigned uniform_u32(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
// Lane j's value for the whole warp, as a uniform value (see uniform_u32).
__device__ __forceinline__ unsigned uniform_from_lane(unsigned v, int lane, int j)
{
    return __reduce_or_sync(0xffffffffu, lane == j ? v : 0u);
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Integrates one ray's frequency bins over its visited records.  Lane `lane` holds bins
// k = kbase + lane + 32*q, q < KS; lanes past the last bin recompute bin K-1 (their results are
// never stored), so the inner loop carries no lane predicate and the warp votes are exact.
// Returns the failure code of the ray (0, 2 = negative, 3 = NaN), warp-uniform.
template <int KS>
__device__ __forceinline__ int integrate_ray(const DevProblem &P, const SegRec *seg, unsigned meta,
                                             int lane, int kbase, double (&Iv)[KS],
                                             const double *T)
{
    const int lo = meta & 0xfff, hi = (meta >> 12) & 0xfff;
    const int K = P.K;
    const ArrayConsts KC{ P.kfp, T };
    int koff[KS];
#pragma unroll
    for (int q = 0; q < KS; q++)
        koff[q] = min(kbase + lane + 32 * q, K - 1);
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
                const float *row = gvp + (size_t) rv.z * K;
                float g[KS];
#pragma unroll
                for (int q = 0; q < KS; q++)
                    g[q] = __ldg(row + koff[q]);
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    const float glf = __fmul_rn(gvl, g[q]);
                    const float elf = __fmul_rn(evl, g[q]);
                    const float ag = fabsf(glf);
                    const bool small = ag < 1e-3f; // == (fabs((double) glf) < 1e-3)
                    const unsigned b_small = __ballot_sync(0xffffffffu, small);
                    const unsigned b_odd = __ballot_sync(0xffffffffu, !(ag < 700.0f));
                    const double gl = (double) glf, el = (double) elf;
                    if (b_odd != 0u) { // some lane has |gl| >= 700, inf or NaN: library semantics
                        Iv[q] = ase_update_library(Iv[q], gl, el);
                    } else {
                        double a = 0.0, b = 0.0;
                        if (b_small != 0u) // warp-uniform: some lane takes the Taylor branch
                            a = ase_update_small(Iv[q], gl, el, KC);
                        if (b_small != 0xffffffffu) // warp-uniform: some lane takes the exp branch
                            b = ase_update_large(Iv[q], gl, el, rcp_approx(glf), KC);
                        Iv[q] = small ? a : b;
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
                const float *row = gvp + (size_t) rv.z * K;
#pragma unroll
                for (int q = 0; q < KS; q++)
                    gl[q] = __dadd_rn(gl[q], __dmul_rn(gvl, (double) __ldg(row + koff[q])));
            }
        }
#pragma unroll
        for (int q = 0; q < KS; q++)
            Iv[q] *= exp_any(gl[q], KC);
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

// Hot-path ray integrator of the owner kernel (ASE mode, one pass over K <= 32*KS bins).
//  * the ray's records are fetched with ONE coalesced 16-byte load per 32 records (lane j holds
//    record c0 + j) and handed out by warp shuffles, instead of one dependent global load per
//    record;
//  * the lineshape row of record j+1 is requested before record j is integrated;
//  * the gv base pointers of the planes come from a shared-memory table.
// Constants of the update pinned in registers for the lifetime of the kernel: they are read
// once from the staged blob with VOLATILE loads, so neither the compiler nor ptxas can
// rematerialise them inside the loop as pairs of 32-bit immediates or constant-bank reloads
// (measured: ~30 of 150 instructions per record were such reloads, profiles/r01).  The exp
// table is addressed by an explicit shared-memory load.
struct PinnedConsts {
    double l2e_, nln2_, c1_, c3_, c4_, third_;
    unsigned tab_; // shared-space address of the 2^(j/128) table
    __device__ __forceinline__ static double pin(const double *p)
    {
        double x;
        asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(x) : "l"(p));
        return x;
    }
    __device__ __forceinline__ PinnedConsts(const double *kc, const double *T)
    {
        l2e_ = pin(kc + RTB_K_N_OVER_LN2);
        nln2_ = pin(kc + RTB_K_LN2_OVER_N);
        c1_ = pin(kc + RTB_K_C1);
        c3_ = pin(kc + RTB_K_C3);
        c4_ = pin(kc + RTB_K_C4);
        third_ = pin(kc + RTB_K_THIRD);
        // the shuffle makes the shared-memory address opaque: otherwise the compiler rebuilds it
        // from the CTA's shared window (4 uniform-datapath instructions) at every use
        tab_ = __shfl_sync(0xffffffffu, (unsigned) __cvta_generic_to_shared(T), 0);
    }
    __device__ __forceinline__ double l2e() const { return l2e_; }
    __device__ __forceinline__ double nln2() const { return nln2_; }
    __device__ __forceinline__ double c1() const { return c1_; }
    __device__ __forceinline__ double c3() const { return c3_; }
    __device__ __forceinline__ double c4() const { return c4_; }
    __device__ __forceinline__ double third() const { return third_; }
    __device__ __forceinline__ double tab(int j) const
    {
        double v;
        asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(tab_ + ((unsigned) j << 3)));
        return v;
    }
};

// Loads of the hand-off (records, meta words) in the owner kernels.  The data is read exactly once,
// so it need not be kept in L1 next to the lineshape rows; and when the kernel overlaps the march
// (Outputs::pix_done) it MUST not be: a line that straddles two pixels could have been brought into
// the SM's L1 by the first pixel's CTA before the march had written the second pixel's part.
#ifdef RTB_HANDOFF_NC
#define RTB_HANDOFF_LD(p) __ldg(p)
#else
#define RTB_HANDOFF_LD(p) __ldcg(p)
#endif

template <int KS>
__device__ __forceinline__ int integrate_ray_ase_fast(const DevProblem &P, unsigned sgv_shared,
                                                      const SegRec *seg, unsigned meta, int lane,
                                                      const int (&koff)[KS], double (&Iv)[KS],
                                                      const PinnedConsts &KC, unsigned slab)
{
    // sgv_shared: shared-space address of the CTA's table of lineshape base pointers ([N] x 64 bit)
    // slab: shared-space address of this warp's 32 x 16-byte record slab (opaque register)
    const int lo = meta & 0xfff, hi = (meta >> 12) & 0xfff;
    const int K = P.K;
    auto slab_load = [slab](int j) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(slab + 16u * (unsigned) j));
        return v;
    };
    for (int c0 = lo; c0 < hi; c0 += 32) {
        const int cnt = min(32, hi - c0);
        // Lane j fetches record c0 + j (one coalesced 16-byte load per lane) and resolves the
        // address of its lineshape row; the warp then walks the records through its
        // shared-memory slab: one 16-byte broadcast read per record replaces three shuffles,
        // the plane look-up and the 64-bit address arithmetic.
        __syncwarp();
        unsigned gvl_abs = 0u;
        bool nonzero = false;
        if (lane < cnt) {
            const int4 rv = RTB_HANDOFF_LD(reinterpret_cast<const int4 *>(&seg[c0 + lane]));
            // (the table of lineshape base pointers through its shared-space address, kept opaque in
            // a uniform register: otherwise it is rebuilt from the CTA's shared window at every use)
            unsigned long long gvp;
            asm volatile("ld.shared.u64 %0, [%1];"
                         : "=l"(gvp)
                         : "r"(sgv_shared + 8u * (unsigned) ((c0 + lane) / RTB_N_SUB + 1)));
            const float *row = reinterpret_cast<const float *>(gvp) + (size_t) rv.z * K;
            const unsigned long long ra = reinterpret_cast<unsigned long long>(row);
            gvl_abs = (unsigned) rv.x & 0x7fffffffu;
            // gvl == 0 && evl == 0 (either sign of zero): gl = el = 0, the update is the identity
            nonzero = (((unsigned) rv.x | (unsigned) rv.y) & 0x7fffffffu) != 0u;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slab + 16u * (unsigned) lane),
                         "r"((unsigned) rv.x), "r"((unsigned) rv.y), "r"((unsigned) ra),
                         "r"((unsigned) (ra >> 32))
                         : "memory");
        }
        __syncwarp();
        // The records that change anything, as a warp-uniform bit mask: the walk below visits
        // exactly these (30 % of the records of ASE_medium-synth lie outside the plasma), and
        // everything that steers it - mask, loop counters, vote results - is uniform by
        // construction, so no branch of the walk needs a divergence check or a reconvergence point.
        unsigned todo = __ballot_sync(0xffffffffu, nonzero);
        if (todo == 0u)
            continue;
        // exp-range test of the records (|gl| >= 700, inf, NaN -> library exp): needed for none
        // of these records if max|gvl| * max|gv| stays below 700 (bit patterns order like the
        // magnitudes, NaN above everything; the comparison is false for NaN)
        const float gvl_max = __uint_as_float(__reduce_max_sync(0xffffffffu, gvl_abs));
        const bool need_range_test =
            !(__fmul_rn(__fmul_rn(gvl_max, __uint_as_float(P.gv_absmax_bits)), 1.000001f) < 700.0f);
        // The update of one record; `g` are the lineshape values of this lane's bins.
        // (generic over a compile-time flag: the walk below is instantiated with and without the
        // range test, so the common walk - no record of the chunk can reach |gl| = 700 - carries
        // neither the test nor its branch)
        auto update = [&](auto range_tag, float gvl, float evl, const float (&g)[KS]) {
            constexpr bool range_test = decltype(range_tag)::value;
            // Branch decisions by warp votes: their results are uniform predicates, so the
            // dispatch below costs a branch each and nothing else.
            float glf[KS], elf[KS];
            bool small[KS];
#pragma unroll
            for (int q = 0; q < KS; q++) {
                glf[q] = __fmul_rn(gvl, g[q]);
                elf[q] = __fmul_rn(evl, g[q]);
                small[q] = fabsf(glf[q]) < 1e-3f; // == (fabs((double) glf) < 1e-3)
            }
            bool out_of_range = false;
            if (range_test) { // warp-uniform; one test for all slots: sum of |gl| (NaN, inf propagate)
                float ag_sum = 0.0f;
#pragma unroll
                for (int q = 0; q < KS; q++)
                    ag_sum += fabsf(glf[q]);
                out_of_range = __any_sync(0xffffffffu, !(ag_sum < 700.0f));
            }
            if (out_of_range) { // conservative: the library path is always valid
#pragma unroll
                for (int q = 0; q < KS; q++)
                    Iv[q] = ase_update_library(Iv[q], (double) glf[q], (double) elf[q]);
                return;
            }
#pragma unroll
            for (int q = 0; q < KS; q++) {
                const double gl = (double) glf[q], el = (double) elf[q];
                // three straight-line variants: the common one (every lane of the slot on the
                // exp branch) carries no select and no dead Taylor result
                if (__builtin_expect(!__any_sync(0xffffffffu, small[q]), 1)) {
                    Iv[q] = ase_update_large(Iv[q], gl, el, rcp_approx(glf[q]), KC);
                } else if (__builtin_expect(__all_sync(0xffffffffu, small[q]), 0)) {
                    Iv[q] = ase_update_small(Iv[q], gl, el, KC);
                } else {
                    const double a = ase_update_small(Iv[q], gl, el, KC);
                    const double b = ase_update_large(Iv[q], gl, el, rcp_approx(glf[q]), KC);
                    Iv[q] = small[q] ? a : b;
                }
            }
        };
        auto fetch = [&](int j, uint4 &e, float (&g)[KS]) {
            e = slab_load(j);
            const float *row = reinterpret_cast<const float *>(((unsigned long long) e.w << 32) | e.z);
#pragma unroll
            for (int q = 0; q < KS; q++)
                g[q] = __ldg(row + koff[q]);
        };
        // next record of the mask (the mask is not empty)
        auto pop = [&todo]() {
            const int j = __ffs((int) todo) - 1;
            todo &= todo - 1u;
            return j;
        };
        // Two records per trip with ping-pong registers: the row of the next record is
        // requested before the current one is integrated, without rotating registers.  Past the
        // last record the prefetch re-reads it (unused), which costs less than a guarded fetch.
        auto walk = [&](auto range_tag) {
            uint4 eA, eB;
            float gA[KS], gB[KS];
            int j = pop();
            fetch(j, eA, gA);
            for (;;) {
                const bool lastA = todo == 0u;
                if (!lastA)
                    j = pop();
                fetch(j, eB, gB);
                update(range_tag, __uint_as_float(eA.x), __uint_as_float(eA.y), gA);
                if (lastA)
                    break;
                const bool lastB = todo == 0u;
                if (!lastB)
                    j = pop();
                fetch(j, eA, gA);
                update(range_tag, __uint_as_float(eB.x), __uint_as_float(eB.y), gB);
                if (lastB)
                    break;
            }
        };
        if (__builtin_expect(need_range_test, 0))
            walk(std::true_type{});
        else
            walk(std::false_type{});
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

// Gain-only ray integrator of the seeded path (RayTraceImageHelper.h:569-581):
//     Iv[k] *= exp( sum over records of (double) gvl * (double) gv[cell][k] ).
// Lane j fetches record c0 + j (one coalesced 16-byte load), widens gvl and resolves the address
// of its lineshape row, and parks both in the warp's shared-memory slab; the warp then walks
// the slab with one 16-byte broadcast read per record.  The product of two floats is exact in
// double (48 significant bits, no underflow), so RN(gl + RN(gvl*gv)) == fma(gvl, gv, gl): one
// DFMA per bin and record, bit-identical to the reference's multiply-then-add.
template <int KS>
__device__ __forceinline__ int integrate_ray_gain_fast(const DevProblem &P, const float *const *s_gv,
                                                       const SegRec *seg, unsigned meta, int lane,
                                                       int kbase, double (&Iv)[KS],
                                                       const ArrayConsts &KC, unsigned slab)
{
    const int lo = meta & 0xfff, hi = (meta >> 12) & 0xfff;
    const int K = P.K;
    int koff[KS];
    double gl[KS];
#pragma unroll
    for (int q = 0; q < KS; q++) {
        koff[q] = min(kbase + lane + 32 * q, K - 1);
        gl[q] = 0.0;
    }
    for (int c0 = lo; c0 < hi; c0 += 32) {
        const int cnt = min(32, hi - c0);
        __syncwarp();
        if (lane < cnt) {
            const int4 rv = __ldg(reinterpret_cast<const int4 *>(&seg[c0 + lane]));
            const float *row = s_gv[(c0 + lane) / RTB_N_SUB + 1] + (size_t) rv.z * K;
            const unsigned long long ra = reinterpret_cast<unsigned long long>(row);
            const unsigned long long gd =
                (unsigned long long) __double_as_longlong((double) __int_as_float(rv.x));
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slab + 16u * (unsigned) lane),
                         "r"((unsigned) gd), "r"((unsigned) (gd >> 32)), "r"((unsigned) ra),
                         "r"((unsigned) (ra >> 32))
                         : "memory");
        }
        __syncwarp();
        for (int j = 0; j < cnt; j++) {
            uint4 e;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                         : "r"(slab + 16u * (unsigned) j));
            // a record with gvl == 0 adds exactly +0 to every bin: fma(0, gv, gl) == gl
            const double gvl = __longlong_as_double((long long) (((unsigned long long) e.y << 32) | e.x));
            const float *row = reinterpret_cast<const float *>(((unsigned long long) e.w << 32) | e.z);
#pragma unroll
            for (int q = 0; q < KS; q++)
                gl[q] = __fma_rn(gvl, (double) __ldg(row + koff[q]), gl[q]);
        }
    }
    bool neg = false, nan = false;
#pragma unroll
    for (int q = 0; q < KS; q++) {
        Iv[q] *= exp_any(gl[q], KC);
        neg = neg || Iv[q] < 0.0;
        nan = nan || Iv[q] != Iv[q];
    }
    const bool any_neg = __any_sync(0xffffffffu, neg);
    const bool any_nan = __any_sync(0xffffffffu, nan);
    return any_neg ? 2 : (any_nan ? 3 : 0);
}

// The same integrator for records that are ALREADY parked in the warp's slab (entries lo..hi-1 of
// this ray: {(double) gvl, lineshape row address}): the scatter kernel fetches the records of a
// whole batch of rays with coalesced loads before it walks the batch, so the DRAM latency of the
// hand-off records is paid once per batch instead of once per ray.
#define RTB_SLAB_RECORDS 256 // per warp
template <int KS>
__device__ __forceinline__ int integrate_ray_gain_slab(const DevProblem &P, unsigned meta, int lane, int kbase,
                                                       double (&Iv)[KS], const ArrayConsts &KC, unsigned slab)
{
    const int lo = meta & 0xfff, hi = (meta >> 12) & 0xfff;
    const int K = P.K;
    int koff[KS];
    double gl[KS];
#pragma unroll
    for (int q = 0; q < KS; q++) {
        koff[q] = min(kbase + lane + 32 * q, K - 1);
        gl[q] = 0.0;
    }
    for (int j = lo; j < hi; j++) {
        uint4 e;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                     : "r"(slab + 16u * (unsigned) j));
        // a record with gvl == 0 adds exactly +0 to every bin: fma(0, gv, gl) == gl
        const double gvl = __longlong_as_double((long long) (((unsigned long long) e.y << 32) | e.x));
        const float *row = reinterpret_cast<const float *>(((unsigned long long) e.w << 32) | e.z);
#pragma unroll
        for (int q = 0; q < KS; q++)
            gl[q] = __fma_rn(gvl, (double) __ldg(row + koff[q]), gl[q]);
    }
    bool neg = false, nan = false;
#pragma unroll
    for (int q = 0; q < KS; q++) {
        Iv[q] *= exp_any(gl[q], KC);
        neg = neg || Iv[q] < 0.0;
        nan = nan || Iv[q] != Iv[q];
    }
    const bool any_neg = __any_sync(0xffffffffu, neg);
    const bool any_nan = __any_sync(0xffffffffu, nan);
    return any_neg ? 2 : (any_nan ? 3 : 0);
}

// Overlapped launch (Outputs::pix_done): the CTA waits until the march has closed every ray slot
// of its pixel.  One thread polls with acquire semantics; the march counted with release
// semantics after the slot's records and meta word.  (The march's CTAs are all resident or done
// when this kernel is allowed to start, so the wait cannot starve them.)
__device__ __forceinline__ void owner_wait_for_pixel(const Outputs &o, unsigned q, unsigned expected)
{
    if (o.pix_done == nullptr)
        return;
    if (threadIdx.x == 0) {
        unsigned v;
        for (;;) { // relaxed polling (an acquire load would invalidate the SM's L1 at every poll)
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(o.pix_done + q) : "memory");
            if (v >= expected)
                break;
            __nanosleep(200);
        }
        if (o.gv_flag != nullptr) { // ... and until the lineshape tables of this image have arrived
            for (;;) {
                asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(o.gv_flag) : "memory");
                if (v == o.gv_epoch)
                    break;
                __nanosleep(500);
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory"); // (the writer is a copy engine)
        } else {
            asm volatile("fence.acq_rel.gpu;" ::: "memory"); // once per CTA
        }
    }
    __syncthreads();
}

// What the owner kernels address per ray, as uniform values that are built once per CTA: the
// pixel's first hand-off record (a ray's records are then one 32-bit multiply-add away; left
// alone, the 64-bit product slot * S is rebuilt from the kernel parameters for every ray) and the
// shared-space address of the table of lineshape base pointers.
struct OwnerBase {
    unsigned seg_lo, seg_hi, ray_bytes, sgv_shared;
    __device__ __forceinline__ OwnerBase(const SegRec *seg_pix, int S, const float *const *s_gv)
    {
        const unsigned long long a = reinterpret_cast<unsigned long long>(seg_pix);
        seg_lo = uniform_u32((unsigned) a);
        seg_hi = uniform_u32((unsigned) (a >> 32));
        ray_bytes = (unsigned) S * (unsigned) sizeof(SegRec);
        sgv_shared = uniform_u32((unsigned) __cvta_generic_to_shared(s_gv));
    }
    // records of the pixel's ray t (a pixel's records fit 32 bits of bytes: the host sizes chunks so
    // that slots * S of a whole chunk stay below 2^31)
    __device__ __forceinline__ const SegRec *ray(int t) const
    {
        const unsigned long long a =
            (((unsigned long long) seg_hi << 32) | seg_lo) + (unsigned long long) ((unsigned) t * ray_bytes);
        return reinterpret_cast<const SegRec *>(a);
    }
};

#define RTB_OWNER_WARPS 8
// Resident CTAs per SM the compiler budgets registers for: up to two lane slots (K <= 64) the
// kernel fits 48 registers with a handful of spills outside the walk, and 40 warps per SM hide
// more of the FP64 latency than 32 (-2.4 % on the headline); with more slots it keeps 64.
#ifndef RTB_OWNER_MINBLOCKS
#define RTB_OWNER_MINBLOCKS 4
#endif
#ifndef RTB_OWNER_MINBLOCKS_SMALL
#define RTB_OWNER_MINBLOCKS_SMALL 5
#endif

template <int KS>
__global__ void __launch_bounds__(RTB_OWNER_WARPS * 32, KS <= 2 ? RTB_OWNER_MINBLOCKS_SMALL : RTB_OWNER_MINBLOCKS)
    integrate_ase_owner_kernel(const DevProblem P, const Chunk c, const Handoff h, const Outputs o)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double part[RTB_OWNER_WARPS][KS * 32];
    __shared__ double exp_tab[RTB_EXP_TABLE_SIZE];
    __shared__ uint4 rec_slab[RTB_OWNER_WARPS][32]; // per-warp record + row-address slab
    const float **s_gv = reinterpret_cast<const float **>(smem_raw); // [N] gv base pointers
    for (int i = threadIdx.x; i < P.N; i += blockDim.x)
        s_gv[i] = P.planes[i].gv;
    load_exp_table(exp_tab); // includes __syncthreads()
    const int lane = threadIdx.x & 31, warp = (int) uniform_u32(threadIdx.x >> 5);
    owner_wait_for_pixel(o, blockIdx.x, (unsigned) P.ab_max);
    const long long p = phys_pixel(P, c, c.pix0 + blockIdx.x);
    const PixelRays pr = pixel_rays(P, p);
    const int S = (P.N - 1) * RTB_N_SUB;
    const int K = P.K;
    const long long slot0 = (long long) blockIdx.x * P.ab_max;
    double pix[KS], dv2[KS];
    int koff[KS];
#pragma unroll
    for (int q = 0; q < KS; q++) {
        pix[q] = 0.0;
        const int k = lane + 32 * q;
        dv2[q] = k < K ? __ldg(&P.dv2[k]) : 0.0;
        // Lanes past the last bin recompute a bin that another lane of the SAME slot already
        // holds (never stored): duplicates cannot add a branch to the slot's flag set, whereas
        // a fixed bin K-1 (far wing, Taylor branch) made the last slot take both branches.
        const int live = K - 32 * q; // bins of this slot
        koff[q] = k < K ? k : (live > 0 ? 32 * q + (k - K) % live : K - 1);
    }
    const PinnedConsts KC(P.kfp_g, exp_tab);
    const unsigned slab_addr = uniform_u32((unsigned) __cvta_generic_to_shared(&rec_slab[warp][0]));
    const OwnerBase ob(h.seg + slot0 * S, S, s_gv);
    // The warp's rays are t = warp, warp + 8, ...; what is per ray and not per bin (hand-off meta
    // word, angular bin) is looked up by one lane per ray, 32 rays at a time.
    for (int t0 = warp; t0 < pr.cnt; t0 += 32 * RTB_OWNER_WARPS) {
        unsigned meta_l = RTB_META_INVALID;
        int bin_l = -1;
        {
            const int t = t0 + lane * RTB_OWNER_WARPS;
            if (t < pr.cnt) {
                meta_l = RTB_HANDOFF_LD(&h.meta[slot0 + t]);
                const int ab = pr.ab0 + t * (int) P.n_parallel;
                const int ka = ab / P.snb, m = ab - ka * P.snb;
                const int ba = __ldg(&P.binA[ka]), bb = __ldg(&P.binB[m]);
                bin_l = (ba >= 0 && bb >= 0) ? ba + bb * P.na : -1;
            }
        }
        const int n_here = min(32, (pr.cnt - t0 + RTB_OWNER_WARPS - 1) / RTB_OWNER_WARPS);
        for (int j = 0; j < n_here; j++) {
            const int t = t0 + j * RTB_OWNER_WARPS;
            const unsigned meta = uniform_from_lane(meta_l, lane, j);
            const int bin = (int) uniform_from_lane((unsigned) bin_l, lane, j);
            if (meta & RTB_META_INVALID)
                continue; // error -1, reported by the march
            double Iv[KS];
#pragma unroll
            for (int q = 0; q < KS; q++)
                Iv[q] = 0.0;
            const int code = integrate_ray_ase_fast<KS>(P, ob.sgv_shared, ob.ray(t), meta, lane, koff, Iv, KC,
                                                        slab_addr);
            if (code != 0) {
                if (lane == 0) {
                    const int ab = pr.ab0 + t * (int) P.n_parallel;
                    const int ka = ab / P.snb, m = ab % P.snb;
                    report_failure(o.fail, code, P.sxf[pr.i], P.syf[pr.j], P.saf[ka], P.sbf[m]);
                }
                continue;
            }
            double w = 0.0;
#pragma unroll
            for (int q = 0; q < KS; q++) {
                w += dv2[q] * Iv[q];
                pix[q] += Iv[q] * P.scale;
            }
            w = warp_sum(w);
            if (lane == 0 && bin >= 0)
                atomicAdd(&o.I_ang[bin], w);
        }
    }
#pragma unroll
    for (int q = 0; q < KS; q++)
        part[warp][q * 32 + lane] = pix[q];
    __syncthreads();
    const int pi = __ldg(&P.pixI[pr.i]), pj = __ldg(&P.pixJ[pr.j]);
    if (!o.compact && (pi < 0 || pj < 0))
        return;
    // destination pixel, or (compact) the launch's logical pixel
    const size_t opix = o.compact ? (size_t) (c.pix0 + blockIdx.x) : (size_t) pi + (size_t) pj * P.nx;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < RTB_OWNER_WARPS; w++)
            sum += part[w][k];
        o.image[(size_t) K * opix + k] = sum;
    }
}

// The same kernel for more than 128 frequency bins (spectral sweeps): the bins are covered in
// tiles of 128 (four lane slots); per tile every ray of the pixel is integrated over the tile's
// bins and the pixel's tile of the spectrum is written.  A ray's records are re-read once per
// tile (they are small next to the lineshape rows, which are read exactly once).  I_ang gets one
// atomic per (ray, tile).  A ray that fails (negative / NaN intensity) is reported by every tile
// that sees it; the reference aborts on the first failure anyway (src/RayTraceImage.cpp:427-430).
__global__ void __launch_bounds__(RTB_OWNER_WARPS * 32, 3)
    integrate_ase_owner_tiled_kernel(const DevProblem P, const Chunk c, const Handoff h, const Outputs o)
{
    constexpr int KS = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double part[RTB_OWNER_WARPS][KS * 32];
    __shared__ double exp_tab[RTB_EXP_TABLE_SIZE];
    __shared__ uint4 rec_slab[RTB_OWNER_WARPS][32];
    const float **s_gv = reinterpret_cast<const float **>(smem_raw); // [N] gv base pointers
    for (int i = threadIdx.x; i < P.N; i += blockDim.x)
        s_gv[i] = P.planes[i].gv;
    load_exp_table(exp_tab); // includes __syncthreads()
    const int lane = threadIdx.x & 31, warp = (int) uniform_u32(threadIdx.x >> 5);
    owner_wait_for_pixel(o, blockIdx.x, (unsigned) P.ab_max);
    const long long p = phys_pixel(P, c, c.pix0 + blockIdx.x);
    const PixelRays pr = pixel_rays(P, p);
    const int S = (P.N - 1) * RTB_N_SUB;
    const int K = P.K;
    const long long slot0 = (long long) blockIdx.x * P.ab_max;
    const PinnedConsts KC(P.kfp_g, exp_tab);
    const unsigned slab_addr = uniform_u32((unsigned) __cvta_generic_to_shared(&rec_slab[warp][0]));
    const OwnerBase ob(h.seg + slot0 * S, S, s_gv);
    const int pi = __ldg(&P.pixI[pr.i]), pj = __ldg(&P.pixJ[pr.j]);
    const bool store = o.compact || (pi >= 0 && pj >= 0);
    const size_t opix = o.compact ? (size_t) (c.pix0 + blockIdx.x) : (size_t) pi + (size_t) pj * P.nx;
    for (int kbase = 0; kbase < K; kbase += 32 * KS) {
        double pix[KS], dv2[KS];
        int koff[KS];
#pragma unroll
        for (int q = 0; q < KS; q++) {
            pix[q] = 0.0;
            const int k = kbase + lane + 32 * q;
            dv2[q] = k < K ? __ldg(&P.dv2[k]) : 0.0;
            const int live = K - kbase - 32 * q; // bins of this slot (see the one-pass kernel)
            koff[q] = k < K ? k : (live > 0 ? kbase + 32 * q + (k - K) % live : K - 1);
        }
        for (int t0 = warp; t0 < pr.cnt; t0 += 32 * RTB_OWNER_WARPS) {
            unsigned meta_l = RTB_META_INVALID;
            int bin_l = -1;
            {
                const int t = t0 + lane * RTB_OWNER_WARPS;
                if (t < pr.cnt) {
                    meta_l = RTB_HANDOFF_LD(&h.meta[slot0 + t]);
                    const int ab = pr.ab0 + t * (int) P.n_parallel;
                    const int ka = ab / P.snb, m = ab - ka * P.snb;
                    const int ba = __ldg(&P.binA[ka]), bb = __ldg(&P.binB[m]);
                    bin_l = (ba >= 0 && bb >= 0) ? ba + bb * P.na : -1;
                }
            }
            const int n_here = min(32, (pr.cnt - t0 + RTB_OWNER_WARPS - 1) / RTB_OWNER_WARPS);
            for (int j = 0; j < n_here; j++) {
                const int t = t0 + j * RTB_OWNER_WARPS;
                const unsigned meta = uniform_from_lane(meta_l, lane, j);
                const int bin = (int) uniform_from_lane((unsigned) bin_l, lane, j);
                if (meta & RTB_META_INVALID)
                    continue; // error -1, reported by the march
                double Iv[KS];
#pragma unroll
                for (int q = 0; q < KS; q++)
                    Iv[q] = 0.0;
                const int code = integrate_ray_ase_fast<KS>(P, ob.sgv_shared, ob.ray(t), meta, lane, koff, Iv, KC,
                                                            slab_addr);
                if (code != 0) {
                    if (lane == 0) {
                        const int ab = pr.ab0 + t * (int) P.n_parallel;
                        const int ka = ab / P.snb, m = ab % P.snb;
                        report_failure(o.fail, code, P.sxf[pr.i], P.syf[pr.j], P.saf[ka], P.sbf[m]);
                    }
                    continue;
                }
                double w = 0.0;
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    w += dv2[q] * Iv[q];
                    pix[q] += Iv[q] * P.scale;
                }
                w = warp_sum(w);
                if (lane == 0 && bin >= 0)
                    atomicAdd(&o.I_ang[bin], w);
            }
        }
        __syncthreads(); // the previous tile's partials have been read
#pragma unroll
        for (int q = 0; q < KS; q++)
            part[warp][q * 32 + lane] = pix[q];
        __syncthreads();
        if (store) {
            for (int k = threadIdx.x; k < 32 * KS && kbase + k < K; k += blockDim.x) {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < RTB_OWNER_WARPS; w++)
                    sum += part[w][k];
                o.image[(size_t) K * opix + kbase + k] = sum;
            }
        }
    }
}

// One CTA per source row j: row j was traced by device j % world as its compact row j / world.
__global__ void __launch_bounds__(256) unpermute_rows_kernel(const DevProblem P, const double *gathered, int world,
                                                             long long rows_per_dev, double *image)
{
    const int j = blockIdx.x;
    const int pj = __ldg(&P.pixJ[j]);
    if (pj < 0)
        return;
    const int K = P.K;
    const double *src = gathered + ((size_t) (j % world) * (size_t) rows_per_dev + (size_t) (j / world)) * (size_t) P.snx * K;
    const int n = P.snx * K;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int i = e / K, k = e - i * K;
        const int pi = __ldg(&P.pixI[i]);
        if (pi >= 0)
            image[(size_t) K * ((size_t) pi + (size_t) pj * P.nx) + k] = src[e];
    }
}

void launch_unpermute_rows(const DevProblem &P, const double *gathered, int world, long long rows_per_dev,
                           double *image, cudaStream_t st)
{
    if (P.sny > 0)
        unpermute_rows_kernel<<<(unsigned) P.sny, 256, 0, st>>>(P, gathered, world, rows_per_dev, image);
}

template <class Kern>
static void launch_owner(Kern kern, unsigned blocks, int threads, size_t smem, cudaStream_t st, bool overlap,
                         const DevProblem &P, const Chunk &c, const Handoff &h, const Outputs &o)
{
    if (!overlap) {
        kern<<<blocks, threads, smem, st>>>(P, c, h, o);
        return;
    }
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3((unsigned) threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, P, c, h, o);
}

void launch_integrate_ase_owner(const DevProblem &P, const Chunk &c, const Handoff &h,
                                const Outputs &o, cudaStream_t st, bool overlap)
{
    const long long npix = c.pix1 - c.pix0;
    if (npix <= 0)
        return;
    const unsigned blocks = (unsigned) npix;
    const int threads = RTB_OWNER_WARPS * 32;
    const int ks = (P.K + 31) / 32;
    const size_t smem = sizeof(float *) * (size_t) P.N;
    switch (ks) {
    case 1: launch_owner(integrate_ase_owner_kernel<1>, blocks, threads, smem, st, overlap, P, c, h, o); break;
    case 2: launch_owner(integrate_ase_owner_kernel<2>, blocks, threads, smem, st, overlap, P, c, h, o); break;
    case 3: launch_owner(integrate_ase_owner_kernel<3>, blocks, threads, smem, st, overlap, P, c, h, o); break;
    case 4: launch_owner(integrate_ase_owner_kernel<4>, blocks, threads, smem, st, overlap, P, c, h, o); break;
    default: launch_owner(integrate_ase_owner_tiled_kernel, blocks, threads, smem, st, overlap, P, c, h, o); break; // K > 128
    }
}

// getIndex (RayTraceImageCPU.cpp:11-16) on the device, for the exit ray: -1 outside the grid
// +- half a cell, else findfirstsingle(x, n, y - dx/2) = the first index with x[idx] >= Y
// (0 below the grid, n above it).  The euv grids are uniform (validated), so the index is
// guessed in closed form and then fixed up against the real coordinates: the result is the
// reference's bisection result for any monotone grid.
__device__ __forceinline__ int dev_get_index(int n, const double *x, double dx, double y)
{
    const double x0 = __ldg(&x[0]), xn = __ldg(&x[n - 1]);
    if (y < x0 - 0.5 * dx || y > xn + 0.5 * dx)
        return -1;
    const double Y = y - 0.5 * dx;
    if (Y < x0)
        return 0;
    if (Y > xn)
        return n;
    if (!(Y > x0)) // Y == x[0]: the bisection never returns 0 for Y >= x[0] (hi stops at 1)
        return n > 1 ? 1 : 0;
    int k = (int) ceil((Y - x0) * (1.0 / dx)); // guess (1/dx: dx is warp-uniform per grid)
    k = k < 1 ? 1 : (k > n - 1 ? n - 1 : k);
    while (k > 1 && __ldg(&x[k - 1]) >= Y)
        --k;
    while (k < n - 1 && !(__ldg(&x[k]) >= Y))
        ++k;
    return k;
}

// findfirstsingle / interp_pchip / calc_seed_inline (RayTraceImageHelper.h:101-117, :168-247)
// on the device, for explicit ray lists with a seed beam.  The seed amplitude multiplies the
// spectrum (it feeds no discrete decision), so ordinary FP64 arithmetic is sufficient.
__device__ int dev_findfirstsingle(const double *X, int n, double Y)
{
    if (Y < __ldg(&X[0]))
        return 0;
    if (Y > __ldg(&X[n - 1]))
        return n;
    int lo = 0, hi = n - 1;
    while (hi - lo != 1) {
        const int mid = (hi + lo) / 2;
        if (__ldg(&X[mid]) >= Y)
            hi = mid;
        else
            lo = mid;
    }
    return hi;
}

__device__ double dev_interp_pchip(int N, const double *xi, const double *yi, double x)
{
    if (x <= xi[0] || N <= 2) {
        const double t = (x - xi[0]) / (xi[1] - xi[0]);
        return (1.0 - t) * yi[0] + t * yi[1];
    }
    if (x >= xi[N - 1]) {
        const double t = (x - xi[N - 2]) / (xi[N - 1] - xi[N - 2]);
        return (1.0 - t) * yi[N - 2] + t * yi[N - 1];
    }
    const int i = dev_findfirstsingle(xi, N, x);
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
        const double s1 = fabs(f1 - f0) / h1, s2 = fabs(f2 - f1) / h2;
        const double g_max = 2 * h2 * (s1 < s2 ? s1 : s2);
        g1 = ((g1 >= 0) ? 1 : -1) * (fabs(g1) < g_max ? fabs(g1) : g_max);
    }
    if (i >= N - 1) {
        g2 = f2 - f1;
    } else if ((f2 < f1 && f2 > yi[i + 1]) || (f2 > f1 && f2 < yi[i + 1])) {
        const double f0 = yi[i + 1];
        const double h1 = xi[i] - xi[i - 1], h2 = xi[i + 1] - xi[i];
        const double a1 = -h2 / (h1 + h2), a2 = (h2 - h1) / h2;
        g2 = a1 * (f1 - f0) + a2 * (f2 - f0);
        const double s1 = fabs(f2 - f1) / h1, s2 = fabs(f0 - f2) / h2;
        const double g_max = 2 * h1 * (s1 < s2 ? s1 : s2);
        g2 = ((g2 >= 0) ? 1 : -1) * (fabs(g2) < g_max ? fabs(g2) : g_max);
    }
    const double t2 = t * t;
    return f1 + t2 * (2 * t - 3) * (f1 - f2) + t * g1 - t2 * (g1 + (1 - t) * (g1 + g2));
}

__device__ double dev_calc_seed(const DevProblem &P, double x, double y, double a, double b)
{
    const double v[4] = { x, y, a, b };
    double f = P.seed_f0;
    for (int d = 0; d < 4; d++) {
        const int n = P.sd_dim[d];
        if (!(v[d] >= P.sd_x[d][0] && v[d] <= P.sd_x[d][n - 1]))
            return 0.0;
    }
    for (int d = 0; d < 4; d++)
        f *= dev_interp_pchip(P.sd_dim[d], P.sd_x[d], P.sd_f[d], v[d]);
    return f < 0.0 ? 0.0 : f;
}

// One warp per ray slot; K is covered in passes of 32*KS bins (one pass for K <= 128).  Handles both integration modes,
// both ray sources, scatter binning and the per-ray dumps.
#ifndef RTB_SCATTER_MINBLOCKS
#define RTB_SCATTER_MINBLOCKS 4
#endif
template <bool LIST, int KS>
__global__ void __launch_bounds__(256, RTB_SCATTER_MINBLOCKS)
    integrate_scatter_kernel(const DevProblem P, const Chunk c, const Handoff h, const Outputs o)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double exp_tab[RTB_EXP_TABLE_SIZE];
    // dynamic shared memory: per-warp record + row-address slabs (8 x RTB_SLAB_RECORDS x 16 B),
    // then the gv base pointers of the planes
    uint4 *rec_slab = reinterpret_cast<uint4 *>(smem_raw);
    const float **s_gv = reinterpret_cast<const float **>(smem_raw + 8 * RTB_SLAB_RECORDS * sizeof(uint4)); // [N]
    for (int i = threadIdx.x; i < P.N; i += blockDim.x)
        s_gv[i] = P.planes[i].gv;
    load_exp_table(exp_tab); // includes __syncthreads()
    const ArrayConsts KC{ P.kfp, exp_tab };
    const bool gain_only = P.use_emis == 0;
    const int lane = threadIdx.x & 31;
    const unsigned slab =
        __shfl_sync(0xffffffffu, (unsigned) __cvta_generic_to_shared(rec_slab + (threadIdx.x >> 5) * RTB_SLAB_RECORDS), 0);
    // slot counts of one chunk fit 32 bits (the host sizes chunks that way): 32-bit loop counters
    const int warp_id = (int) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int n_warps = (int) ((gridDim.x * blockDim.x) >> 5);
    const int n_slots = (int) (LIST ? (c.ray1 - c.ray0) : (c.pix1 - c.pix0) * P.ab_max);
    const int S = (P.N - 1) * RTB_N_SUB;
    const int K = P.K;
    // Each warp takes RUN consecutive ray slots at a time.  Consecutive rays (b fastest, then a)
    // leave the plasma next to each other, so they mostly fall into the same image pixel and
    // the same angular bin: their contributions are summed in registers and flushed with one
    // set of FP64 atomics when the destination changes, instead of 1 + K atomics per ray onto
    // addresses that thousands of concurrent rays share (measured: the atomics, not the
    // arithmetic, bound the seeded path, profiles/r01_seed).
    constexpr int RUN = 64;
    double acc[KS];
#pragma unroll
    for (int q = 0; q < KS; q++)
        acc[q] = 0.0;
    double acc_w = 0.0;
    int cur_pix = -1; // destination pixels fit 32 bits (validated on the host)
    int cur_bin = -1;
    const bool combine = K <= 32 * KS; // single pass over the bins
    auto flush_pix = [&]() {
#pragma unroll
        for (int q = 0; q < KS; q++) {
            const int k = lane + 32 * q;
            if (cur_pix >= 0 && k < K)
                atomicAdd(&o.image[(size_t) K * (size_t) cur_pix + k], acc[q]);
            acc[q] = 0.0; // also after a run of rays that left the image (cur_pix < 0): dropped
        }
        cur_pix = -1;
    };
    auto flush_bin = [&]() { // acc_w: this lane's share of the run's angular-bin sum
        if (cur_bin >= 0) {
            const double w = warp_sum(acc_w);
            if (lane == 0)
                atomicAdd(&o.I_ang[cur_bin], w);
        }
        acc_w = 0.0;
        cur_bin = -1;
    };
    // The source coordinates of a slot (only needed again to report a failed ray).
    auto slot_ray = [&](long long slot, float &rx, float &ry, float &ra, float &rb, int &pi, int &pj,
                        int &ka, int &m) {
        if (LIST) {
            const float4 r = __ldg(&c.rays[c.ray0 + slot]);
            rx = r.x, ry = r.y, ra = r.z, rb = r.w;
            pi = pj = ka = m = 0;
        } else {
            const unsigned lq = (unsigned) slot / (unsigned) P.ab_max; // slots fit 32 bits
            const long long p = phys_pixel(P, c, c.pix0 + lq);
            const int t = (int) ((unsigned) slot - lq * (unsigned) P.ab_max);
            const PixelRays pr = pixel_rays(P, p);
            const int ab = pr.ab0 + t * (int) P.n_parallel;
            ka = (int) ((unsigned) ab / (unsigned) P.snb);
            m = ab - ka * P.snb;
            pi = pr.i;
            pj = pr.j;
            rx = __ldg(&P.sxf[pi]);
            ry = __ldg(&P.syf[pj]);
            ra = __ldg(&P.saf[ka]);
            rb = __ldg(&P.sbf[m]);
        }
    };
    // Everything about a ray slot that is the same for all frequency bins: the seed amplitude
    // and the destination pixel / angular bin.  Evaluated by ONE lane per slot, 32 slots at a
    // time, and handed to the warp by shuffles (it used to be repeated by all 32 lanes of the
    // warp for every ray: ~250 of ~800 instructions per ray).
    auto prologue = [&](long long slot, unsigned meta, double &f, int &pix, int &bin) {
        float rx, ry, ra, rb;
        int pi, pj, ka, m;
        slot_ray(slot, rx, ry, ra, rb, pi, pj, ka, m);
        f = 0.0;
        if (LIST) {
            if (P.seed_fv && !(meta & (RTB_META_ESCAPED | RTB_META_INVALID))) {
                if (P.method == 1) { // backward: seed at the exit point (:525-529)
                    const float4 e = h.exit_ray[slot];
                    f = dev_calc_seed(P, (double) e.x, (double) e.y, (double) e.z, (double) e.w);
                } else { // forward: seed at the entry point (:530-533)
                    f = dev_calc_seed(P, (double) rx, (double) ry, (double) ra, (double) rb);
                }
            }
        } else if (P.seed_fx && !(meta & RTB_META_ESCAPED)) {
            // calc_seed_inline (:230-247) from the per-index tables
            const double fx = __ldg(&P.seed_fx[pi]), fy = __ldg(&P.seed_fy[pj]);
            const double fa = __ldg(&P.seed_fa[ka]), fb = __ldg(&P.seed_fb[m]);
            if (fx == fx && fy == fy && fa == fa && fb == fb) {
                f = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(P.seed_f0, fx), fy), fa), fb);
                f = f < 0.0 ? 0.0 : f;
            }
        }
        pix = -1;
        bin = -1;
        if (!(meta & RTB_META_INVALID) && (o.image || o.I_ang)) {
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
            const int i1 = dev_get_index(P.nx, P.ex, P.edx, (double) bx);
            const int i2 = dev_get_index(P.ny, P.ey, P.edy, (double) by);
            const int i3 = dev_get_index(P.na, P.ea, P.eda, (double) ba);
            const int i4 = dev_get_index(P.nb, P.eb, P.edb, (double) bb);
            if (o.image && i1 >= 0 && i2 >= 0)
                pix = i1 + i2 * P.nx;
            if (i3 >= 0 && i4 >= 0)
                bin = i3 + i4 * P.na;
        }
    };
    // Rays whose records are fetched together (gain-only integration): as many as fit the slab.
    int batch = 32;
    while (batch > 1 && batch * S > RTB_SLAB_RECORDS)
        batch >>= 1;
    const bool batched = gain_only && S >= 1 && S <= RTB_SLAB_RECORDS;
    const int n_runs = (n_slots + RUN - 1) / RUN;
    for (int run = warp_id; run < n_runs; run += n_warps) {
    const int slot_end = min((run + 1) * RUN, n_slots);
    for (int base = run * RUN; base < slot_end; base += 32) {
    unsigned meta_l = RTB_META_INACTIVE;
    double f_l = 0.0;
    int pix_l = -1;
    int bin_l = -1;
    if (base + lane < slot_end) {
        meta_l = __ldg(&h.meta[base + lane]);
        if (!(meta_l & RTB_META_INACTIVE))
            prologue(base + lane, meta_l, f_l, pix_l, bin_l);
    }
    __syncwarp();
    const int n_here = min(slot_end - base, 32);
    for (int j = 0; j < n_here; j++) {
        const long long slot = base + j; // 64-bit where it scales addresses
        if (batched && (j & (batch - 1)) == 0) {
            // one coalesced sweep over the records of the next `batch` rays (they are contiguous
            // in the hand-off arena); entries outside a ray's visited range are never read
            __syncwarp();
            const int nrec = min(batch, n_here - j) * S;
            const SegRec *src = h.seg + slot * S;
            for (int r = lane; r < nrec; r += 32) {
                const int4 rv = __ldg(reinterpret_cast<const int4 *>(&src[r]));
                const int sg = r % S;
                const float *row = s_gv[sg / RTB_N_SUB + 1] + (size_t) rv.z * K;
                const unsigned long long ra = reinterpret_cast<unsigned long long>(row);
                const unsigned long long gd =
                    (unsigned long long) __double_as_longlong((double) __int_as_float(rv.x));
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slab + 16u * (unsigned) r),
                             "r"((unsigned) gd), "r"((unsigned) (gd >> 32)), "r"((unsigned) ra),
                             "r"((unsigned) (ra >> 32))
                             : "memory");
            }
            __syncwarp();
        }
        const unsigned ray_slab = slab + 16u * (unsigned) ((j & (batch - 1)) * S);
        const unsigned meta = __shfl_sync(0xffffffffu, meta_l, j);
        if (meta & RTB_META_INACTIVE)
            continue;
        const double f = __shfl_sync(0xffffffffu, f_l, j);
        const int pix = __shfl_sync(0xffffffffu, pix_l, j);
        const int bin = __shfl_sync(0xffffffffu, bin_l, j);
        const bool invalid = (meta & RTB_META_INVALID) != 0;
        int code = invalid ? 1 : 0;
        double w = 0.0;
        bool bad = false;
        for (int kbase = 0; kbase < K && !bad; kbase += 32 * KS) {
            double Iv[KS];
#pragma unroll
            for (int q = 0; q < KS; q++) {
                const int k = kbase + lane + 32 * q;
                // (the seed spectrum is re-read per ray: holding it in registers costs more in
                // spills than the three L1 hits, measured 5.46 vs 5.19 ms on seed_small)
                Iv[q] = (f != 0.0 && k < K) ? __dmul_rn(f, __ldg(&P.seed_fv[k])) : 0.0;
            }
            if (!invalid) {
                const int cc = batched ? integrate_ray_gain_slab<KS>(P, meta, lane, kbase, Iv, KC, ray_slab)
                               : gain_only
                                   ? integrate_ray_gain_fast<KS>(P, s_gv, h.seg + slot * S, meta, lane, kbase, Iv, KC, slab)
                                   : integrate_ray<KS>(P, h.seg + slot * S, meta, lane, kbase, Iv, exp_tab);
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
            if (combine) {
                if (code == 0 && !invalid) {
                    if (pix != cur_pix) {
                        flush_pix();
                        cur_pix = pix;
                    }
#pragma unroll
                    for (int q = 0; q < KS; q++) {
                        const int k = lane + 32 * q;
                        if (k < K) {
                            w += __ldg(&P.dv2[k]) * Iv[q];
                            acc[q] += Iv[q] * P.scale;
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
                if (batched)
                    integrate_ray_gain_slab<KS>(P, meta, lane, kbase, Iv, KC, ray_slab);
                else if (gain_only)
                    integrate_ray_gain_fast<KS>(P, s_gv, h.seg + slot * S, meta, lane, kbase, Iv, KC, slab);
                else
                    integrate_ray<KS>(P, h.seg + slot * S, meta, lane, kbase, Iv, exp_tab);
#pragma unroll
                for (int q = 0; q < KS; q++) {
                    const int k = kbase + lane + 32 * q;
                    if (k < K) {
                        w += __ldg(&P.dv2[k]) * Iv[q];
                        if (pix >= 0)
                            atomicAdd(&o.image[(size_t) K * (size_t) pix + k], Iv[q] * P.scale);
                    }
                }
            }
        }
        if (code == 0 && !invalid && o.I_ang) {
            if (bin != cur_bin) { // one warp reduction per run of rays with the same bin
                flush_bin();
                cur_bin = bin;
            }
            acc_w += w;
        }
        if (lane == 0) {
            if (o.error)
                o.error[slot] = -code;
            if (code >= 2) {
                float rx, ry, ra, rb;
                int pi, pj, ka, m;
                slot_ray(slot, rx, ry, ra, rb, pi, pj, ka, m);
                report_failure(o.fail, code, rx, ry, ra, rb);
            }
        }
    }
    }
    flush_pix();
    flush_bin();
    }
}

// ------------------------------------------------------------------------------------------
// Seeded image, grid mode (method 2: gain-only integration of the seed spectrum along every ray
// of the seed beam, binned by the EXIT ray; src/RayTraceImageCPU.cpp:40-60,
// RayTraceImageHelper.h:569-581): the lean form of integrate_scatter_kernel for what a seeded
// create_image needs - no per-ray dumps, no explicit ray list, one pass over K <= 32 KS bins.
// One warp per 64 consecutive ray slots, lane = frequency bin.  What is per ray and not per bin
// (seed amplitude, destination pixel and angular bin: four index searches) is evaluated by one
// lane per ray, 32 rays at a time, and handed out as warp-uniform values; the hand-off records of
// a batch of rays are fetched with coalesced loads into the warp's shared-memory slab; a ray then
// costs one 16-byte broadcast read, KS lineshape loads and KS DFMAs per record, and KS exp at the
// end.  Consecutive rays mostly leave the plasma through the same pixel: their spectra are summed
// in registers and flushed with one set of FP64 atomics when the destination changes.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void seeded_slot_ray(const DevProblem &P, const Chunk &c, long long slot, float &rx,
                                                float &ry, float &ra, float &rb, int &pi, int &pj, int &ka, int &m)
{
    const unsigned lq = (unsigned) slot / (unsigned) P.ab_max; // slots fit 32 bits
    const long long p = phys_pixel(P, c, c.pix0 + lq);
    const int t = (int) ((unsigned) slot - lq * (unsigned) P.ab_max);
    const PixelRays pr = pixel_rays(P, p);
    const int ab = pr.ab0 + t * (int) P.n_parallel;
    ka = (int) ((unsigned) ab / (unsigned) P.snb);
    m = ab - ka * P.snb;
    pi = pr.i;
    pj = pr.j;
    rx = __ldg(&P.sxf[pi]);
    ry = __ldg(&P.syf[pj]);
    ra = __ldg(&P.saf[ka]);
    rb = __ldg(&P.sbf[m]);
}

template <int KS>
__global__ void __launch_bounds__(256, RTB_SCATTER_MINBLOCKS)
    integrate_seeded_kernel(const DevProblem P, const Chunk c, const Handoff h, const Outputs o)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double exp_tab[RTB_EXP_TABLE_SIZE];
    __shared__ double s_seed[128], s_dv2[128]; // seed spectrum and 2 dv per bin (zero past the last bin)
    uint4 *rec_slab = reinterpret_cast<uint4 *>(smem_raw);
    // the lineshape tables in double (DevPlane::gvd: packed for every gain-only problem)
    const double **s_gv = reinterpret_cast<const double **>(smem_raw + 8 * RTB_SLAB_RECORDS * sizeof(uint4)); // [N]
    for (int i = threadIdx.x; i < P.N; i += blockDim.x)
        s_gv[i] = P.planes[i].gvd;
    for (int k = threadIdx.x; k < 128; k += blockDim.x) {
        s_seed[k] = k < P.K ? __ldg(&P.seed_fv[k]) : 0.0;
        s_dv2[k] = k < P.K ? __ldg(&P.dv2[k]) : 0.0;
    }
    load_exp_table(exp_tab); // includes __syncthreads()
    // (constants of exp pinned like in the owner kernel: uniform registers instead of four
    // constant-bank loads and a rebuilt shared-memory window per ray)
    const PinnedConsts KC(P.kfp_g, exp_tab);
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = (int) uniform_u32(threadIdx.x >> 5);
    const unsigned slab = uniform_u32((unsigned) __cvta_generic_to_shared(rec_slab + warp_in_cta * RTB_SLAB_RECORDS));
    const int warp_id = (int) (blockIdx.x * (blockDim.x >> 5)) + warp_in_cta;
    const int n_warps = (int) ((gridDim.x * blockDim.x) >> 5);
    const int n_slots = (int) ((c.pix1 - c.pix0) * P.ab_max); // fits 31 bits (the host sizes chunks that way)
    const int S = (P.N - 1) * RTB_N_SUB;
    const int K = P.K;
    constexpr int RUN = 64;
    // this lane's bins: k = lane + 32 q; lanes past the last bin shadow bin K-1 with a zero seed
    int koff[KS];
    bool live[KS];
    double acc[KS];
#pragma unroll
    for (int q = 0; q < KS; q++) {
        const int k = lane + 32 * q;
        live[q] = k < K;
        koff[q] = min(k, K - 1);
        acc[q] = 0.0;
    }
    // (per-bin constants are re-read from shared memory per ray: cheaper than the registers)
    const unsigned tab_seed = uniform_u32((unsigned) __cvta_generic_to_shared(s_seed)) + 8u * (unsigned) lane;
    const unsigned tab_dv2 = uniform_u32((unsigned) __cvta_generic_to_shared(s_dv2)) + 8u * (unsigned) lane;
    auto lds64 = [](unsigned a) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
        return v;
    };
    double acc_w = 0.0;
    int cur_pix = -1, cur_bin = -1; // (warp-uniform)
    auto flush_pix = [&]() {
#pragma unroll
        for (int q = 0; q < KS; q++) {
            if (cur_pix >= 0 && live[q])
                atomicAdd(&o.image[(size_t) K * (size_t) cur_pix + (size_t) (lane + 32 * q)], acc[q]);
            acc[q] = 0.0; // also after a run of rays that left the image (cur_pix < 0): dropped
        }
        cur_pix = -1;
    };
    auto flush_bin = [&]() { // acc_w: this lane's share of the run's angular-bin sum
        if (cur_bin >= 0) {
            const double w = warp_sum(acc_w);
            if (lane == 0)
                atomicAdd(&o.I_ang[cur_bin], w);
        }
        acc_w = 0.0;
        cur_bin = -1;
    };
    // rays whose records are fetched together: as many as fit the slab
    int batch = 32;
    while (batch > 1 && batch * S > RTB_SLAB_RECORDS)
        batch >>= 1;
    const int n_runs = (n_slots + RUN - 1) / RUN;
    for (int run = warp_id; run < n_runs; run += n_warps) {
        const int slot_end = min((run + 1) * RUN, n_slots);
        for (int base = run * RUN; base < slot_end; base += 32) {
            // ---- per ray, one lane each: meta word, seed amplitude, destination ----
            unsigned meta_l = RTB_META_INACTIVE;
            double f_l = 0.0;
            int pix_l = -1, bin_l = -1;
            if (base + lane < slot_end) {
                const long long slot = base + lane;
                meta_l = __ldg(&h.meta[slot]);
                if (!(meta_l & RTB_META_INACTIVE)) {
                    float rx, ry, ra, rb;
                    int pi, pj, ka, m;
                    seeded_slot_ray(P, c, slot, rx, ry, ra, rb, pi, pj, ka, m);
                    if (P.seed_fx && !(meta_l & RTB_META_ESCAPED)) {
                        // calc_seed_inline (:230-247) from the per-index tables
                        const double fx = __ldg(&P.seed_fx[pi]), fy = __ldg(&P.seed_fy[pj]);
                        const double fa = __ldg(&P.seed_fa[ka]), fb = __ldg(&P.seed_fb[m]);
                        if (fx == fx && fy == fy && fa == fa && fb == fb) {
                            f_l = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(P.seed_f0, fx), fy), fa), fb);
                            f_l = f_l < 0.0 ? 0.0 : f_l;
                        }
                    }
                    if (!(meta_l & RTB_META_INVALID)) {
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
                        const int i1 = dev_get_index(P.nx, P.ex, P.edx, (double) bx);
                        const int i2 = dev_get_index(P.ny, P.ey, P.edy, (double) by);
                        const int i3 = dev_get_index(P.na, P.ea, P.eda, (double) ba);
                        const int i4 = dev_get_index(P.nb, P.eb, P.edb, (double) bb);
                        if (i1 >= 0 && i2 >= 0)
                            pix_l = i1 + i2 * P.nx;
                        if (i3 >= 0 && i4 >= 0)
                            bin_l = i3 + i4 * P.na;
                    }
                }
            }
            __syncwarp();
            const int n_here = min(slot_end - base, 32);
            for (int j0 = 0; j0 < n_here; j0 += batch) {
                // ---- one coalesced sweep over the records of the next `batch` rays (contiguous in
                //      the arena); entries outside a ray's visited range are never read ----
                const int nb = min(batch, n_here - j0);
                {
                    const int nrec = nb * S;
                    const SegRec *src = h.seg + (size_t) (base + j0) * S;
                    for (int r = lane; r < nrec; r += 32) {
                        const int4 rv = __ldg(reinterpret_cast<const int4 *>(&src[r]));
                        const int sg = r % S;
                        const double *row = s_gv[sg / RTB_N_SUB + 1] + (size_t) rv.z * K;
                        const unsigned long long ra = reinterpret_cast<unsigned long long>(row);
                        const unsigned long long gd =
                            (unsigned long long) __double_as_longlong((double) __int_as_float(rv.x));
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slab + 16u * (unsigned) r),
                                     "r"((unsigned) gd), "r"((unsigned) (gd >> 32)), "r"((unsigned) ra),
                                     "r"((unsigned) (ra >> 32))
                                     : "memory");
                    }
                }
                __syncwarp();
                for (int jj = 0; jj < nb; jj++) {
                    const int j = j0 + jj;
                    const unsigned meta = uniform_from_lane(meta_l, lane, j);
                    if (meta & RTB_META_INACTIVE)
                        continue;
                    const int pix = (int) uniform_from_lane((unsigned) pix_l, lane, j);
                    const int bin = (int) uniform_from_lane((unsigned) bin_l, lane, j);
                    const double f = __shfl_sync(0xffffffffu, f_l, j);
                    if (meta & RTB_META_INVALID)
                        continue; // error -1, reported by the march
                    // ---- gain-only integration: gl[k] = sum over records of gvl * gv[cell][k]
                    //      (the product of two floats is exact in double, so one DFMA per bin and
                    //      record is the reference's multiply-then-add bit for bit) ----
                    const int lo = meta & 0xfff, hi = (meta >> 12) & 0xfff;
                    const unsigned ray_slab = slab + 16u * (unsigned) (jj * S);
                    double gl[KS];
#pragma unroll
                    for (int q = 0; q < KS; q++)
                        gl[q] = 0.0;
                    auto fetch = [&](int s, double &gvl, double (&g)[KS]) {
                        uint4 e;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                                     : "r"(ray_slab + 16u * (unsigned) s));
                        gvl = __longlong_as_double((long long) (((unsigned long long) e.y << 32) | e.x));
                        const double *row = reinterpret_cast<const double *>(((unsigned long long) e.w << 32) | e.z);
#pragma unroll
                        for (int q = 0; q < KS; q++)
                            g[q] = __ldg(row + koff[q]);
                    };
                    if (lo < hi) {
                        // the row of the next record is requested before the current one is added
                        double gvlA, gvlB;
                        double gA[KS], gB[KS];
                        fetch(lo, gvlA, gA);
                        for (int s = lo;; s += 2) {
                            fetch(min(s + 1, hi - 1), gvlB, gB);
#pragma unroll
                            for (int q = 0; q < KS; q++)
                                gl[q] = __fma_rn(gvlA, gA[q], gl[q]);
                            if (s + 1 >= hi)
                                break;
                            fetch(min(s + 2, hi - 1), gvlA, gA);
#pragma unroll
                            for (int q = 0; q < KS; q++)
                                gl[q] = __fma_rn(gvlB, gB[q], gl[q]);
                            if (s + 2 >= hi)
                                break;
                        }
                    }
                    // exp: one range test for all the ray's bins (|gl| >= 700, inf, NaN take the
                    // library routine), so the KS evaluations run side by side without branches
                    double Iv[KS];
                    bool in_range = true;
#pragma unroll
                    for (int q = 0; q < KS; q++)
                        in_range = in_range && fabs(gl[q]) < 700.0;
                    if (__all_sync(0xffffffffu, in_range)) {
#pragma unroll
                        for (int q = 0; q < KS; q++)
                            Iv[q] = __dmul_rn(f, lds64(tab_seed + 256u * q)) * exp_core(gl[q], KC);
                    } else {
#pragma unroll
                        for (int q = 0; q < KS; q++)
                            Iv[q] = __dmul_rn(f, lds64(tab_seed + 256u * q)) * exp_any(gl[q], KC);
                    }
                    // failed ray: negative (code 2) or NaN (code 3) intensity in a live bin
                    bool bad = false;
#pragma unroll
                    for (int q = 0; q < KS; q++)
                        bad = bad || (live[q] && !(Iv[q] >= 0.0));
                    if (__any_sync(0xffffffffu, bad)) {
                        bool neg = false;
#pragma unroll
                        for (int q = 0; q < KS; q++)
                            neg = neg || (live[q] && Iv[q] < 0.0);
                        const bool any_neg = __any_sync(0xffffffffu, neg); // negative (2) wins over NaN (3)
                        if (lane == 0) {
                            float rx, ry, ra, rb;
                            int pi, pj, ka, m;
                            seeded_slot_ray(P, c, (long long) base + j, rx, ry, ra, rb, pi, pj, ka, m);
                            report_failure(o.fail, any_neg ? 2 : 3, rx, ry, ra, rb);
                        }
                        continue;
                    }
                    // ---- binning: run-length combined in registers ----
                    if (pix != cur_pix) {
                        flush_pix();
                        cur_pix = pix;
                    }
                    if (bin != cur_bin) { // one warp reduction per run of rays with the same bin
                        flush_bin();
                        cur_bin = bin;
                    }
#pragma unroll
                    for (int q = 0; q < KS; q++) {
                        acc_w = __fma_rn(lds64(tab_dv2 + 256u * q), Iv[q], acc_w);
                        acc[q] = __fma_rn(Iv[q], P.scale, acc[q]);
                    }
                }
            }
        }
        flush_pix();
        flush_bin();
    }
}

// Trajectory intensities of RayTrace::calc_ray_path (RAY_DEBUG path of RayTrace_calc_ray,
// :536-566): one warp per explicit ray, lane = frequency bin, K <= 128.  With a debug buffer the
// reference always integrates emission-style (:543), over ALL (segment, sub-segment) records in
// order, and stores I = sum_k (float)(2*Iv[k]*dv[k]) after each one.
__global__ void __launch_bounds__(256)
    path_intensity_kernel(const DevProblem P, const Chunk c, const Handoff h, float *path_I, int *error)
{
    constexpr int KS = 4;
    __shared__ double exp_tab[RTB_EXP_TABLE_SIZE];
    load_exp_table(exp_tab);
    const ArrayConsts KC{ P.kfp, exp_tab };
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long) gridDim.x * blockDim.x) >> 5;
    const long long n = c.ray1 - c.ray0;
    const int S = (P.N - 1) * RTB_N_SUB, N2 = S + 1, K = P.K;
    for (long long slot = warp_id; slot < n; slot += n_warps) {
        const unsigned meta = __ldg(&h.meta[slot]);
        float *out = path_I + slot * N2;
        if (meta & RTB_META_INVALID) { // error -1: returned before any intensity is computed
            if (lane == 0)
                error[slot] = -1;
            continue;
        }
        const float4 r = __ldg(&c.rays[c.ray0 + slot]);
        double f = 0.0;
        if (P.seed_fv && !(meta & RTB_META_ESCAPED)) {
            if (P.method == 1) {
                const float4 e = h.exit_ray[slot];
                f = dev_calc_seed(P, (double) e.x, (double) e.y, (double) e.z, (double) e.w);
            } else {
                f = dev_calc_seed(P, (double) r.x, (double) r.y, (double) r.z, (double) r.w);
            }
        }
        double Iv[KS], dv[KS];
        int koff[KS];
#pragma unroll
        for (int q = 0; q < KS; q++) {
            const int k = lane + 32 * q;
            koff[q] = min(k, K - 1);
            dv[q] = k < K ? 0.5 * __ldg(&P.dv2[k]) : 0.0;
            Iv[q] = (f != 0.0 && k < K) ? __dmul_rn(f, __ldg(&P.seed_fv[k])) : 0.0;
        }
        auto intensity = [&]() {
            float part = 0.0f;
#pragma unroll
            for (int q = 0; q < KS; q++)
                part += (float) (2 * Iv[q] * dv[q]);
            for (int o = 16; o > 0; o >>= 1)
                part += __shfl_xor_sync(0xffffffffu, part, o);
            return part;
        };
        float I0 = intensity();
        if (lane == 0)
            out[0] = I0;
        const int lo = meta & 0xfff, hi = (meta >> 12) & 0xfff;
        const SegRec *seg = h.seg + slot * S;
        for (int sg = 0; sg < S; sg++) {
            if (sg >= lo && sg < hi) {
                const int4 rv = __ldg(reinterpret_cast<const int4 *>(&seg[sg]));
                const float gvl = __int_as_float(rv.x), evl = __int_as_float(rv.y);
                if (!(gvl == 0.0f && evl == 0.0f)) {
                    const float *row = P.planes[sg / RTB_N_SUB + 1].gv + (size_t) rv.z * K;
#pragma unroll
                    for (int q = 0; q < KS; q++) {
                        const float g = __ldg(row + koff[q]);
                        const float glf = __fmul_rn(gvl, g), elf = __fmul_rn(evl, g);
                        const double gl = (double) glf, el = (double) elf;
                        if (!(fabsf(glf) < 700.0f))
                            Iv[q] = ase_update_library(Iv[q], gl, el);
                        else if (fabsf(glf) < 1e-3f)
                            Iv[q] = ase_update_small(Iv[q], gl, el, KC);
                        else
                            Iv[q] = ase_update_large(Iv[q], gl, el, rcp_approx(glf), KC);
                    }
                }
            }
            const float Is = intensity();
            if (lane == 0)
                out[sg + 1] = Is;
        }
        bool neg = false, nan = false;
#pragma unroll
        for (int q = 0; q < KS; q++) {
            neg = neg || Iv[q] < 0.0;
            nan = nan || Iv[q] != Iv[q];
        }
        const bool any_neg = __any_sync(0xffffffffu, neg), any_nan = __any_sync(0xffffffffu, nan);
        if (lane == 0)
            error[slot] = any_neg ? -2 : (any_nan ? -3 : 0);
    }
}

void launch_path_intensity(const DevProblem &P, const Chunk &c, const Handoff &h, float *path_I,
                           int *error, cudaStream_t st)
{
    const long long n = c.ray1 - c.ray0;
    if (n <= 0)
        return;
    long long blocks = std::min<long long>((n + 7) / 8, 148LL * 8 * 4);
    path_intensity_kernel<<<(unsigned) blocks, 256, 0, st>>>(P, c, h, path_I, error);
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
    const int ks = std::min(4, (P.K + 31) / 32);
    const size_t smem = 8 * RTB_SLAB_RECORDS * sizeof(uint4) + sizeof(float *) * (size_t) P.N;
    const int S = (P.N - 1) * RTB_N_SUB;
#ifndef RTB_NO_SEEDED_KERNEL
    // the seeded image of create_image: the lean kernel
    if (!list_mode && P.use_emis == 0 && o.image && o.I_ang && !o.Iv && !o.error && P.K <= 32 * ks && S >= 1 &&
        S <= RTB_SLAB_RECORDS) {
        switch (ks) {
        case 1: integrate_seeded_kernel<1><<<(unsigned) blocks, threads, smem, st>>>(P, c, h, o); break;
        case 2: integrate_seeded_kernel<2><<<(unsigned) blocks, threads, smem, st>>>(P, c, h, o); break;
        case 3: integrate_seeded_kernel<3><<<(unsigned) blocks, threads, smem, st>>>(P, c, h, o); break;
        default: integrate_seeded_kernel<4><<<(unsigned) blocks, threads, smem, st>>>(P, c, h, o); break;
        }
        return;
    }
#endif
#define RTB_LAUNCH_SCATTER(KS_)                                                                  \
    do {                                                                                         \
        if (list_mode)                                                                           \
            integrate_scatter_kernel<true, KS_><<<(unsigned) blocks, threads, smem, st>>>(P, c, h, o);  \
        else                                                                                     \
            integrate_scatter_kernel<false, KS_><<<(unsigned) blocks, threads, smem, st>>>(P, c, h, o); \
    } while (0)
    switch (ks) {
    case 1: RTB_LAUNCH_SCATTER(1); break;
    case 2: RTB_LAUNCH_SCATTER(2); break;
    case 3: RTB_LAUNCH_SCATTER(3); break;
    default: RTB_LAUNCH_SCATTER(4); break;
    }
#undef RTB_LAUNCH_SCATTER
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

// ------------------------------------------------------------------------------------------
// Exhaustive check of fdiv_refined against the correctly rounded quotient: divisor significands [b_first, b_first +
// b_count) x ALL 2^23 numerator significands, at binary exponents ea / eb.  One thread per
// (divisor, 128 consecutive numerators).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fdiv_check_kernel(unsigned b_first, unsigned b_count, int ea, int eb, int variant,
                                                         unsigned long long *out /* [0] mismatches, [1] a bits, [2] b bits */)
{
    const unsigned long long tid = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long total = (unsigned long long) b_count << 16; // 2^23 / 128 numerator groups
    unsigned long long bad = 0;
    for (unsigned long long w = tid; w < total; w += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned bm = b_first + (unsigned) (w >> 16);
        const unsigned a0 = (unsigned) (w & 0xffffu) << 7;
        const float b = __uint_as_float(((unsigned) (eb + 127) << 23) | (bm & 0x7fffffu));
        const float r = frcp_refined(b);
#if defined(__CUDA_ARCH__) && !defined(RTB_NO_F32X2)
        if (variant == 3) { // the packed form (FFMA2): both lanes, reciprocal refined in the pair as well
            for (unsigned i = 0; i < 128u; i += 2) {
                const float a = __uint_as_float(((unsigned) (ea + 127) << 23) | (a0 + i));
                const float a2 = __uint_as_float(((unsigned) (ea + 127) << 23) | (a0 + i + 1u));
                float q, q2, p, p2;
                fdiv_refined2(a, b, a2, b, q, q2);
                fdiv_refined2_by(a2, a, b, r, p2, p);
                const float want = __double2float_rn(__ddiv_rn((double) a, (double) b));
                const float want2 = __double2float_rn(__ddiv_rn((double) a2, (double) b));
                if (__float_as_uint(q) != __float_as_uint(want) || __float_as_uint(p) != __float_as_uint(want)) {
                    bad++;
                    out[1] = __float_as_uint(a);
                    out[2] = __float_as_uint(b);
                }
                if (__float_as_uint(q2) != __float_as_uint(want2) || __float_as_uint(p2) != __float_as_uint(want2)) {
                    bad++;
                    out[1] = __float_as_uint(a2);
                    out[2] = __float_as_uint(b);
                }
            }
            continue;
        }
#endif
#pragma unroll 4
        for (unsigned i = 0; i < 128u; i++) {
            const float a = __uint_as_float(((unsigned) (ea + 127) << 23) | (a0 + i));
            // Independent oracle: the FP64 quotient rounded to float.  53 >= 2*24 + 2 bits, so the
            // double rounding is innocuous and this IS the correctly rounded float quotient
            // (comparing with __fdiv_rn would compare the sequence with itself).
            // variant 1 (detector self-test): the uncorrected product a*r, which is only faithful
            const float q = variant == 1 ? __fmul_rn(a, r) : fdiv_refined(a, b, r);
            const float want = __double2float_rn(__ddiv_rn((double) a, (double) b));
            if (__float_as_uint(q) != __float_as_uint(want)) {
                bad++;
                out[1] = __float_as_uint(a);
                out[2] = __float_as_uint(b);
            }
        }
    }
    if (bad)
        atomicAdd(&out[0], bad);
}

// Exhaustive check of fsqrt_refined and of the reciprocal that follows it in normalize_s:
// significands [b_first, b_first + b_count) at binary exponents eb and eb + 1 (both parities).
__global__ void __launch_bounds__(256) fsqrt_check_kernel(unsigned b_first, unsigned b_count, int eb,
                                                          unsigned long long *out)
{
    const unsigned long long tid = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad = 0;
    for (unsigned long long w = tid; w < 2ull * b_count; w += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned bm = b_first + (unsigned) (w >> 1);
        const float x = __uint_as_float(((unsigned) (eb + (int) (w & 1u) + 127) << 23) | (bm & 0x7fffffu));
        const float got = fsqrt_refined(x), want = __fsqrt_rn(x);
        const float rgot = fdiv_refined(1.0f, got, frcp_refined(got));
        const float rwant = __double2float_rn(__ddiv_rn(1.0, (double) want));
        if (__float_as_uint(got) != __float_as_uint(want) || __float_as_uint(rgot) != __float_as_uint(rwant)) {
            bad++;
            out[1] = __float_as_uint(x);
            out[2] = __float_as_uint(got);
        }
    }
    if (bad)
        atomicAdd(&out[0], bad);
}

void launch_fdiv_check(unsigned b_first, unsigned b_count, int ea, int eb, int variant,
                       unsigned long long *out, cudaStream_t st)
{
    if (variant == 2) {
        fsqrt_check_kernel<<<148 * 16, 256, 0, st>>>(b_first, b_count, eb, out);
        return;
    }
    fdiv_check_kernel<<<148 * 16, 256, 0, st>>>(b_first, b_count, ea, eb, variant, out);
}

// ------------------------------------------------------------------------------------------
// Cell records (CellRec) of every plane, derived on the device: record i1 = (k1-1) + (k2-1)*Nx
// holds the corners i1, i1+1, i1+Nx, i1+Nx+1 in the reference's order
// (RayTraceImageHelper.h:474-477) - the float corner values, their four double differences
// (:333-334) and the cell's copy of the two axis intervals.  One thread per node index; the last
// row and column are never addressed (zero).  Same IEEE operations as fill_cell_records.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_cell_records_kernel(const DevPlane *planes)
{
    const DevPlane &D = planes[blockIdx.y];
    const int Nx = D.Nx, Ny = D.Ny;
    const long long i1 = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i1 >= (long long) Nx * Ny)
        return;
    const int j = (int) (i1 / Nx), i = (int) (i1 - (long long) j * Nx);
    CellRec r;
    if (i + 1 >= Nx || j + 1 >= Ny) {
        memset(&r, 0, sizeof(r));
    } else {
        const long long c[4] = { i1, i1 + 1, i1 + Nx, i1 + Nx + 1 };
        double nc[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const Node nd = load_node(&D.node[c[q]]);
            nc[q] = nd.n;
            r.nf[q] = __double2float_rn(nd.n);
            r.g0[q] = nd.g0;
            r.E0[q] = nd.E0;
        }
        r.n10 = __dsub_rn(nc[1], nc[0]);
        r.n32 = __dsub_rn(nc[3], nc[2]);
        r.n20 = __dsub_rn(nc[2], nc[0]);
        r.n31 = __dsub_rn(nc[3], nc[1]);
        const AxisCell &ax = D.cx[i + 1], &ay = D.cy[j + 1];
        r.xl = ax.lo;
        r.dxd = ax.dd;
        r.rdx = ax.rd;
        r.yl = ay.lo;
        r.dyd = ay.dd;
        r.rdy = ay.rd;
    }
    const_cast<CellRec *>(D.cell)[i1] = r;
}

void launch_build_cell_records(const DevPlane *planes, int N, long long max_nodes, cudaStream_t st)
{
    if (N <= 0 || max_nodes <= 0)
        return;
    const dim3 grid((unsigned) ((max_nodes + 255) / 256), (unsigned) N);
    build_cell_records_kernel<<<grid, 256, 0, st>>>(planes);
}

// The lineshape tables of a gain-only problem once more in double (the seeded kernel's operand
// type): widened on the device from the uploaded float tables, so the host neither converts nor
// uploads them.
__global__ void __launch_bounds__(256) widen_gv_kernel(const DevPlane *planes, int K)
{
    const DevPlane &D = planes[blockIdx.y];
    const size_t n = (size_t) D.Nx * D.Ny * (size_t) K;
    double *dst = const_cast<double *>(D.gvd);
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        dst[i] = (double) __ldg(&D.gv[i]);
}

void launch_widen_gv(const DevPlane *planes, int N, int K, cudaStream_t st)
{
    if (N <= 0 || K <= 0)
        return;
    widen_gv_kernel<<<dim3(148 * 2, (unsigned) N), 256, 0, st>>>(planes, K);
}

void launch_fp64_peak(double *out, int iters, cudaStream_t st, int *blocks, int *threads)
{
    *blocks = 148 * 8;
    *threads = 256;
    fp64_peak_kernel<<<*blocks, *threads, 0, st>>>(out, iters);
}

} // namespace rtb
