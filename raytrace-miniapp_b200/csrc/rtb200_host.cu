// rtb200_host.cu — host layer of librtb200.so: context (stream, pinned staging, device arena),
// problem staging, chunked launch of the march / integration kernels, and the extern "C" ABI
// declared in include/rtb200.h.  C++ in the reference's own style; no CPU fallback: every
// compute entry point fails with RTB200_ERR_CUDA when the device is not usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "rtb200_kernels.cuh"
#include "rtb200_pack.h"

using namespace rtb;

#define RTB_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
            return RTB200_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)

namespace {

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0; // elements
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc((void **) &p, want * sizeof(T));
        if (e == cudaSuccess)
            cap = want;
        return e;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinBuf {
    char *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = n + n / 4 + 4096;
        cudaError_t e = cudaMallocHost((void **) &p, want);
        if (e == cudaSuccess)
            cap = want;
        return e;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

} // namespace

struct rtb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // uploads the lineshape tables beside the running march
    cudaEvent_t gv_ready = nullptr;
    // Ordering between calls that may use different streams (the caller's stream of
    // rtb200_launch*, the context's own stream, the copy stream): every launch ends with
    // launch_done, every staging with stage_done, and whatever touches the shared arenas
    // (blob, lineshape tables, hand-off, work counter, failure state) next waits for them.
    cudaEvent_t stage_done = nullptr, launch_done = nullptr;
    bool have_launch_done = false;
    std::string err;
    // staging
    PinBuf h_blob, h_gv;
    DevBuf<char> d_blob, d_gv; // d_gv: the lineshape tables (read by the integration only)
    DevBuf<char> d_cells;      // per-cell records, derived on the device (CellBlob, rtb200_pack.h)
    DevBuf<char> d_gvd;        // lineshape tables in double (gain-only problems), widened on the device
    // Overlapped march + integration (owner kernel; RTB200_OVERLAP=0 turns it off): the march counts
    // the closed ray slots of every pixel, the integration is launched with programmatic stream
    // serialization right behind it and each of its CTAs waits for its own pixel.
    DevBuf<unsigned> d_pix_done; // per-pixel completion counts
    DevBuf<unsigned> d_gv_flag;  // epoch of the lineshape tables resident in d_gv (written by a copy)
    unsigned *h_gv_epoch = nullptr; // pinned source of that copy
    unsigned gv_epoch = 0;
    bool overlap = true;
    const rtb200_problem *gv_pending = nullptr; // tables not filled/uploaded yet (create_image)
    size_t gv_bytes = 0;
    DevProblem prob;
    bool staged = false;
    bool owner_ok = false; // ASE owner kernel usable (identity-like, injective pixel map)
    long long staged_pixels = 0, staged_rays = 0;
    // hand-off
    DevBuf<SegRec> d_seg;
    DevBuf<unsigned> d_meta;
    DevBuf<float4> d_exit;
    long long slots_per_chunk = 0;
    // outputs for the host-buffer API
    DevBuf<double> d_image, d_iang, d_Iv;
    DevBuf<int> d_err;
    DevBuf<float4> d_rays;
    DevBuf<float2> d_tans;
    DevBuf<float2> d_path_xy;
    DevBuf<float> d_path_I;
    PinBuf h_out;
    FailState *d_fail = nullptr;
    FailState *h_fail = nullptr; // pinned
    // timing
    std::vector<cudaEvent_t> ev;
    size_t ev_used = 0;
    std::vector<std::pair<size_t, size_t>> ev_march, ev_integ; // (start, stop) indices
    std::pair<size_t, size_t> ev_h2d{ 0, 0 }, ev_d2h{ 0, 0 };
    bool have_h2d = false, have_d2h = false;
    rtb200_timings last;
    int launches = 0;
    bool count_steps = false;
    bool ieee_div = false; // never take the reciprocal-table division path (ddiv_by)
    int march_blocks = 0; // grid of the persistent march; 0 = resident CTAs x SMs (tuning override)
    unsigned long long *d_work = nullptr;
    size_t handoff_bytes = (size_t) 16384 << 20; // B200 has 180 GB: one chunk for every shipped / synthetic size (allocated to need)
};

namespace {

size_t new_event(rtb200_ctx *ctx, cudaStream_t st)
{
    if (ctx->ev_used == ctx->ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev.push_back(e);
    }
    cudaEventRecord(ctx->ev[ctx->ev_used], st);
    return ctx->ev_used++;
}

void reset_timing(rtb200_ctx *ctx)
{
    ctx->ev_used = 0;
    ctx->ev_march.clear();
    ctx->ev_integ.clear();
    ctx->have_h2d = ctx->have_d2h = false;
    ctx->launches = 0;
}

float ev_ms(rtb200_ctx *ctx, std::pair<size_t, size_t> p)
{
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[p.first], ctx->ev[p.second]);
    return ms;
}

void collect_timing(rtb200_ctx *ctx)
{
    rtb200_timings &t = ctx->last;
    std::memset(&t, 0, sizeof(t));
    if (ctx->have_h2d)
        t.h2d_ms = ev_ms(ctx, ctx->ev_h2d);
    if (ctx->have_d2h)
        t.d2h_ms = ev_ms(ctx, ctx->ev_d2h);
    for (auto &p : ctx->ev_march)
        t.march_ms += ev_ms(ctx, p);
    for (auto &p : ctx->ev_integ)
        t.integrate_ms += ev_ms(ctx, p);
    if (ctx->ev_used >= 2)
        t.total_ms = ev_ms(ctx, { 0, ctx->ev_used - 1 });
    t.kernel_launches = ctx->launches;
    t.n_rays = (uint64_t) ctx->staged_rays;
    t.march_steps = ctx->h_fail ? ctx->h_fail->march_steps : 0;
}

// The kernels read seed->f[4][k] for every frequency bin k < nv and interpolate in the four
// spatial / angular tables: reject tables that are missing or too short.
int validate_seed(rtb200_ctx *ctx, const rtb200_seed *s, int nv)
{
    if (!s)
        return RTB200_OK;
    for (int d = 0; d < 5; d++)
        if (s->dim[d] < (d < 4 ? 2 : 1) || !s->x[d] || !s->f[d]) {
            ctx->err = "invalid seed (NULL table or fewer than 2 points per axis)";
            return RTB200_ERR_ARG;
        }
    if (s->dim[4] != nv) {
        ctx->err = "seed spectrum length differs from the number of frequency bins";
        return RTB200_ERR_ARG;
    }
    return RTB200_OK;
}

// check_grid (src/RayTraceImage.cpp:237-242)
bool grid_error(int n, double dx, const double *x)
{
    bool error = false;
    for (int i = 1; i < n; i++)
        error = error || (std::fabs((x[i] - x[i - 1]) - dx) > 1e-12 * dx);
    return error;
}

int validate(rtb200_ctx *ctx, const rtb200_problem *p, unsigned flags)
{
    if (!p || !p->euv_beam || !p->gain || p->N < 1 || p->N_parallel < 1 || p->N_start < 0) {
        ctx->err = "invalid problem (NULL member, N < 1, N_parallel < 1 or N_start < 0)";
        return RTB200_ERR_ARG;
    }
    if (p->seed && !p->seed_beam) {
        ctx->err = "seed given without seed_beam";
        return RTB200_ERR_ARG;
    }
    if (int rs = validate_seed(ctx, p->seed, p->euv_beam->nv))
        return rs;
    const rtb200_beam &e = *p->euv_beam;
    if (!(flags & RTB200_FLAG_NO_LIMITS)) { // src/RayTraceImage.cpp:229-232
        if (p->N > RTB200_N_MAX) {
            ctx->err = "Exceeded maximum number of length segments";
            return RTB200_ERR_LIMITS;
        }
        if (e.nv >= RTB200_K_MAX) {
            ctx->err = "Exceeded maximum number of frequencies";
            return RTB200_ERR_LIMITS;
        }
    }
    {
        const rtb200_beam &g = p->seed ? *p->seed_beam : e;
        if ((long long) g.nx * g.ny >= (1LL << 31) || (long long) g.na * g.nb >= (1LL << 31) ||
            (long long) e.nx * e.ny >= (1LL << 31) || (long long) e.na * e.nb >= (1LL << 31)) {
            ctx->err = "more than 2^31 source pixels or angles per pixel";
            return RTB200_ERR_LIMITS;
        }
    }
    if ((p->N - 1) * RTB200_N_SUB > RTB_MAX_SEGS) {
        ctx->err = "too many length segments for the hand-off record";
        return RTB200_ERR_LIMITS;
    }
    if (grid_error(e.nx, e.dx, e.x) || grid_error(e.ny, e.dy, e.y) ||
        grid_error(e.na, e.da, e.a) || grid_error(e.nb, e.db, e.b)) { // :243-250
        ctx->err = "Only uniform grid spacings are currently supported (euv_beam)";
        return RTB200_ERR_GRID;
    }
    if (p->seed_beam) { // :253-264
        const rtb200_beam &s = *p->seed_beam;
        if (grid_error(s.nx, s.dx, s.x) || grid_error(s.ny, s.dy, s.y) ||
            grid_error(s.na, s.da, s.a) || grid_error(s.nb, s.db, s.b)) {
            ctx->err = "Only uniform grid spacings are currently supported (seed_beam)";
            return RTB200_ERR_GRID;
        }
        if ((e.y[0] >= 0.0) != (s.y[0] >= 0.0)) {
            ctx->err = "Negitive y positions in seed_beam or euv_beam, but not both";
            return RTB200_ERR_GRID;
        }
    }
    for (int i = 0; i < p->N; i++) {
        const rtb200_gain_plane &g = p->gain[i];
        if (g.Nx < 2 || g.Ny < 2 || g.Nv != e.nv || !g.x || !g.y || !g.n || !g.g0 || !g.gv) {
            ctx->err = "invalid gain plane (Nx, Ny < 2, Nv != nv or NULL array)";
            return RTB200_ERR_ARG;
        }
    }
    return RTB200_OK;
}

bool injective(const int *t, int n, int range)
{
    std::vector<char> seen((size_t) std::max(range, 1), 0);
    for (int i = 0; i < n; i++) {
        if (t[i] < 0)
            continue;
        if (t[i] >= range || seen[t[i]])
            return false;
        seen[t[i]] = 1;
    }
    return true;
}

// Fills and uploads the lineshape tables if that is still pending (see stage_impl).
// device_flag: `st` does NOT wait for the copy; instead the epoch word d_gv_flag is written behind
// it on the copy stream and the integration kernel polls that word (overlapped launch).
int flush_gv(rtb200_ctx *ctx, cudaStream_t st, bool device_flag = false)
{
    if (!ctx->gv_pending)
        return RTB200_OK;
    ctx->prob.gv_absmax_bits = pack_gv(*ctx->gv_pending, ctx->h_gv.p);
    ctx->gv_pending = nullptr;
    // The copy runs on its own stream next to the march; `st` only waits for its completion.
    // (d_gv is not read by anything that is still in flight: the previous image's integration
    // finished before its results were read back.)
    RTB_CUDA(cudaMemcpyAsync(ctx->d_gv.p, ctx->h_gv.p, ctx->gv_bytes, cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    if (!ctx->prob.use_emis) { // (the plane descriptors were uploaded on ctx->stream: stage_done)
        RTB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_done, 0));
        launch_widen_gv(ctx->prob.planes, ctx->prob.N, ctx->prob.K, ctx->copy_stream);
    }
    if (device_flag) {
        *ctx->h_gv_epoch = ++ctx->gv_epoch;
        RTB_CUDA(cudaMemcpyAsync(ctx->d_gv_flag.p, ctx->h_gv_epoch, sizeof(unsigned), cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
        return RTB200_OK;
    }
    RTB_CUDA(cudaEventRecord(ctx->gv_ready, ctx->copy_stream));
    RTB_CUDA(cudaStreamWaitEvent(st, ctx->gv_ready, 0));
    return RTB200_OK;
}

// defer_gv: leave the lineshape tables (most of the bytes, not needed by the march) to
// flush_gv(), which the caller runs on the host while the march kernel is already executing.
int stage_impl(rtb200_ctx *ctx, const rtb200_problem *p, bool explicit_rays, int method,
               double scale, bool defer_gv = false)
{
    RTB_CUDA(cudaSetDevice(ctx->device));
    // a staging that fails part-way must not leave the previous one "staged": its buffers may
    // already have been freed or overwritten
    ctx->staged = false;
    ctx->gv_pending = nullptr;
    if (ctx->have_launch_done) { // kernels of an earlier launch (any stream) still read the arenas
        RTB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->launch_done, 0));
        RTB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->launch_done, 0));
    }
    DevProblem tmp;
    GvBlob gvb{ nullptr, nullptr, false, 0 };
    CellBlob cellb{ nullptr, 0 }, gvdb{ nullptr, 0 };
    const size_t bytes = pack_problem(*p, explicit_rays, method, scale, nullptr, nullptr, tmp, &gvb, &cellb, &gvdb);
    RTB_CUDA(ctx->h_blob.reserve(bytes));
    RTB_CUDA(ctx->d_blob.reserve(bytes));
    RTB_CUDA(ctx->h_gv.reserve(gvb.bytes));
    RTB_CUDA(ctx->d_gv.reserve(gvb.bytes));
    RTB_CUDA(ctx->d_cells.reserve(cellb.bytes));
    RTB_CUDA(ctx->d_gvd.reserve(gvdb.bytes));
    gvdb.dev = ctx->d_gvd.p;
    gvb.host = ctx->h_gv.p;
    gvb.dev = ctx->d_gv.p;
    gvb.copy = !defer_gv;
    cellb.dev = ctx->d_cells.p;
    pack_problem(*p, explicit_rays, method, scale, ctx->h_blob.p, ctx->d_blob.p, ctx->prob, &gvb, &cellb, &gvdb);
    ctx->gv_bytes = gvb.bytes;
    ctx->gv_pending = defer_gv ? p : nullptr;
    if (ctx->ieee_div) {
        DevPlane *hp = reinterpret_cast<DevPlane *>(ctx->h_blob.p + ((const char *) ctx->prob.planes - ctx->d_blob.p));
        PlaneLite *hl = reinterpret_cast<PlaneLite *>(ctx->h_blob.p + ((const char *) ctx->prob.lite - ctx->d_blob.p));
        for (int i = 0; i < ctx->prob.N; i++) {
            hp[i].fast_div = 0;
            hl[i].flags &= ~2;
        }
    }
    ctx->ev_h2d.first = new_event(ctx, ctx->stream);
    RTB_CUDA(cudaMemsetAsync(ctx->d_fail, 0, sizeof(FailState), ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(ctx->d_blob.p, ctx->h_blob.p, bytes, cudaMemcpyHostToDevice,
                             ctx->stream));
    {
        long long max_nodes = 0;
        for (int i = 0; i < p->N; i++)
            max_nodes = std::max(max_nodes, (long long) p->gain[i].Nx * p->gain[i].Ny);
        launch_build_cell_records(ctx->prob.planes, ctx->prob.N, max_nodes, ctx->stream);
    }
    if (!defer_gv) {
        RTB_CUDA(cudaMemcpyAsync(ctx->d_gv.p, ctx->h_gv.p, gvb.bytes, cudaMemcpyHostToDevice,
                                 ctx->stream));
        if (!ctx->prob.use_emis)
            launch_widen_gv(ctx->prob.planes, ctx->prob.N, ctx->prob.K, ctx->stream);
    }
    ctx->ev_h2d.second = new_event(ctx, ctx->stream);
    RTB_CUDA(cudaEventRecord(ctx->stage_done, ctx->stream));
    ctx->have_h2d = true;
    const DevProblem &P = ctx->prob;
    if (!explicit_rays) {
        ctx->staged_pixels = (long long) P.snx * P.sny;
        const long long Nt = ctx->staged_pixels * P.sna * P.snb;
        ctx->staged_rays = P.n_start < Nt ? (Nt - P.n_start + P.n_parallel - 1) / P.n_parallel : 0;
        // The owner kernel needs each destination pixel to be fed by at most one source pixel.
        const char *hb = ctx->h_blob.p;
        const char *db = ctx->d_blob.p;
        const int *pixI = reinterpret_cast<const int *>(hb + ((const char *) P.pixI - db));
        const int *pixJ = reinterpret_cast<const int *>(hb + ((const char *) P.pixJ - db));
        ctx->owner_ok = P.method == 1 && P.use_emis && injective(pixI, P.snx, P.nx) &&
                        injective(pixJ, P.sny, P.ny);
    } else {
        ctx->staged_pixels = 0;
        ctx->staged_rays = 0;
        ctx->owner_ok = false;
    }
    ctx->staged = true;
    return RTB200_OK;
}

int ensure_handoff(rtb200_ctx *ctx, long long slots, bool need_exit)
{
    const int S = (ctx->prob.N - 1) * RTB_N_SUB;
    RTB_CUDA(ctx->d_seg.reserve((size_t) slots * (size_t) std::max(S, 1)));
    RTB_CUDA(ctx->d_meta.reserve((size_t) slots));
    if (need_exit)
        RTB_CUDA(ctx->d_exit.reserve((size_t) slots));
    return RTB200_OK;
}

// Launches the kernels for source pixels [pix0, pix1) in chunks whose hand-off fits the budget.
int launch_pixels(rtb200_ctx *ctx, long long pix0, long long pix1, const Outputs &out,
                  cudaStream_t st, int row_off = 0, int row_stride = 1)
{
    const DevProblem &P = ctx->prob;
    if (pix1 <= pix0)
        return RTB200_OK;
    if (ctx->have_launch_done) // the hand-off arena and the work counter are shared by all launches
        RTB_CUDA(cudaStreamWaitEvent(st, ctx->launch_done, 0));
    const int S = (P.N - 1) * RTB_N_SUB;
    const bool need_exit = P.method != 1;
    const size_t per_slot = (size_t) std::max(S, 1) * sizeof(SegRec) + sizeof(unsigned) +
                            (need_exit ? sizeof(float4) : 0);
    long long pix_per_chunk =
        std::max<long long>(1, (long long) (ctx->handoff_bytes / per_slot) / std::max(P.ab_max, 1));
    pix_per_chunk = std::min(pix_per_chunk, pix1 - pix0);
    // the kernels count the slots AND the hand-off records of one chunk in 32 bits
    pix_per_chunk = std::min<long long>(pix_per_chunk, ((1LL << 31) - 1) / std::max(P.ab_max, 1) / std::max(S, 1));
    pix_per_chunk = std::max<long long>(pix_per_chunk, 1);
    int rc = ensure_handoff(ctx, pix_per_chunk * P.ab_max, need_exit);
    if (rc)
        return rc;
    Handoff h{ ctx->d_seg.p, ctx->d_meta.p, need_exit ? ctx->d_exit.p : nullptr, nullptr, nullptr };
    // (the owner kernels address a pixel's records with 32-bit byte offsets: OwnerBase::ray)
    const bool owner = ctx->owner_ok && !out.Iv && !out.error &&
                       (long long) P.ab_max * std::max(S, 1) * (long long) sizeof(SegRec) < (1LL << 32);
    // Overlapped form (owner kernel): the march counts the closed ray slots of every pixel, the
    // integration is launched with programmatic stream serialization right behind it - nothing
    // between the two launches in the stream - and each of its CTAs waits for its pixel (and, when
    // the lineshape tables are still on their way, for the word that follows them).
#ifdef RTB_HANDOFF_NC // (hand-off read through L1: not safe next to a running march)
    const bool overlap = false;
#else
    const bool overlap = ctx->overlap && owner && !ctx->count_steps;
#endif
    if (overlap)
        RTB_CUDA(ctx->d_pix_done.reserve((size_t) pix_per_chunk));
    bool wait_gv = false; // this launch's integration kernels poll the tables' epoch word
    for (long long a = pix0; a < pix1; a += pix_per_chunk) {
        Chunk c;
        std::memset(&c, 0, sizeof(c));
        c.pix0 = a;
        c.pix1 = std::min(pix1, a + pix_per_chunk);
        c.row_off = row_off;
        c.row_stride = row_stride;
        if (overlap) {
            RTB_CUDA(cudaMemsetAsync(ctx->d_pix_done.p, 0, sizeof(unsigned) * (size_t) (c.pix1 - c.pix0), st));
            Handoff ho = h;
            ho.pix_done = ctx->d_pix_done.p;
            Outputs oo = out;
            oo.pix_done = ctx->d_pix_done.p;
            const size_t e0 = new_event(ctx, st);
            launch_march(P, c, false, ho, ctx->d_fail, ctx->count_steps, st, ctx->d_work, ctx->march_blocks);
            if (ctx->gv_pending) { // host packs the lineshape tables while the march runs
                rc = flush_gv(ctx, st, true);
                if (rc)
                    return rc;
                wait_gv = true;
            }
            if (wait_gv) {
                oo.gv_flag = ctx->d_gv_flag.p;
                oo.gv_epoch = ctx->gv_epoch;
            }
            launch_integrate_ase_owner(P, c, ho, oo, st, true);
            const size_t e2 = new_event(ctx, st);
            ctx->ev_march.push_back({ e0, e2 }); // (the two kernels overlap: one interval)
            ctx->ev_integ.push_back({ e2, e2 });
            ctx->launches += 2;
            continue;
        }
        const size_t e0 = new_event(ctx, st);
        launch_march(P, c, false, h, ctx->d_fail, ctx->count_steps, st, ctx->d_work, ctx->march_blocks);
        const size_t e1m = new_event(ctx, st);
        rc = flush_gv(ctx, st); // host packs the lineshape tables while the march runs
        if (rc)
            return rc;
        const size_t e1 = new_event(ctx, st);
        if (owner)
            launch_integrate_ase_owner(P, c, h, out, st);
        else
            launch_integrate_scatter(P, c, false, h, out, st);
        const size_t e2 = new_event(ctx, st);
        ctx->ev_march.push_back({ e0, e1m });
        ctx->ev_integ.push_back({ e1, e2 });
        ctx->launches += 2;
    }
    RTB_CUDA(cudaGetLastError());
    RTB_CUDA(cudaEventRecord(ctx->launch_done, st));
    ctx->have_launch_done = true;
    return RTB200_OK;
}

int finish(rtb200_ctx *ctx, unsigned *failure_code, rtb200_ray *failed, int max_failed,
           int *n_failed)
{
    RTB_CUDA(cudaMemcpyAsync(ctx->h_fail, ctx->d_fail, sizeof(FailState), cudaMemcpyDeviceToHost,
                             ctx->stream));
    RTB_CUDA(cudaMemsetAsync(ctx->d_fail, 0, sizeof(FailState), ctx->stream));
    new_event(ctx, ctx->stream);
    RTB_CUDA(cudaStreamSynchronize(ctx->stream));
    collect_timing(ctx);
    const FailState &f = *ctx->h_fail;
    if (failure_code)
        *failure_code = f.failure_code;
    if (n_failed)
        *n_failed = (int) f.n_failed;
    if (failed) {
        const int n = std::min<int>({ (int) f.n_failed, max_failed, 32 });
        for (int i = 0; i < n; i++) {
            failed[i].x = f.failed[4 * i + 0];
            failed[i].y = f.failed[4 * i + 1];
            failed[i].a = f.failed[4 * i + 2];
            failed[i].b = f.failed[4 * i + 3];
        }
    }
    return f.failure_code ? RTB200_RAYS_FAILED : RTB200_OK;
}

} // namespace

// =============================================================================================
// extern "C" ABI
// =============================================================================================
extern "C" {

const char *rtb200_version(void) { return "rtb200 0.1.0 (sm_100a)"; }

int rtb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rtb200_create(int device, rtb200_ctx **out)
{
    if (!out)
        return RTB200_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        return RTB200_ERR_CUDA;
    }
    rtb200_ctx *ctx = new rtb200_ctx;
    ctx->device = device;
    std::memset(&ctx->last, 0, sizeof(ctx->last));
    std::memset(&ctx->prob, 0, sizeof(ctx->prob));
    auto fail = [&](cudaError_t e) {
        (void) e;
        cudaGetLastError();
        rtb200_destroy(ctx);
        return RTB200_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return fail(e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail(e);
    if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return fail(e);
    if ((e = cudaEventCreateWithFlags(&ctx->gv_ready, cudaEventDisableTiming)) != cudaSuccess)
        return fail(e);
    if ((e = cudaEventCreateWithFlags(&ctx->stage_done, cudaEventDisableTiming)) != cudaSuccess)
        return fail(e);
    if ((e = cudaEventCreateWithFlags(&ctx->launch_done, cudaEventDisableTiming)) != cudaSuccess)
        return fail(e);
    if ((e = cudaMalloc((void **) &ctx->d_fail, sizeof(FailState))) != cudaSuccess)
        return fail(e);
    if ((e = cudaMallocHost((void **) &ctx->h_fail, sizeof(FailState))) != cudaSuccess)
        return fail(e);
    if ((e = cudaMalloc((void **) &ctx->d_work, sizeof(unsigned long long))) != cudaSuccess)
        return fail(e);
    std::memset(ctx->h_fail, 0, sizeof(FailState));
    if (const char *s = getenv("RTB200_MARCH_BLOCKS")) // tuning: grid of the persistent march
        ctx->march_blocks = std::max(1, atoi(s));
    if (const char *s = getenv("RTB200_HANDOFF_MB"))
        ctx->handoff_bytes = (size_t) std::max(1, atoi(s)) << 20;
    if ((e = cudaMallocHost((void **) &ctx->h_gv_epoch, sizeof(unsigned))) != cudaSuccess)
        return fail(e);
    *ctx->h_gv_epoch = 0;
    if ((e = ctx->d_gv_flag.reserve(1)) != cudaSuccess || (e = cudaMemset(ctx->d_gv_flag.p, 0, sizeof(unsigned))) != cudaSuccess)
        return fail(e);
    if (const char *s = getenv("RTB200_OVERLAP")) // 0: march and integration strictly one after the other
        ctx->overlap = atoi(s) != 0;
    if (const char *s = getenv("RTB200_COUNT_STEPS"))
        ctx->count_steps = atoi(s) != 0;
    if (const char *s = getenv("RTB200_IEEE_DIV")) // tests: plain IEEE divisions by the cell widths
        ctx->ieee_div = atoi(s) != 0;
    *out = ctx;
    return RTB200_OK;
}

void rtb200_destroy(rtb200_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    ctx->h_blob.release();
    ctx->d_blob.release();
    ctx->h_gv.release();
    ctx->d_gv.release();
    ctx->d_cells.release();
    ctx->d_gvd.release();
    ctx->d_pix_done.release();
    ctx->d_gv_flag.release();
    if (ctx->h_gv_epoch)
        cudaFreeHost(ctx->h_gv_epoch);
    ctx->h_gv_epoch = nullptr;
    ctx->d_seg.release();
    ctx->d_meta.release();
    ctx->d_exit.release();
    ctx->d_image.release();
    ctx->d_iang.release();
    ctx->d_Iv.release();
    ctx->d_err.release();
    ctx->d_rays.release();
    ctx->d_tans.release();
    ctx->d_path_xy.release();
    ctx->d_path_I.release();
    ctx->h_out.release();
    if (ctx->d_fail)
        cudaFree(ctx->d_fail);
    if (ctx->d_work)
        cudaFree(ctx->d_work);
    if (ctx->h_fail)
        cudaFreeHost(ctx->h_fail);
    for (auto e : ctx->ev)
        cudaEventDestroy(e);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream)
        cudaStreamDestroy(ctx->copy_stream);
    if (ctx->gv_ready)
        cudaEventDestroy(ctx->gv_ready);
    if (ctx->stage_done)
        cudaEventDestroy(ctx->stage_done);
    if (ctx->launch_done)
        cudaEventDestroy(ctx->launch_done);
    delete ctx;
}

const char *rtb200_last_error(const rtb200_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rtb200_stage(rtb200_ctx *ctx, const rtb200_problem *problem, unsigned flags)
{
    if (!ctx)
        return RTB200_ERR_ARG;
    int rc = validate(ctx, problem, flags);
    if (rc)
        return rc;
    reset_timing(ctx);
    new_event(ctx, ctx->stream);
    return stage_impl(ctx, problem, false, 0, 0.0, (flags & RTB200_FLAG_LAZY_TABLES) != 0);
}

int rtb200_staged_info(const rtb200_ctx *ctx, rtb200_staged *out)
{
    if (!ctx || !out || !ctx->staged)
        return RTB200_ERR_ARG;
    const DevProblem &P = ctx->prob;
    out->method = P.method;
    out->owner = ctx->owner_ok ? 1 : 0;
    out->snx = P.snx;
    out->sny = P.sny;
    out->nx = P.nx;
    out->ny = P.ny;
    out->na = P.na;
    out->nb = P.nb;
    out->nv = P.K;
    out->reserved = 0;
    return RTB200_OK;
}

int64_t rtb200_staged_pixels(const rtb200_ctx *ctx) { return ctx && ctx->staged ? ctx->staged_pixels : 0; }
int64_t rtb200_staged_rays(const rtb200_ctx *ctx) { return ctx && ctx->staged ? ctx->staged_rays : 0; }

int rtb200_launch(rtb200_ctx *ctx, int64_t pix_begin, int64_t pix_end, double *d_image,
                  double *d_I_ang, void *cuda_stream)
{
    if (!ctx || !ctx->staged || !d_image || !d_I_ang || pix_begin < 0 ||
        pix_end > ctx->staged_pixels) {
        if (ctx)
            ctx->err = "rtb200_launch: nothing staged or bad argument";
        return RTB200_ERR_ARG;
    }
    RTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : ctx->stream;
    if (st != ctx->stream) // order after the staging upload
        RTB_CUDA(cudaStreamWaitEvent(st, ctx->stage_done, 0));
    if (ctx->ev_used > 64 + 3 * 4096) // a long series of launches on one staging: keep the pool bounded
        reset_timing(ctx);
    Outputs out{ d_image, d_I_ang, nullptr, nullptr, ctx->d_fail, 0 };
    return launch_pixels(ctx, pix_begin, pix_end, out, st);
}

static int launch_rows_impl(rtb200_ctx *ctx, int row_offset, int row_stride, double *d_image,
                            double *d_I_ang, void *cuda_stream, bool compact)
{
    if (!ctx || !ctx->staged || !d_image || !d_I_ang || row_stride < 1 || row_offset < 0 ||
        row_offset >= row_stride) {
        if (ctx)
            ctx->err = "rtb200_launch_rows: nothing staged or bad argument";
        return RTB200_ERR_ARG;
    }
    const DevProblem &P = ctx->prob;
    if (compact && !(ctx->owner_ok)) {
        ctx->err = "rtb200_launch_rows_compact: the staged problem is not traced by pixel owners "
                   "(rtb200_staged_info().owner == 0): use rtb200_launch_rows and a sum";
        return RTB200_ERR_ARG;
    }
    RTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : ctx->stream;
    if (st != ctx->stream) // order after the staging upload
        RTB_CUDA(cudaStreamWaitEvent(st, ctx->stage_done, 0));
    if (ctx->ev_used > 64 + 3 * 4096)
        reset_timing(ctx);
    const long long rows = P.sny > row_offset ? (P.sny - row_offset + row_stride - 1) / row_stride : 0;
    Outputs out{ d_image, d_I_ang, nullptr, nullptr, ctx->d_fail, compact ? 1 : 0 };
    return launch_pixels(ctx, 0, rows * P.snx, out, st, row_offset, row_stride);
}

int rtb200_launch_rows(rtb200_ctx *ctx, int row_offset, int row_stride, double *d_image,
                       double *d_I_ang, void *cuda_stream)
{
    return launch_rows_impl(ctx, row_offset, row_stride, d_image, d_I_ang, cuda_stream, false);
}

int rtb200_launch_rows_compact(rtb200_ctx *ctx, int row_offset, int row_stride, double *d_rows,
                               double *d_I_ang, void *cuda_stream)
{
    return launch_rows_impl(ctx, row_offset, row_stride, d_rows, d_I_ang, cuda_stream, true);
}

int rtb200_unpermute_rows(rtb200_ctx *ctx, const double *d_gathered, int world, int64_t rows_per_dev,
                          double *d_image, void *cuda_stream)
{
    if (!ctx || !ctx->staged || !d_gathered || !d_image || world < 1 ||
        rows_per_dev < (ctx->prob.sny + world - 1) / world) {
        if (ctx)
            ctx->err = "rtb200_unpermute_rows: nothing staged or bad argument";
        return RTB200_ERR_ARG;
    }
    RTB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : ctx->stream;
    if (st != ctx->stream)
        RTB_CUDA(cudaStreamWaitEvent(st, ctx->stage_done, 0));
    launch_unpermute_rows(ctx->prob, d_gathered, world, rows_per_dev, d_image, st);
    RTB_CUDA(cudaGetLastError());
    return RTB200_OK;
}

int rtb200_sync(rtb200_ctx *ctx, unsigned *failure_code, rtb200_ray *failed, int max_failed,
                int *n_failed)
{
    if (!ctx)
        return RTB200_ERR_ARG;
    RTB_CUDA(cudaSetDevice(ctx->device));
    RTB_CUDA(cudaDeviceSynchronize());
    return finish(ctx, failure_code, failed, max_failed, n_failed);
}

int rtb200_reset_timings(rtb200_ctx *ctx)
{
    if (!ctx)
        return RTB200_ERR_ARG;
    reset_timing(ctx);
    return RTB200_OK;
}

int rtb200_get_timings(const rtb200_ctx *ctx, rtb200_timings *out)
{
    if (!ctx || !out)
        return RTB200_ERR_ARG;
    *out = ctx->last;
    return RTB200_OK;
}

int rtb200_create_image(rtb200_ctx *ctx, const rtb200_problem *problem, unsigned flags,
                        double *image, double *I_ang, unsigned *failure_code,
                        rtb200_ray *failed, int max_failed, int *n_failed)
{
    if (!ctx || !image || !I_ang) {
        if (ctx)
            ctx->err = "rtb200_create_image: NULL argument";
        return RTB200_ERR_ARG;
    }
    int rc = validate(ctx, problem, flags);
    if (rc)
        return rc;
    reset_timing(ctx);
    new_event(ctx, ctx->stream);
    rc = stage_impl(ctx, problem, false, 0, 0.0, true);
    if (rc)
        return rc;
    const DevProblem &P = ctx->prob;
    const size_t n_img = (size_t) P.nx * P.ny * P.K, n_ang = (size_t) P.na * P.nb;
    RTB_CUDA(ctx->d_image.reserve(n_img));
    RTB_CUDA(ctx->d_iang.reserve(n_ang));
    RTB_CUDA(cudaMemsetAsync(ctx->d_image.p, 0, n_img * sizeof(double), ctx->stream));
    RTB_CUDA(cudaMemsetAsync(ctx->d_iang.p, 0, n_ang * sizeof(double), ctx->stream));
    Outputs out{ ctx->d_image.p, ctx->d_iang.p, nullptr, nullptr, ctx->d_fail, 0 };
    rc = launch_pixels(ctx, 0, ctx->staged_pixels, out, ctx->stream);
    if (rc == RTB200_OK)
        rc = flush_gv(ctx, ctx->stream); // nothing was launched (no source pixels): finish the staging
    ctx->gv_pending = nullptr;
    if (rc) {
        ctx->staged = false; // the tables may not have been uploaded
        return rc;
    }
    ctx->ev_d2h.first = new_event(ctx, ctx->stream);
    RTB_CUDA(cudaMemcpyAsync(image, ctx->d_image.p, n_img * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(I_ang, ctx->d_iang.p, n_ang * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    ctx->ev_d2h.second = new_event(ctx, ctx->stream);
    ctx->have_d2h = true;
    return finish(ctx, failure_code, failed, max_failed, n_failed);
}

extern "C" int rtb200_parse_dat_view(const void *bytes, size_t n_bytes, rtb200_problem **problem);

int rtb200_create_image_from_dat(rtb200_ctx *ctx, const void *bytes, size_t n_bytes, unsigned flags,
                                 double *image, double *I_ang, unsigned *failure_code,
                                 rtb200_ray *failed, int max_failed, int *n_failed)
{
    if (!ctx || !bytes) {
        if (ctx)
            ctx->err = "rtb200_create_image_from_dat: NULL argument";
        return RTB200_ERR_ARG;
    }
    rtb200_problem *p = nullptr;
    int rc = rtb200_parse_dat_view(bytes, n_bytes, &p);
    if (rc != RTB200_OK) {
        ctx->err = "malformed .dat byte stream";
        return rc;
    }
    // the large arrays of `p` alias `bytes`: the packer reads them once, into the pinned blob
    rc = rtb200_create_image(ctx, p, flags, image, I_ang, failure_code, failed, max_failed, n_failed);
    rtb200_free_problem(p);
    return rc;
}

namespace {

// tanf(1e-3f*a) per ray through a small open-addressing cache keyed by the float's bits: ray
// lists enumerate a grid, so only na + nb distinct angles occur.
struct TanCache {
    static const int CAP = 8192;
    uint32_t key[CAP];
    float val[CAP];
    bool used[CAP];
    int count = 0;
    TanCache() { std::memset(used, 0, sizeof(used)); }
    float get(float a)
    {
        uint32_t bits;
        std::memcpy(&bits, &a, 4);
        uint32_t h = (bits * 2654435761u) >> 19; // 13 bits
        for (int probe = 0; probe < 16; probe++) {
            const uint32_t s = (h + probe) & (CAP - 1);
            if (used[s] && key[s] == bits)
                return val[s];
            if (!used[s]) {
                if (count > CAP / 2)
                    break;
                used[s] = true;
                key[s] = bits;
                val[s] = tanf(1e-3f * a);
                count++;
                return val[s];
            }
        }
        return tanf(1e-3f * a);
    }
};

int upload_rays(rtb200_ctx *ctx, const rtb200_ray *rays, size_t n)
{
    RTB_CUDA(ctx->d_rays.reserve(n));
    RTB_CUDA(ctx->d_tans.reserve(n));
    RTB_CUDA(ctx->h_out.reserve(n * sizeof(float2)));
    float2 *t = reinterpret_cast<float2 *>(ctx->h_out.p);
    TanCache cache;
    for (size_t i = 0; i < n; i++) {
        t[i].x = cache.get(rays[i].a);
        t[i].y = cache.get(rays[i].b);
    }
    RTB_CUDA(cudaMemcpyAsync(ctx->d_rays.p, rays, n * sizeof(float4), cudaMemcpyHostToDevice,
                             ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(ctx->d_tans.p, t, n * sizeof(float2), cudaMemcpyHostToDevice,
                             ctx->stream));
    return RTB200_OK;
}

// Marches + integrates an explicit ray list in chunks.
int launch_list(rtb200_ctx *ctx, size_t n_rays, const Outputs &out_all, bool keep_handoff)
{
    const DevProblem &P = ctx->prob;
    const int S = (P.N - 1) * RTB_N_SUB;
    const size_t per_slot = (size_t) std::max(S, 1) * sizeof(SegRec) + sizeof(unsigned) + sizeof(float4);
    long long per_chunk = keep_handoff ? (long long) n_rays
                                       : std::max<long long>(1, (long long) (ctx->handoff_bytes / per_slot));
    per_chunk = std::min<long long>(per_chunk, (long long) n_rays);
    const long long max_slots = ((1LL << 31) - 64) / std::max(S, 1); // slots and records of a chunk are counted in 32 bits
    if (keep_handoff && per_chunk > max_slots) {
        ctx->err = "more than 2^31 rays in one call that keeps the per-ray intermediates";
        return RTB200_ERR_LIMITS;
    }
    per_chunk = std::min(per_chunk, max_slots);
    int rc = ensure_handoff(ctx, per_chunk, true);
    if (rc)
        return rc;
    Handoff h{ ctx->d_seg.p, ctx->d_meta.p, ctx->d_exit.p, nullptr };
    for (long long a = 0; a < (long long) n_rays; a += per_chunk) {
        Chunk c;
        std::memset(&c, 0, sizeof(c));
        c.ray0 = a;
        c.ray1 = std::min<long long>((long long) n_rays, a + per_chunk);
        c.rays = ctx->d_rays.p;
        c.tans = ctx->d_tans.p;
        Outputs out = out_all;
        if (out.Iv)
            out.Iv += (size_t) a * P.K;
        if (out.error)
            out.error += a;
        const size_t e0 = new_event(ctx, ctx->stream);
        launch_march(P, c, true, h, ctx->d_fail, ctx->count_steps, ctx->stream, ctx->d_work, ctx->march_blocks);
        const size_t e1 = new_event(ctx, ctx->stream);
        launch_integrate_scatter(P, c, true, h, out, ctx->stream);
        const size_t e2 = new_event(ctx, ctx->stream);
        ctx->ev_march.push_back({ e0, e1 });
        ctx->ev_integ.push_back({ e1, e2 });
        ctx->launches += 2;
    }
    RTB_CUDA(cudaGetLastError());
    return RTB200_OK;
}

} // namespace

int rtb200_trace_rays(rtb200_ctx *ctx, int N, const rtb200_beam *beam,
                      const rtb200_gain_plane *gain, const rtb200_seed *seed, int method,
                      const rtb200_ray *rays, size_t n_rays, double scale, double *image,
                      double *I_ang, unsigned *failure_code, rtb200_ray *failed,
                      int max_failed, int *n_failed)
{
    if (!ctx || !beam || !gain || !image || !I_ang || (n_rays && !rays) ||
        (method != 1 && method != 2)) {
        if (ctx)
            ctx->err = "rtb200_trace_rays: bad argument";
        return RTB200_ERR_ARG;
    }
    rtb200_problem p;
    std::memset(&p, 0, sizeof(p));
    p.N = N;
    p.N_start = 0;
    p.N_parallel = 1;
    p.euv_beam = beam;
    p.gain = gain;
    p.seed = seed;
    p.seed_beam = seed ? beam : nullptr; // only consulted by the validation (grids are the euv ones)
    int rc = validate(ctx, &p, RTB200_FLAG_NO_LIMITS);
    p.seed_beam = nullptr;
    if (rc)
        return rc;
    reset_timing(ctx);
    new_event(ctx, ctx->stream);
    rc = stage_impl(ctx, &p, true, method, scale);
    if (rc)
        return rc;
    const DevProblem &P = ctx->prob;
    const size_t n_img = (size_t) P.nx * P.ny * P.K, n_ang = (size_t) P.na * P.nb;
    RTB_CUDA(ctx->d_image.reserve(n_img));
    RTB_CUDA(ctx->d_iang.reserve(n_ang));
    RTB_CUDA(cudaMemsetAsync(ctx->d_image.p, 0, n_img * sizeof(double), ctx->stream));
    RTB_CUDA(cudaMemsetAsync(ctx->d_iang.p, 0, n_ang * sizeof(double), ctx->stream));
    if (n_rays) {
        rc = upload_rays(ctx, rays, n_rays);
        if (rc)
            return rc;
        Outputs out{ ctx->d_image.p, ctx->d_iang.p, nullptr, nullptr, ctx->d_fail, 0 };
        rc = launch_list(ctx, n_rays, out, false);
        if (rc)
            return rc;
    }
    // accumulate into the caller's buffers, like the CPU loop (src/RayTraceImageCPU.cpp:56-68)
    std::vector<double> tmp(n_img + n_ang);
    ctx->ev_d2h.first = new_event(ctx, ctx->stream);
    RTB_CUDA(cudaMemcpyAsync(tmp.data(), ctx->d_image.p, n_img * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(tmp.data() + n_img, ctx->d_iang.p, n_ang * sizeof(double),
                             cudaMemcpyDeviceToHost, ctx->stream));
    ctx->ev_d2h.second = new_event(ctx, ctx->stream);
    ctx->have_d2h = true;
    rc = finish(ctx, failure_code, failed, max_failed, n_failed);
    if (rc < 0)
        return rc;
    for (size_t i = 0; i < n_img; i++)
        image[i] += tmp[i];
    for (size_t i = 0; i < n_ang; i++)
        I_ang[i] += tmp[n_img + i];
    return rc;
}

int rtb200_calc_rays(rtb200_ctx *ctx, int N, double dz, const rtb200_gain_plane *gain,
                     const rtb200_seed *seed, int K, int method, const rtb200_ray *rays,
                     size_t n_rays, double *Iv, rtb200_ray *ray2, int *error, float *gvl,
                     float *evl, int32_t *ivl)
{
    if (!ctx || !gain || (n_rays && !rays) || (method != 1 && method != 2) || K < 1) {
        if (ctx)
            ctx->err = "rtb200_calc_rays: bad argument";
        return RTB200_ERR_ARG;
    }
    rtb200_beam beam;
    std::memset(&beam, 0, sizeof(beam));
    beam.nv = K;
    beam.dz = dz;
    rtb200_problem p;
    std::memset(&p, 0, sizeof(p));
    p.N = N;
    p.N_parallel = 1;
    p.euv_beam = &beam;
    p.gain = gain;
    p.seed = seed;
    if (N < 1 || (N - 1) * RTB200_N_SUB > RTB_MAX_SEGS) {
        ctx->err = "rtb200_calc_rays: bad N";
        return RTB200_ERR_ARG;
    }
    for (int i = 0; i < N; i++)
        if (gain[i].Nx < 2 || gain[i].Ny < 2 || gain[i].Nv != K) {
            ctx->err = "rtb200_calc_rays: invalid gain plane";
            return RTB200_ERR_ARG;
        }
    if (int rs = validate_seed(ctx, seed, K))
        return rs;
    reset_timing(ctx);
    new_event(ctx, ctx->stream);
    int rc = stage_impl(ctx, &p, true, method, 1.0);
    if (rc)
        return rc;
    if (!n_rays)
        return RTB200_OK;
    const int S = (N - 1) * RTB_N_SUB;
    rc = upload_rays(ctx, rays, n_rays);
    if (rc)
        return rc;
    RTB_CUDA(ctx->d_Iv.reserve(n_rays * (size_t) K));
    RTB_CUDA(ctx->d_err.reserve(n_rays));
    Outputs out{ nullptr, nullptr, ctx->d_Iv.p, ctx->d_err.p, ctx->d_fail, 0 };
    rc = launch_list(ctx, n_rays, out, true);
    if (rc)
        return rc;
    std::vector<SegRec> seg((size_t) n_rays * (size_t) std::max(S, 1));
    std::vector<unsigned> meta(n_rays);
    std::vector<float4> ex(n_rays);
    std::vector<int> err(n_rays);
    if (Iv)
        RTB_CUDA(cudaMemcpyAsync(Iv, ctx->d_Iv.p, n_rays * (size_t) K * sizeof(double),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(err.data(), ctx->d_err.p, n_rays * sizeof(int), cudaMemcpyDeviceToHost,
                             ctx->stream));
    if (S > 0)
        RTB_CUDA(cudaMemcpyAsync(seg.data(), ctx->d_seg.p, n_rays * (size_t) S * sizeof(SegRec),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(meta.data(), ctx->d_meta.p, n_rays * sizeof(unsigned),
                             cudaMemcpyDeviceToHost, ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(ex.data(), ctx->d_exit.p, n_rays * sizeof(float4),
                             cudaMemcpyDeviceToHost, ctx->stream));
    rc = finish(ctx, nullptr, nullptr, 0, nullptr);
    if (rc < 0)
        return rc;
    for (size_t r = 0; r < n_rays; r++) {
        const int lo = meta[r] & 0xfff, hi = (meta[r] >> 12) & 0xfff;
        for (int s = 0; s < S; s++) {
            const bool in = s >= lo && s < hi;
            const SegRec &q = seg[r * (size_t) S + s];
            if (gvl)
                gvl[r * (size_t) S + s] = in ? q.gvl : 0.0f;
            if (evl)
                evl[r * (size_t) S + s] = in ? q.evl : 0.0f;
            if (ivl)
                ivl[r * (size_t) S + s] = in ? q.cell : 0;
        }
        if (error)
            error[r] = err[r];
        if (ray2) {
            const bool invalid = (meta[r] & RTB_META_INVALID) != 0;
            ray2[r].x = invalid ? 0.0f : ex[r].x;
            ray2[r].y = invalid ? 0.0f : ex[r].y;
            ray2[r].a = invalid ? 0.0f : ex[r].z;
            ray2[r].b = invalid ? 0.0f : ex[r].w;
        }
    }
    return RTB200_OK;
}

int rtb200_calc_ray_paths(rtb200_ctx *ctx, int N, double dz, const rtb200_gain_plane *gain,
                          const rtb200_seed *seed, int K, const double *dv, int method, double c,
                          const rtb200_ray *rays, size_t n_rays, float *xr, float *yr, float *Ir,
                          int *error)
{
    if (!ctx || !gain || !dv || (n_rays && !rays) || (method != 1 && method != 2) || K < 1 ||
        N < 1 || (N - 1) * RTB200_N_SUB > RTB_MAX_SEGS) {
        if (ctx)
            ctx->err = "rtb200_calc_ray_paths: bad argument";
        return RTB200_ERR_ARG;
    }
    if (K > 128) {
        ctx->err = "rtb200_calc_ray_paths: more than 128 frequencies";
        return RTB200_ERR_LIMITS;
    }
    for (int i = 0; i < N; i++)
        if (gain[i].Nx < 2 || gain[i].Ny < 2 || gain[i].Nv != K) {
            ctx->err = "rtb200_calc_ray_paths: invalid gain plane";
            return RTB200_ERR_ARG;
        }
    if (int rs = validate_seed(ctx, seed, K))
        return rs;
    rtb200_beam beam;
    std::memset(&beam, 0, sizeof(beam));
    beam.nv = K;
    beam.dz = dz;
    beam.dv = dv;
    rtb200_problem p;
    std::memset(&p, 0, sizeof(p));
    p.N = N;
    p.N_parallel = 1;
    p.euv_beam = &beam;
    p.gain = gain;
    p.seed = seed;
    reset_timing(ctx);
    new_event(ctx, ctx->stream);
    int rc = stage_impl(ctx, &p, true, method, 1.0);
    if (rc)
        return rc;
    ctx->prob.c = (float) c;
    if (!n_rays)
        return RTB200_OK;
    const int S = (N - 1) * RTB_N_SUB, N2 = S + 1;
    rc = upload_rays(ctx, rays, n_rays);
    if (rc)
        return rc;
    rc = ensure_handoff(ctx, (long long) n_rays, true);
    if (rc)
        return rc;
    RTB_CUDA(ctx->d_path_xy.reserve(n_rays * (size_t) N2));
    RTB_CUDA(ctx->d_path_I.reserve(n_rays * (size_t) N2));
    RTB_CUDA(ctx->d_err.reserve(n_rays));
    RTB_CUDA(cudaMemsetAsync(ctx->d_path_xy.p, 0, n_rays * (size_t) N2 * sizeof(float2), ctx->stream));
    RTB_CUDA(cudaMemsetAsync(ctx->d_path_I.p, 0, n_rays * (size_t) N2 * sizeof(float), ctx->stream));
    Handoff h{ ctx->d_seg.p, ctx->d_meta.p, ctx->d_exit.p, ctx->d_path_xy.p };
    Chunk ck;
    std::memset(&ck, 0, sizeof(ck));
    ck.ray1 = (long long) n_rays;
    ck.rays = ctx->d_rays.p;
    ck.tans = ctx->d_tans.p;
    launch_march(ctx->prob, ck, true, h, ctx->d_fail, false, ctx->stream, ctx->d_work, ctx->march_blocks);
    launch_path_intensity(ctx->prob, ck, h, ctx->d_path_I.p, ctx->d_err.p, ctx->stream);
    ctx->launches += 2;
    RTB_CUDA(cudaGetLastError());
    std::vector<float2> xy(n_rays * (size_t) N2);
    std::vector<int> err(n_rays);
    RTB_CUDA(cudaMemcpyAsync(xy.data(), ctx->d_path_xy.p, xy.size() * sizeof(float2),
                             cudaMemcpyDeviceToHost, ctx->stream));
    if (Ir)
        RTB_CUDA(cudaMemcpyAsync(Ir, ctx->d_path_I.p, n_rays * (size_t) N2 * sizeof(float),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    RTB_CUDA(cudaMemcpyAsync(err.data(), ctx->d_err.p, n_rays * sizeof(int), cudaMemcpyDeviceToHost,
                             ctx->stream));
    rc = finish(ctx, nullptr, nullptr, 0, nullptr);
    if (rc < 0)
        return rc;
    int n_errors = 0;
    for (size_t r = 0; r < n_rays; r++) {
        for (int q = 0; q < N2; q++) {
            if (xr)
                xr[r * (size_t) N2 + q] = xy[r * (size_t) N2 + q].x;
            if (yr)
                yr[r * (size_t) N2 + q] = xy[r * (size_t) N2 + q].y;
        }
        if (error)
            error[r] = err[r];
        n_errors += err[r] != 0;
    }
    return n_errors ? RTB200_RAYS_FAILED : RTB200_OK;
}

int rtb200_measure_fp64_peak(rtb200_ctx *ctx, double *rate)
{
    if (!ctx || !rate)
        return RTB200_ERR_ARG;
    RTB_CUDA(cudaSetDevice(ctx->device));
    RTB_CUDA(ctx->d_Iv.reserve(16));
    int blocks = 0, threads = 0;
    const int iters = 4096;
    launch_fp64_peak(ctx->d_Iv.p, 64, ctx->stream, &blocks, &threads); // warm-up
    cudaEvent_t a, b;
    RTB_CUDA(cudaEventCreate(&a));
    RTB_CUDA(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        RTB_CUDA(cudaEventRecord(a, ctx->stream));
        launch_fp64_peak(ctx->d_Iv.p, iters, ctx->stream, &blocks, &threads);
        RTB_CUDA(cudaEventRecord(b, ctx->stream));
        RTB_CUDA(cudaEventSynchronize(b));
        float ms = 0;
        RTB_CUDA(cudaEventElapsedTime(&ms, a, b));
        const double r = (double) blocks * threads * (double) iters * 8.0 / (ms * 1e-3);
        best = std::max(best, r);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *rate = best;
    return RTB200_OK;
}

int rtb200_check_fdiv(rtb200_ctx *ctx, unsigned b_first, unsigned b_count, int exp_a, int exp_b,
                      int variant, unsigned long long *mismatches, float *witness_a, float *witness_b)
{
    if (!ctx || !mismatches || b_first >= (1u << 23) || b_count > (1u << 23) - b_first ||
        exp_a < -60 || exp_a > 59 || exp_b < -60 || exp_b > 59)
        return RTB200_ERR_ARG;
    RTB_CUDA(cudaSetDevice(ctx->device));
    RTB_CUDA(ctx->d_Iv.reserve(16));
    unsigned long long *d = reinterpret_cast<unsigned long long *>(ctx->d_Iv.p);
    RTB_CUDA(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), ctx->stream));
    launch_fdiv_check(b_first, b_count, exp_a, exp_b, variant, d, ctx->stream);
    unsigned long long h[3] = { 0, 0, 0 };
    RTB_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    RTB_CUDA(cudaStreamSynchronize(ctx->stream));
    RTB_CUDA(cudaGetLastError());
    *mismatches = h[0];
    const unsigned ab = (unsigned) h[1], bb = (unsigned) h[2];
    if (witness_a)
        std::memcpy(witness_a, &ab, 4);
    if (witness_b)
        std::memcpy(witness_b, &bb, 4);
    return RTB200_OK;
}

} // extern "C"
