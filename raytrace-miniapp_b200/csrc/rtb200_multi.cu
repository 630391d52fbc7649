// rtb200_multi.cu — one create_image call on several devices of one box, behind the C ABI.
//
// Single process, one rtb200 context per device, one NCCL communicator over them
// (ncclCommInitAll).  The image is sharded by source-pixel ROWS, row-cyclically (device r of W
// traces rows r, r + W, ...: balanced, rows near the target surface escape early); every device
// packs and uploads the problem itself, on its own host thread, so staging runs in parallel.
// Exchange step, all on the devices' streams, then ONE download from device 0:
//   ASE (pixel-owner kernel) : each device's rows, compacted, sent to device 0 (grouped
//                              ncclSend / ncclRecv over NVLink), un-permuted there into the image;
//   seeded / scatter binning : ncclReduce(sum) of the full-size partial images;
//   I_ang                    : ncclReduce(sum) of na*nb doubles.
// This is the reference's `cuda-multigpu` method (src/RayTraceImage.cpp:396-405: ThreadLoop, one
// worker per GPU, partial images summed on the host, and - because cudaSetDevice is called in
// the parent thread, :116-119 - every worker on device 0) and the application's
// intensity_step_struct::sum_reduce (src/RayTraceStructures.cpp:1603-1646), restated for
// device-resident partials.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): librtb200.so has no link-time dependency
// on it, and a single-device rtb200_multi needs no NCCL at all.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtb200.h"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t,
                           cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err)
    {
        if (lib)
            return true;
        const char *names[] = { getenv("RTB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char *n : names) {
            if (n && (lib = dlopen(n, RTLD_NOW | RTLD_LOCAL)))
                break;
        }
        if (!lib) {
            err = std::string("NCCL not found (libnccl.so.2): ") + dlerror();
            return false;
        }
#define RTB_NCCL_SYM(f)                                                                       \
    do {                                                                                      \
        *(void **) (&f) = dlsym(lib, "nccl" #f);                                              \
        if (!f) {                                                                             \
            err = "symbol nccl" #f " missing in the NCCL library";                            \
            return false;                                                                     \
        }                                                                                     \
    } while (0)
        RTB_NCCL_SYM(CommInitAll);
        RTB_NCCL_SYM(CommDestroy);
        RTB_NCCL_SYM(GroupStart);
        RTB_NCCL_SYM(GroupEnd);
        RTB_NCCL_SYM(Send);
        RTB_NCCL_SYM(Recv);
        RTB_NCCL_SYM(Reduce);
        RTB_NCCL_SYM(GetErrorString);
#undef RTB_NCCL_SYM
        return true;
    }
};

struct DevBuf {
    double *p = nullptr;
    size_t cap = 0; // doubles
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **) &p, (n + n / 8 + 64) * sizeof(double));
        if (e == cudaSuccess)
            cap = n + n / 8 + 64;
        return e;
    }
};

} // namespace

struct rtb200_multi {
    int n = 0;
    std::vector<int> dev;
    std::vector<rtb200_ctx *> ctx;
    std::vector<cudaStream_t> st;
    std::vector<ncclComm_t> comm;
    std::vector<DevBuf> part, iang; // per device: its rows (or its full-size partial image), its I_ang share
    DevBuf gather, image, iang_sum;  // device 0
    std::vector<cudaEvent_t> ev;     // device 0: [0] start, [1] kernels done, [2] exchange done, [3] download done
    std::vector<rtb200_timings> tm;
    float exchange_ms = 0.f, total_ms = 0.f;
    NcclApi nccl;
    std::string err;
};

#define RTBM_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            m->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            return RTB200_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)
#define RTBM_NCCL(call)                                                                        \
    do {                                                                                       \
        ncclResult_t r_ = (call);                                                              \
        if (r_ != ncclSuccess) {                                                               \
            m->err = std::string(#call) + ": " + m->nccl.GetErrorString(r_);                   \
            return RTB200_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)

extern "C" {

void rtb200_multi_destroy(rtb200_multi *m)
{
    if (!m)
        return;
    for (int i = 0; i < (int) m->ctx.size(); i++) {
        cudaSetDevice(m->dev[i]);
        if (i < (int) m->st.size() && m->st[i])
            cudaStreamSynchronize(m->st[i]);
        if (i < (int) m->comm.size() && m->comm[i] && m->nccl.CommDestroy)
            m->nccl.CommDestroy(m->comm[i]);
        if (i < (int) m->part.size() && m->part[i].p)
            cudaFree(m->part[i].p);
        if (i < (int) m->iang.size() && m->iang[i].p)
            cudaFree(m->iang[i].p);
        if (i == 0) {
            for (DevBuf *b : { &m->gather, &m->image, &m->iang_sum })
                if (b->p)
                    cudaFree(b->p);
            for (cudaEvent_t e : m->ev)
                cudaEventDestroy(e);
        }
        if (i < (int) m->st.size() && m->st[i])
            cudaStreamDestroy(m->st[i]);
        rtb200_destroy(m->ctx[i]);
    }
    delete m;
}

int rtb200_multi_create(const int *devices, int n_dev, rtb200_multi **out)
{
    if (!out || n_dev < 1)
        return RTB200_ERR_ARG;
    *out = nullptr;
    if (n_dev > rtb200_device_count())
        return RTB200_ERR_CUDA;
    rtb200_multi *m = new rtb200_multi;
    m->n = n_dev;
    for (int i = 0; i < n_dev; i++)
        m->dev.push_back(devices ? devices[i] : i);
    m->part.resize(n_dev);
    m->iang.resize(n_dev);
    m->tm.resize(n_dev);
    for (int i = 0; i < n_dev; i++) {
        rtb200_ctx *c = nullptr;
        if (rtb200_create(m->dev[i], &c) != RTB200_OK) {
            rtb200_multi_destroy(m);
            return RTB200_ERR_CUDA;
        }
        m->ctx.push_back(c);
        cudaStream_t s = nullptr;
        if (cudaSetDevice(m->dev[i]) != cudaSuccess ||
            cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
            rtb200_multi_destroy(m);
            return RTB200_ERR_CUDA;
        }
        m->st.push_back(s);
    }
    cudaSetDevice(m->dev[0]);
    for (int i = 0; i < 4; i++) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) {
            rtb200_multi_destroy(m);
            return RTB200_ERR_CUDA;
        }
        m->ev.push_back(e);
    }
    if (n_dev > 1) {
        m->comm.assign(n_dev, nullptr);
        if (!m->nccl.load(m->err) ||
            m->nccl.CommInitAll(m->comm.data(), n_dev, m->dev.data()) != ncclSuccess) {
            // keep the object alive so that the caller can read the reason
            if (m->err.empty())
                m->err = "ncclCommInitAll failed";
            m->comm.clear();
            *out = m;
            return RTB200_ERR_CUDA;
        }
    }
    *out = m;
    return RTB200_OK;
}

const char *rtb200_multi_last_error(const rtb200_multi *m) { return m ? m->err.c_str() : "null object"; }
int rtb200_multi_device_count(const rtb200_multi *m) { return m ? m->n : 0; }

int rtb200_multi_create_image(rtb200_multi *m, const rtb200_problem *problem, unsigned flags,
                              double *image, double *I_ang, unsigned *failure_code,
                              rtb200_ray *failed, int max_failed, int *n_failed)
{
    if (!m || !problem || !image || !I_ang) {
        if (m)
            m->err = "rtb200_multi_create_image: NULL argument";
        return RTB200_ERR_ARG;
    }
    const int W = m->n;
    if (W > 1 && m->comm.empty()) {
        m->err = "no NCCL communicator (rtb200_multi_create failed to set one up)";
        return RTB200_ERR_CUDA;
    }
    RTBM_CUDA(cudaSetDevice(m->dev[0]));
    RTBM_CUDA(cudaEventRecord(m->ev[0], m->st[0]));

    // ---- every device stages the problem and traces its rows, one host thread each ----------
    std::vector<int> rc(W, RTB200_OK);
    std::vector<std::string> errs(W);
    rtb200_staged info;
    std::memset(&info, 0, sizeof(info));
    auto worker = [&](int r) {
        rtb200_ctx *c = m->ctx[r];
        auto fail = [&](int code, const char *what) {
            rc[r] = code;
            errs[r] = what ? what : rtb200_last_error(c);
        };
        if (cudaSetDevice(m->dev[r]) != cudaSuccess)
            return fail(RTB200_ERR_CUDA, "cudaSetDevice");
        int e = rtb200_stage(c, problem, flags | RTB200_FLAG_LAZY_TABLES);
        if (e != RTB200_OK)
            return fail(e, nullptr);
        rtb200_staged s;
        rtb200_staged_info(c, &s);
        if (r == 0)
            info = s;
        const size_t n_ang = (size_t) s.na * s.nb;
        const size_t rows_per_dev = ((size_t) s.sny + W - 1) / W;
        const size_t n_part = s.owner ? rows_per_dev * (size_t) s.snx * s.nv : (size_t) s.nx * s.ny * s.nv;
        if (m->part[r].reserve(n_part) != cudaSuccess || m->iang[r].reserve(n_ang) != cudaSuccess)
            return fail(RTB200_ERR_CUDA, "cudaMalloc (partial results)");
        cudaMemsetAsync(m->iang[r].p, 0, n_ang * sizeof(double), m->st[r]);
        if (!s.owner) // scatter binning adds into a full-size partial image
            cudaMemsetAsync(m->part[r].p, 0, n_part * sizeof(double), m->st[r]);
        e = s.owner ? rtb200_launch_rows_compact(c, r, W, m->part[r].p, m->iang[r].p, m->st[r])
                    : rtb200_launch_rows(c, r, W, m->part[r].p, m->iang[r].p, m->st[r]);
        if (e != RTB200_OK)
            return fail(e, nullptr);
    };
    if (W == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int r = 0; r < W; r++)
            th.emplace_back(worker, r);
        for (auto &t : th)
            t.join();
    }
    for (int r = 0; r < W; r++)
        if (rc[r] != RTB200_OK) {
            m->err = "device " + std::to_string(m->dev[r]) + ": " + errs[r];
            for (int q = 0; q < W; q++) { // let whatever was launched drain
                cudaSetDevice(m->dev[q]);
                cudaStreamSynchronize(m->st[q]);
            }
            return rc[r];
        }

    // ---- exchange: rows / partial images and I_ang to device 0 over NCCL -------------------
    const size_t n_img = (size_t) info.nx * info.ny * info.nv, n_ang = (size_t) info.na * info.nb;
    const size_t rows_per_dev = ((size_t) info.sny + W - 1) / W;
    const size_t n_rows = rows_per_dev * (size_t) info.snx * info.nv;
    RTBM_CUDA(cudaSetDevice(m->dev[0]));
    RTBM_CUDA(m->image.reserve(n_img));
    RTBM_CUDA(m->iang_sum.reserve(n_ang));
    RTBM_CUDA(cudaEventRecord(m->ev[1], m->st[0])); // (device 0's kernels; the others overlap the exchange)
    const double *img_src = nullptr, *ang_src = nullptr;
    if (W == 1) {
        if (info.owner) {
            RTBM_CUDA(cudaMemsetAsync(m->image.p, 0, n_img * sizeof(double), m->st[0]));
            int e = rtb200_unpermute_rows(m->ctx[0], m->part[0].p, 1, (int64_t) rows_per_dev, m->image.p, m->st[0]);
            if (e != RTB200_OK) {
                m->err = rtb200_last_error(m->ctx[0]);
                return e;
            }
            img_src = m->image.p;
        } else {
            img_src = m->part[0].p;
        }
        ang_src = m->iang[0].p;
    } else {
        if (info.owner) {
            RTBM_CUDA(m->gather.reserve(n_rows * W));
            RTBM_CUDA(cudaMemsetAsync(m->image.p, 0, n_img * sizeof(double), m->st[0]));
        }
        RTBM_NCCL(m->nccl.GroupStart());
        for (int r = 0; r < W; r++) {
            if (info.owner) {
                RTBM_NCCL(m->nccl.Send(m->part[r].p, n_rows, ncclDouble, 0, m->comm[r], m->st[r]));
                RTBM_NCCL(m->nccl.Recv(m->gather.p + (size_t) r * n_rows, n_rows, ncclDouble, r, m->comm[0], m->st[0]));
            } else {
                RTBM_NCCL(m->nccl.Reduce(m->part[r].p, m->image.p, n_img, ncclDouble, ncclSum, 0, m->comm[r], m->st[r]));
            }
        }
        RTBM_NCCL(m->nccl.GroupEnd());
        RTBM_NCCL(m->nccl.GroupStart());
        for (int r = 0; r < W; r++)
            RTBM_NCCL(m->nccl.Reduce(m->iang[r].p, m->iang_sum.p, n_ang, ncclDouble, ncclSum, 0, m->comm[r], m->st[r]));
        RTBM_NCCL(m->nccl.GroupEnd());
        RTBM_CUDA(cudaSetDevice(m->dev[0]));
        if (info.owner) {
            int e = rtb200_unpermute_rows(m->ctx[0], m->gather.p, W, (int64_t) rows_per_dev, m->image.p, m->st[0]);
            if (e != RTB200_OK) {
                m->err = rtb200_last_error(m->ctx[0]);
                return e;
            }
        }
        img_src = m->image.p;
        ang_src = m->iang_sum.p;
    }
    RTBM_CUDA(cudaEventRecord(m->ev[2], m->st[0]));
    // ---- one download -----------------------------------------------------------------------
    RTBM_CUDA(cudaMemcpyAsync(image, img_src, n_img * sizeof(double), cudaMemcpyDeviceToHost, m->st[0]));
    RTBM_CUDA(cudaMemcpyAsync(I_ang, ang_src, n_ang * sizeof(double), cudaMemcpyDeviceToHost, m->st[0]));
    RTBM_CUDA(cudaEventRecord(m->ev[3], m->st[0]));

    // ---- failure reports of all devices (src/RayTraceImage.cpp:427-430) ----------------------
    unsigned fc = 0;
    int nf = 0, result = RTB200_OK;
    for (int r = 0; r < W; r++) {
        RTBM_CUDA(cudaSetDevice(m->dev[r]));
        RTBM_CUDA(cudaStreamSynchronize(m->st[r]));
        unsigned f = 0;
        int k = 0;
        rtb200_ray tmp[RTB200_N_FAILED_MAX];
        int e = rtb200_sync(m->ctx[r], &f, tmp, RTB200_N_FAILED_MAX, &k);
        if (e < 0) {
            m->err = "device " + std::to_string(m->dev[r]) + ": " + rtb200_last_error(m->ctx[r]);
            return e;
        }
        rtb200_get_timings(m->ctx[r], &m->tm[r]);
        fc |= f;
        for (int q = 0; q < std::min(k, RTB200_N_FAILED_MAX); q++)
            if (failed && nf + q < max_failed)
                failed[nf + q] = tmp[q];
        nf += k;
        if (e == RTB200_RAYS_FAILED)
            result = RTB200_RAYS_FAILED;
    }
    RTBM_CUDA(cudaSetDevice(m->dev[0]));
    RTBM_CUDA(cudaEventSynchronize(m->ev[3]));
    cudaEventElapsedTime(&m->exchange_ms, m->ev[1], m->ev[2]);
    cudaEventElapsedTime(&m->total_ms, m->ev[0], m->ev[3]);
    if (failure_code)
        *failure_code = fc;
    if (n_failed)
        *n_failed = nf;
    return result;
}

int rtb200_multi_get_timings(const rtb200_multi *m, rtb200_timings *per_device, float *exchange_ms,
                             float *total_ms)
{
    if (!m)
        return RTB200_ERR_ARG;
    if (per_device)
        for (int r = 0; r < m->n; r++)
            per_device[r] = m->tm[r];
    if (exchange_ms)
        *exchange_ms = m->exchange_ms;
    if (total_ms)
        *total_ms = m->total_ms;
    return RTB200_OK;
}

} // extern "C"
