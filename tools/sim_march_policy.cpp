// sim_march_policy.cpp — ANALYSIS TOOL (not part of librtb200.so): runs the march's own source
// (rtb200_march_flat.cuh, host build) over a sample of ray slots, records for every ray which of
// the three blocks of a trip (CELL, INTERP, STEP) it needs in which order, and replays warps of
// 32 lanes under different block-scheduling policies to count the warp instructions each policy
// would issue.  Used to decide, before touching the kernel, whether gating the low-occupancy
// blocks pays (tools/sim_march_policy.py drives it).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../raytrace-miniapp_b200/csrc/rtb200_fp64.cuh"
#include "../raytrace-miniapp_b200/csrc/rtb200_march_flat.cuh"
#include "../raytrace-miniapp_b200/csrc/rtb200_pack.h"

using namespace rtb;

namespace {
struct NullSink {
    void operator()(int, float, float, int) const {}
    void point(int, float, float) const {}
};
} // namespace

extern "C" {

// Ops of rays [first, first + count) of the grid enumeration in pixel-major slot order (slot =
// pixel * AB + ab, the order the kernel hands them out): ops[off[r] .. off[r+1]) are bytes
// 'C', 'I', 'S'.  Returns the total number of ops or -1.
long long sim_trace(const rtb200_problem *p, long long first, long long count, unsigned char *ops,
                    long long cap, long long *off)
{
    DevProblem P;
    const size_t bytes = pack_problem(*p, false, 0, 0.0, nullptr, nullptr, P);
    std::vector<char> blob(bytes + 256);
    char *base = (char *) (((uintptr_t) blob.data() + 255) & ~(uintptr_t) 255);
    pack_problem(*p, false, 0, 0.0, base, base, P);
    float zt[RTB_N_SUB];
    for (int iz = 0; iz < RTB_N_SUB; iz++)
        zt[iz] = march_sub_limit(iz, P.dz0);
    MarchConsts K;
    march_consts(K, P.lite, zt, P.N, P.method, P.c, P.use_emis != 0);
    const long long AB = (long long) P.sna * P.snb;
    long long n = 0;
    NullSink sink;
    for (long long r = 0; r < count; r++) {
        const long long slot = first + r;
        const long long pix = slot / AB, ab = slot % AB; // pixel p = i + j*snx
        const int i = (int) (pix % P.snx), j = (int) (pix / P.snx);
        const int k = (int) (ab / P.snb), m = (int) (ab % P.snb);
        off[r] = n;
        if (j >= P.sny)
            return -1;
        FlatMarch fm;
        flat_init(fm, K, P.sxf[i], P.syf[j], P.tanA[k], P.tanB[m]);
        while (flat_phase(fm) != PH_DONE) {
            if (n + 3 > cap)
                return -1;
            if (flat_phase(fm) == PH_CELL) {
                ops[n++] = 'C';
                flat_cell(fm, K, sink);
            }
            if (flat_phase(fm) == PH_INTERP) {
                ops[n++] = 'I';
                flat_interp(fm, K, sink);
            }
            if (flat_phase(fm) == PH_STEP) {
                ops[n++] = 'S';
                flat_step(fm, K);
            }
        }
    }
    off[count] = n;
    return n;
}

// Replays the rays as warps: warp w owns slots [w*per_warp, (w+1)*per_warp), lanes are refilled
// `refill_min` at a time like the kernel does.  policy 0: every block runs whenever a lane needs
// it (the kernel today).  policy 1: CELL only runs when at least thrC lanes wait for it, or no
// lane can do anything else, or a lane has waited maxwait trips.  policy 2: additionally INTERP is
// gated by thrI the same way.  out[0..] = trips, execC, execI, execS, lanesC, lanesI, lanesS,
// instructions.
void sim_policy(const unsigned char *ops, const long long *off, long long n_rays, int per_warp, int refill_min,
                int policy, int thrC, int thrI, int maxwait, const double *cost, double *out)
{
    double trips = 0, eC = 0, eI = 0, eS = 0, lC = 0, lI = 0, lS = 0, instr = 0;
    const long long n_warps = n_rays / per_warp;
    for (long long w = 0; w < n_warps; w++) {
        long long next = w * per_warp, end = next + per_warp;
        long long pos[32], stop[32];
        int waited[32];
        bool have[32];
        for (int l = 0; l < 32; l++)
            have[l] = false, pos[l] = stop[l] = 0, waited[l] = 0;
        for (;;) {
            int idle = 0, live = 0;
            for (int l = 0; l < 32; l++) {
                if (have[l] && pos[l] >= stop[l])
                    have[l] = false;
                if (!have[l])
                    idle++;
                else
                    live++;
            }
            if (idle > 0 && next < end && (idle >= refill_min || live == 0)) {
                for (int l = 0; l < 32 && next < end; l++)
                    if (!have[l]) {
                        pos[l] = off[next], stop[l] = off[next + 1];
                        have[l] = pos[l] < stop[l];
                        waited[l] = 0;
                        next++;
                        if (have[l])
                            live++;
                    }
                instr += cost[4]; // refill block
            }
            if (live == 0) {
                if (next >= end)
                    break;
                continue;
            }
            trips += 1;
            instr += cost[3]; // loop overhead
            int nC = 0, nI = 0, nS = 0, wmax = 0;
            for (int l = 0; l < 32; l++)
                if (have[l]) {
                    const unsigned char o = ops[pos[l]];
                    nC += o == 'C', nI += o == 'I', nS += o == 'S';
                    if (o == 'C' && waited[l] > wmax)
                        wmax = waited[l];
                }
            bool runC = nC > 0;
            if (policy >= 1 && nC > 0)
                runC = nC >= thrC || (nI + nS) == 0 || wmax >= maxwait;
            if (runC) {
                eC += 1, lC += nC, instr += cost[0];
                for (int l = 0; l < 32; l++)
                    if (have[l] && ops[pos[l]] == 'C')
                        pos[l]++, waited[l] = 0;
            } else {
                for (int l = 0; l < 32; l++)
                    if (have[l] && ops[pos[l]] == 'C')
                        waited[l]++;
            }
            nI = 0;
            nS = 0;
            for (int l = 0; l < 32; l++)
                if (have[l] && pos[l] < stop[l]) {
                    nI += ops[pos[l]] == 'I';
                    nS += ops[pos[l]] == 'S';
                }
            bool runI = nI > 0;
            if (policy >= 2 && nI > 0)
                runI = nI >= thrI || nS == 0;
            if (runI) {
                eI += 1, lI += nI, instr += cost[1];
                for (int l = 0; l < 32; l++)
                    if (have[l] && pos[l] < stop[l] && ops[pos[l]] == 'I')
                        pos[l]++;
            }
            nS = 0;
            for (int l = 0; l < 32; l++)
                if (have[l] && pos[l] < stop[l] && ops[pos[l]] == 'S')
                    nS++;
            if (nS > 0) {
                eS += 1, lS += nS, instr += cost[2];
                for (int l = 0; l < 32; l++)
                    if (have[l] && pos[l] < stop[l] && ops[pos[l]] == 'S')
                        pos[l]++;
            }
        }
    }
    out[0] = trips, out[1] = eC, out[2] = eI, out[3] = eS, out[4] = lC, out[5] = lI, out[6] = lS, out[7] = instr;
}

} // extern "C"
