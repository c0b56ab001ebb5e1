// Persistent-lane Devroye kernel (the PG(1,z) / sum-of-PG(1) hot path).
//
// Replaces rpg_devroye's serial loop (LogitWrapper.cpp:75-80) and Logit::draw_w
// (Logit.hpp:285-288).  Rejection sampling makes lanes finish at different times;
// instead of letting a warp spin until its slowest lane accepts, every lane is a
// small state machine: one proposal + series test per trip, and a lane whose draw
// is complete immediately pulls the next observation of its warp's chunk through a
// ballot-compacted refill (popc of the lower-lane mask gives each requesting lane
// its offset).  All 32 lanes therefore carry live work on every trip.  Because the
// variate stream is keyed by the observation index (philox.cuh), the result does
// not depend on which lane ends up drawing which observation.
//
// Work split: the batch is cut into chunks of kChunkObs observations dealt
// round-robin to warps, so a drift of z along the array cannot unbalance the SMs.
// HBM traffic: z and n in, omega out -- coalesced in runs of consecutive indices.
#include "engine.h"
#include "pg_devroye_fast.cuh"

namespace bl {

namespace {

constexpr int kThreads = 256;
constexpr int kChunkObs = 128;

__global__ void __launch_bounds__(kThreads)
k_devroye_refill(double *__restrict__ x, const int *__restrict__ n, const double *__restrict__ z,
                 int64_t num, StreamId id)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (kThreads / 32) * kChunkObs;

    int64_t cur = warp * kChunkObs;       // next unassigned observation of this warp (uniform)
    int64_t cend = cur + kChunkObs;       // end of the current chunk (uniform)

    bool active = false;
    int64_t obs = 0;
    int remaining = 0;
    double sum = 0.0;
    DevSetup st;
    PhiloxSource src;

    for (;;) {
        unsigned want = __ballot_sync(full, !active);
        if (want && cur < num) {
            int rank = __popc(want & lt_mask);
            int64_t cand = cur + rank;
            if (cand >= cend) cand += stride - kChunkObs;
            if (!active && cand < num) {
                int ni = n[cand];
                if (ni == 0) {
                    x[cand] = 0.0;                    // LogitWrapper.cpp:76-79
                } else {
                    obs = cand;
                    remaining = ni < 1 ? 1 : ni;      // NTHROW clamp, PolyaGamma.cpp:128-135
                    sum = 0.0;
                    st = dev_setup(z[cand]);
                    src.open(id.seed, id.obs0 + (uint64_t)cand, id.call_id);
                    active = true;
                }
            }
            cur += __popc(want);
            if (cur >= cend) {
                int64_t over = cur - cend;
                cend += stride;
                cur = cend - kChunkObs + over;
            }
        }
        if (!__any_sync(full, active)) {
            if (cur >= num) break;
            continue;
        }
        if (active) {
            double X;
            if (dev_propose(src, st, X)) {
                sum += 0.25 * X;
                if (--remaining == 0) {
                    x[obs] = sum;
                    active = false;
                }
            }
        }
    }
}

}  // namespace

cudaError_t launch_devroye_refill(double *x, const int *n, const double *z, int64_t num,
                                  StreamId id, cudaStream_t st)
{
    if (num <= 0) return cudaSuccess;
    int64_t chunks = (num + kChunkObs - 1) / kChunkObs;
    int64_t blocks = (chunks + (kThreads / 32) - 1) / (kThreads / 32);
    int64_t cap = 148LL * 4;                 // 148 SMs x resident CTAs
    int grid = (int)(blocks < cap ? blocks : cap);
    k_devroye_refill<<<grid, kThreads, 0, st>>>(x, n, z, num, id);
    count_launch();
    return cudaGetLastError();
}

}  // namespace bl
