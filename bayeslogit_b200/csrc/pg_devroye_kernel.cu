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

// Branch-class binning (large batches): the left proposal is the inverse-chi^2 pair loop when
// Z = |z|/2 < 1/0.64 and the inverse-Gaussian loop otherwise (PolyaGamma.cpp:87), a property of
// the observation alone.  A counting sort of the observation indices by that class makes every
// warp's chunk class-pure, so a warp runs two proposal branches instead of three.
constexpr int kBinThreads = 256;

__device__ __forceinline__ int dev_class(double z) { return fabs(z) * 0.5 >= 1.0 / kTrunc ? 1 : 0; }

__global__ void __launch_bounds__(kBinThreads)
k_cls_count(const double *__restrict__ z, int n, int *__restrict__ meta)
{
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += dev_class(z[i]);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0 && tot) atomicAdd(&meta[0], tot);
}

__global__ void __launch_bounds__(kBinThreads)
k_cls_scatter(const double *__restrict__ z, int n, int *__restrict__ meta, int *__restrict__ idx)
{
    __shared__ int wcnt[kBinThreads / 32][2];
    __shared__ int base[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int off1 = n - meta[0];                                  // class 1 starts after all of class 0
    int tiles = (n + kBinThreads - 1) / kBinThreads;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int i = tile * kBinThreads + threadIdx.x;
        int cls = i < n ? dev_class(z[i]) : -1;
        unsigned m1 = __ballot_sync(0xffffffffu, cls == 1), m0 = __ballot_sync(0xffffffffu, cls == 0);
        int rank = __popc((cls == 1 ? m1 : m0) & lt);
        if (lane == 0) { wcnt[warp][0] = __popc(m0); wcnt[warp][1] = __popc(m1); }
        __syncthreads();
        if (threadIdx.x < 2) {
            int t = 0;
            for (int w = 0; w < kBinThreads / 32; ++w) t += wcnt[w][threadIdx.x];
            base[threadIdx.x] = (threadIdx.x ? off1 : 0) + (t ? atomicAdd(&meta[1 + threadIdx.x], t) : 0);
        }
        __syncthreads();
        if (cls >= 0) {
            int off = base[cls];
            for (int w = 0; w < warp; ++w) off += wcnt[w][cls];
            idx[off + rank] = i;
        }
        __syncthreads();
    }
}

template <bool kIndexed>
__global__ void __launch_bounds__(kThreads)
k_devroye_refill(double *__restrict__ x, const int *__restrict__ n, const double *__restrict__ z,
                 int64_t num, StreamId id, const int *__restrict__ idx, int chunk)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (kThreads / 32) * chunk;

    int64_t cur = warp * chunk;       // next unassigned observation of this warp (uniform)
    int64_t cend = cur + chunk;       // end of the current chunk (uniform)

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
            if (cand >= cend) cand += stride - chunk;
            if (!active && cand < num) {
                if (kIndexed) cand = idx[cand];                  // position in the class-sorted list -> observation
                int ni = n[cand];
                if (ni == 0) {
                    x[cand] = 0.0;                    // LogitWrapper.cpp:76-79
                } else {
                    obs = cand;
                    remaining = ni < 1 ? 1 : ni;      // NTHROW clamp, PolyaGamma.cpp:128-135
                    sum = 0.0;
                    st = dev_setup(z[cand]);
                    if (id.chain_len) {
                        const uint32_t ch = (uint32_t)cand / id.chain_len;
                        src.open(id.seed + ch, id.obs0 + ((uint32_t)cand - ch * id.chain_len), id.call_id);
                    } else {
                        src.open(id.seed, id.obs0 + (uint64_t)cand, id.call_id);
                    }
                    active = true;
                }
            }
            cur += __popc(want);
            if (cur >= cend) {
                int64_t over = cur - cend;
                cend += stride;
                cur = cend - chunk + over;
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
                                  StreamId id, cudaStream_t st, void *work)
{
    if (num <= 0) return cudaSuccess;
    int64_t cap = 148LL * 4;                 // 148 SMs x resident CTAs
    // Chunks are dealt round-robin to warps; a warp should see at least ~8 of them or the last
    // round leaves part of the chip idle (N = 1M rows of a Gibbs sweep is only 1.65 chunks of 128
    // per resident warp).
    int chunk = num >= cap * (kThreads / 32) * kChunkObs * 8 ? kChunkObs : kChunkObs / 4;
    int64_t chunks = (num + chunk - 1) / chunk;
    int64_t blocks = (chunks + (kThreads / 32) - 1) / (kThreads / 32);
    int grid = (int)(blocks < cap ? blocks : cap);
    if (work && num >= (1 << 20) && num < (1LL << 31)) {
        int *meta = (int *)work, *idx = meta + 32;
        cudaError_t e = cudaMemsetAsync(meta, 0, 32 * sizeof(int), st);
        if (e != cudaSuccess) return e;
        int tiles = (int)((num + kBinThreads - 1) / kBinThreads);
        int bgrid = tiles < 148 * 8 ? tiles : 148 * 8;
        k_cls_count<<<bgrid, kBinThreads, 0, st>>>(z, (int)num, meta);
        k_cls_scatter<<<bgrid, kBinThreads, 0, st>>>(z, (int)num, meta, idx);
        k_devroye_refill<true><<<grid, kThreads, 0, st>>>(x, n, z, num, id, idx, chunk);
        count_launch(3);
    } else {
        k_devroye_refill<false><<<grid, kThreads, 0, st>>>(x, n, z, num, id, nullptr, chunk);
        count_launch();
    }
    return cudaGetLastError();
}

}  // namespace bl
