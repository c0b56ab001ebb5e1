// rpg_hybrid on the device: regime binning + one kernel per regime.
//
// The reference dispatches per observation inside one serial loop
// (LogitWrapper.cpp:140-162; same logic in PolyaGammaHybrid.h:26-55).  Running that
// switch per lane makes every warp execute every sampler its 32 observations need
// and makes the kernel's code footprint the sum of all samplers (~400 KB of SASS,
// far beyond the instruction cache -- the profile in profiles/r1_00_* shows 65 of
// 70 stall cycles per instruction waiting on instruction fetch).  Instead:
//
//   1. k_hyb_count    histogram of regimes (block-aggregated atomics)
//   2. k_hyb_offsets  exclusive prefix -> list offsets
//   3. k_hyb_scatter  counting sort of observation indices by regime into one
//                     int32 list (block-level ballot ranking, ascending within a
//                     block so gathers stay nearly coalesced); b <= 0 -> x = 0
//   4. one kernel per regime over its index list: every warp runs ONE sampler.
//
// Results are unchanged by the re-ordering: each observation draws from the Philox
// stream keyed by its own global index (philox.cuh).
#include "engine.h"
#include "pg_devroye_fast.cuh"

namespace bl {

namespace {

constexpr int kBinThreads = 256;
constexpr int kMetaCounts = 0;    // meta[0..7]   regime counts
constexpr int kMetaOffsets = 8;   // meta[8..15]  list offsets
constexpr int kMetaCursor = 16;   // meta[16..23] scatter cursors

__global__ void __launch_bounds__(kBinThreads)
k_hyb_count(const double *__restrict__ h, int n, int *__restrict__ meta)
{
    __shared__ int cnt[8];
    if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
    __syncthreads();
    int local[6] = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        local[regime_of(h[i])]++;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        int v = local[r];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[r], v);
    }
    __syncthreads();
    if (threadIdx.x < 6 && cnt[threadIdx.x]) atomicAdd(&meta[kMetaCounts + threadIdx.x], cnt[threadIdx.x]);
}

__global__ void k_hyb_offsets(int *meta)
{
    int acc = 0;
    for (int r = 1; r < 6; ++r) {   // regime 0 (b <= 0) needs no list
        meta[kMetaOffsets + r] = acc;
        meta[kMetaCursor + r] = 0;
        acc += meta[kMetaCounts + r];
    }
}

__global__ void __launch_bounds__(kBinThreads)
k_hyb_scatter(const double *__restrict__ h, int n, int *__restrict__ meta, int *__restrict__ idx,
              double *__restrict__ x)
{
    __shared__ int wcnt[kBinThreads / 32][8];
    __shared__ int base[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int tiles = (n + kBinThreads - 1) / kBinThreads;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        int i = tile * kBinThreads + threadIdx.x;
        int reg = i < n ? regime_of(h[i]) : -1;
        if (reg == kRegZero) x[i] = 0.0;                      // LogitWrapper.cpp:159-161
        int rank = 0;
#pragma unroll
        for (int r = 1; r < 6; ++r) {
            unsigned m = __ballot_sync(0xffffffffu, reg == r);
            if (reg == r) rank = __popc(m & lt);
            if (lane == 0) wcnt[warp][r] = __popc(m);
        }
        __syncthreads();
        if (threadIdx.x >= 1 && threadIdx.x < 6) {
            int r = threadIdx.x, tot = 0;
            for (int w = 0; w < kBinThreads / 32; ++w) tot += wcnt[w][r];
            base[r] = tot ? meta[kMetaOffsets + r] + atomicAdd(&meta[kMetaCursor + r], tot) : 0;
        }
        __syncthreads();
        if (reg > 0) {
            int off = base[reg];
            for (int w = 0; w < warp; ++w) off += wcnt[w][reg];
            idx[off + rank] = i;
        }
        __syncthreads();
    }
}

template <int R>
__global__ void __launch_bounds__(128)
k_hyb_regime(double *__restrict__ x, const double *__restrict__ h, const double *__restrict__ z,
             const int *__restrict__ idx, const int *__restrict__ meta, StreamId id)
{
    const int count = meta[kMetaCounts + R];
    const int *list = idx + meta[kMetaOffsets + R];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
        int i = list[j];
        PhiloxSource s;
        s.open(id.seed, id.obs0 + (uint64_t)i, id.call_id);
        double b = h[i], zi = z[i], v;
        if (R == kRegNormal) {
            double m = pg_m1(b, zi);
            double var = pg_m2(b, zi) - m * m;
            v = m + sqrt(var) * s.norm();
        } else if (R == kRegSP) {
            sp_draw(s, v, b, zi);
        } else if (R == kRegDevroye) {
            v = devroye_sum_fast(s, (int)b, zi);
        } else if (R == kRegAlt) {
            v = alt_draw(s, b, zi);
        } else {
            v = gamma_sum(s, b, zi, 200);
        }
        x[i] = v;
    }
}

// Saddle-point regime as two kernels (pg_sp.cuh): set-up -> 15-double state per draw in HBM
// (struct of arrays over the chunk, coalesced) -> rejection loop.  Each kernel's working set of
// code stays near the 32 KB instruction cache; the state costs 240 B of HBM traffic per draw,
// ~2 % of HBM bandwidth at the rates these kernels reach.
__global__ void __launch_bounds__(128)
k_sp_setup(const double *__restrict__ h, const double *__restrict__ z, const int *__restrict__ idx,
           const int *__restrict__ meta, double *__restrict__ state, int c0, int cap)
{
    const int count = min(meta[kMetaCounts + kRegSP] - c0, cap);
    const int *list = idx + meta[kMetaOffsets + kRegSP] + c0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < count; j += gridDim.x * blockDim.x) {
        int i = list[j];
        SpState s;
        sp_setup(h[i], z[i], s);
#pragma unroll
        for (int k = 0; k < kSpStateDoubles; ++k) state[(size_t)k * cap + j] = s.f[k];
    }
}

// Rejection loop on persistent lanes: a lane makes one trip (sp_trip) per pass and, when its draw
// is complete, takes the next list position of its warp's chunk through a ballot-compacted
// refill, so rejected proposals and slow inner loops of one draw do not idle the other 31 lanes.
constexpr int kSpLoopThreads = 128;
constexpr int kSpLaneChunk = 128;   // consecutive list positions a warp works through

__global__ void __launch_bounds__(kSpLoopThreads)
k_sp_loop(double *__restrict__ x, const double *__restrict__ h, const double *__restrict__ z,
          const int *__restrict__ idx, const int *__restrict__ meta, const double *__restrict__ state,
          int c0, int cap, StreamId id)
{
    const unsigned full = 0xffffffffu;
    const int count = min(meta[kMetaCounts + kRegSP] - c0, cap);
    if (count <= 0) return;
    const int *list = idx + meta[kMetaOffsets + kRegSP] + c0;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int warp = blockIdx.x * (kSpLoopThreads / 32) + (threadIdx.x >> 5);
    const int stride = gridDim.x * (kSpLoopThreads / 32) * kSpLaneChunk;

    int cur = warp * kSpLaneChunk;     // next unassigned list position of this warp (uniform)
    int cend = cur + kSpLaneChunk;     // end of the current chunk (uniform)

    bool active = false;
    int obs = 0;
    double n = 0.0, zh = 0.0;
    SpStateRef st{state, (size_t)cap};
    SpLane L;
    L.start();
    PhiloxSource src;

    for (;;) {
        unsigned want = __ballot_sync(full, !active);
        if (want && cur < count) {
            int rank = __popc(want & lt_mask);
            int cand = cur + rank;
            if (cand >= cend) cand += stride - kSpLaneChunk;
            if (!active && cand < count) {
                obs = list[cand];
                n = h[obs];
                zh = 0.5 * fabs(z[obs]);
                st.o = state + cand;
                L.start();
                src.open(id.seed, id.obs0 + (uint64_t)obs, id.call_id);
                active = true;
            }
            cur += __popc(want);
            if (cur >= cend) {
                int over = cur - cend;
                cend += stride;
                cur = cend - kSpLaneChunk + over;
            }
        }
        if (!__any_sync(full, active)) {
            if (cur >= count) break;
            continue;
        }
        if (active && sp_trip(src, L, n, zh, st)) {
            x[obs] = n * 0.25 * L.X;
            active = false;
        }
    }
}

constexpr int kSpChunk = 1 << 24;   // draws per set-up/loop kernel pair (1.6 GB of state)

template <int R>
void launch_regime(double *x, const double *h, const double *z, const int *idx, const int *meta,
                   StreamId id, int n, int ctas_per_sm, cudaStream_t st)
{
    int need = (n + 127) / 128;
    int cap = 148 * ctas_per_sm;
    k_hyb_regime<R><<<need < cap ? need : cap, 128, 0, st>>>(x, h, z, idx, meta, id);
    count_launch();
}

}  // namespace

// optional per-stage timing of the last binned launch (bench.py's roofline leg)
static bool g_hyb_timing = false;
static cudaEvent_t g_hyb_ev[7];
static bool g_hyb_ev_ready = false;

void hybrid_timing_enable(bool on)
{
    g_hyb_timing = on;
    if (on && !g_hyb_ev_ready) {
        for (auto &e : g_hyb_ev) cudaEventCreate(&e);
        g_hyb_ev_ready = true;
    }
}

// ms of [binning, SP, Alt, sum-of-gammas, normal, Devroye] of the last timed launch
int hybrid_timing_last(double *out6)
{
    if (!g_hyb_ev_ready) return 1;
    if (cudaEventSynchronize(g_hyb_ev[6]) != cudaSuccess) return 1;
    for (int k = 0; k < 6; ++k) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_hyb_ev[k], g_hyb_ev[k + 1]);
        out6[k] = ms;
    }
    return 0;
}

// [meta: 32 ints][index list: num ints, padded to 16 B][saddle-point state: 12 x min(num, chunk) doubles]
static size_t hybrid_state_offset(int64_t num) { return ((32 + (size_t)num) * sizeof(int) + 15) / 16 * 16; }

size_t hybrid_workspace_bytes(int64_t num)
{
    size_t cap = (size_t)(num < kSpChunk ? num : kSpChunk);
    return hybrid_state_offset(num) + cap * kSpStateDoubles * sizeof(double);
}

// One rpg_hybrid batch of at most 2^31-1 observations.  `work` holds
// hybrid_workspace_bytes(num) bytes of device scratch owned by the caller's stream.
cudaError_t launch_hybrid_binned(double *x, const double *h, const double *z, int num, StreamId id,
                                 void *work, cudaStream_t st)
{
    if (num <= 0) return cudaSuccess;
    int *meta = (int *)work;
    int *idx = meta + 32;
    cudaError_t e = cudaMemsetAsync(meta, 0, 32 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    int tiles = (num + kBinThreads - 1) / kBinThreads;
    int grid = tiles < 148 * 8 ? tiles : 148 * 8;
    const bool tm = g_hyb_timing;
    if (tm) cudaEventRecord(g_hyb_ev[0], st);
    k_hyb_count<<<grid, kBinThreads, 0, st>>>(h, num, meta);
    k_hyb_offsets<<<1, 1, 0, st>>>(meta);
    k_hyb_scatter<<<grid, kBinThreads, 0, st>>>(h, num, meta, idx, x);
    count_launch(3);
    // heavy regimes first so the light ones fill the tail
    if (tm) cudaEventRecord(g_hyb_ev[1], st);
    {
        double *state = (double *)((char *)work + hybrid_state_offset(num));
        int cap = num < kSpChunk ? num : kSpChunk;
        int need = (cap + 127) / 128;
        int grid = need < 148 * 8 ? need : 148 * 8;
        for (int c0 = 0; c0 < num; c0 += cap) {   // pairs past the regime's count return at once
            k_sp_setup<<<grid, 128, 0, st>>>(h, z, idx, meta, state, c0, cap);
            k_sp_loop<<<grid, 128, 0, st>>>(x, h, z, idx, meta, state, c0, cap, id);
            count_launch(2);
        }
    }
    if (tm) cudaEventRecord(g_hyb_ev[2], st);
    launch_regime<kRegAlt>(x, h, z, idx, meta, id, num, 4, st);
    if (tm) cudaEventRecord(g_hyb_ev[3], st);
    launch_regime<kRegGamma>(x, h, z, idx, meta, id, num, 8, st);
    if (tm) cudaEventRecord(g_hyb_ev[4], st);
    launch_regime<kRegNormal>(x, h, z, idx, meta, id, num, 8, st);
    if (tm) cudaEventRecord(g_hyb_ev[5], st);
    launch_regime<kRegDevroye>(x, h, z, idx, meta, id, num, 8, st);
    if (tm) cudaEventRecord(g_hyb_ev[6], st);
    return cudaGetLastError();
}

}  // namespace bl
