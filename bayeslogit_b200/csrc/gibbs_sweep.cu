// One pass over X per logit Gibbs iteration: psi = X beta, omega = PG(n, psi) and the weighted Gram
// X' Omega X from the SAME shared-memory tile (kernel K3).
//
// Reference statements: psi = X beta (Logit.hpp:421,431), Logit::draw_w (:283-289), the Gram of
// Logit::draw_beta (:293-301, :325-332).  The three are adjacent in gibbs_block (:421-431) and X is
// the only large operand (N P 8 bytes; omega and psi are N 8 bytes), so the iteration's HBM traffic
// is one read of X when a row tile stays on the SM from its psi to its Gram.
//
// Persistent cooperative kernel, one CTA per SM, a static contiguous run of 32-row tiles per CTA,
// three warp roles around a ring of shared-memory stages:
//   producer (warp 0, one lane)   TMA: cp.async.bulk.tensor.2d of the tile's 16-column boxes
//                                 (128-byte swizzle) into the stage, completion on an mbarrier
//                                 (full[s]); rows past N arrive as zeros.
//   draw warps (D of them)        tile t belongs to draw warp t mod D.  It forms psi of the 32 rows
//                                 from the stage (the Gram's fragment pattern: lane (gid, tig) reads
//                                 X[r][8 j + gid], eight DFMAs against its eight beta entries, a
//                                 three-step butterfly over gid), draws omega = PG(n, psi) with the
//                                 filtered Devroye sampler (streams keyed by the global row), writes
//                                 omega to the stage (and to w_out), arrives on wready[s].
//   Gram warps (8)                tiles in order: wait wready[s], accumulate the 36 live 8 x 8 tiles
//                                 of the 64 x 64 Gram on the FP64 tensor path (DMMA m8n8k4) -- warp
//                                 (slot A, row group g) owns tile rows (A, 7 - A), nine tiles, and rows
//                                 [16 g, 16 g + 16) of every tile -- then arrive on empty[s].
// A draw takes ~6 us of latency on one warp and the Gram consumes a tile every ~0.6 us, so D = 12
// tiles are in flight: one 16 KB stage per draw warp (a stage cycles load -> draw -> Gram -> load).  The draws' instructions (ALU / FMA / MUFU pipes) fill the issue slots the DMMA-bound Gram
// leaves idle; the FP64 pipe carries the Gram's 288 DMMAs per tile plus 64 DFMAs for psi.
//
// Swizzled reads.  A box is [32 rows][16 doubles] with TMA's 128-byte swizzle: the 16-byte chunk c of
// row r sits at chunk c ^ (r & 7).  An MMA k-step takes rows {r0, r0 + 2, r0 + 4, r0 + 6} (tig = 0..3),
// so the eight lanes gid = 0..7 of one tig read chunks (4 (j & 1) + (gid >> 1)) ^ (2 tig + (r0 & 1)):
// across a half-warp the four rows land in four different chunk pairs -- every 128-byte wavefront
// conflict-free, no padding, and the row-to-slot assignment is free because k is summed over.
//
// Epilogue in the same launch: the two row groups fold through shared memory, every CTA stores its
// partial 64 x 64 tile, a grid barrier (cooperative launch), then CTA x < P P / 32 sums 32 entries
// over the CTAs' partials in a fixed order and mirrors them into PP; sharded sweeps publish PP to
// the peer windows from the last CTA (peer_publish).  One launch replaces four.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "engine.h"
#include "gibbs_kernels.cuh"
#include "gibbs_sweep.h"
#include "pg_devroye_fast.cuh"

namespace bl {

namespace {

constexpr int kTileRows = 32;
constexpr int kBoxCols = 16;                       // 16 doubles = 128 bytes = the swizzle span
constexpr int kBoxBytes = kTileRows * kBoxCols * 8;   // 4096
constexpr int kMaxThreads = 640;                   // 1 producer + 8 Gram + up to 11 draw warps: 96 registers per thread (ptxas sizes for multiples of 128 threads)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
{
    unsigned ok, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();      // a broken pipeline must fault, not hang the device
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// The Gram of one tile's 16 rows [16 grp, 16 grp + 16) for warp slot A (tile rows A and 7 - A of the 8 x 8
// grid of MMA tiles): c[p], p < 8 - A, is tile (A, A + p); the others are tiles (7 - A, 7 - A + (p - (8 - A))).
// ONE body for the four slots -- the slot enters through the nine precomputed column offsets boff[] (and two for
// the A operands, aoff[]) and through the count nlo = 8 - A: the four specialised bodies this replaces were
// 29 KB of SASS, and together with the sampler's code the draw warps run they overflowed the SM's instruction
// cache (hit rate 75 %, "no instruction" the top stall of the draw warps).  Cost: every B fragment is its own
// LDS (11 per k-step instead of 8 - A + 1).
template <int kRowsPerGroup>
__device__ __forceinline__ void sweep_gram_tile(double (&c)[9][2], uint32_t tile, uint32_t wts, int grp, int tig,
                                                const uint32_t (&boff)[9], const uint32_t (&aoff)[2], int nlo)
{
#pragma unroll
    for (int kk = 0; kk < kRowsPerGroup / 8; ++kk) {
#pragma unroll
        for (int p0 = 0; p0 < 2; ++p0) {
            const int r = kRowsPerGroup * grp + 8 * kk + 2 * tig + p0;   // this lane's k row; r & 7 = 2 tig + p0
            const double wr = lds_f64(wts + 8 * r);
            const uint32_t row = tile + 128 * r;
            // offsets are those of even rows (p0 = 0); an odd row flips bit 0 of the swizzled chunk index
            const uint32_t fl = p0 ? 16u : 0u;
            const double alo = lds_f64(row + (aoff[0] ^ fl)) * wr, ahi = lds_f64(row + (aoff[1] ^ fl)) * wr;
#pragma unroll
            for (int p = 0; p < 9; ++p) {
                const double b = lds_f64(row + (boff[p] ^ fl));
                // p < 5 is always a tile of row A, p == 8 always one of row 7 - A
                const double av = p < 5 ? alo : p == 8 ? ahi : (p < nlo ? alo : ahi);
                dmma884(c[p][0], c[p][1], av, b);
            }
        }
    }
}

struct SweepArgs {
    double *w_out;             // omega of this iteration [N] (always written: scratch when the caller keeps none)
    const int *shape;          // (int) n_i
    const double *beta;        // the beta psi is formed with [P]
    double *part;              // [grid][64 * 64] per-CTA partial tiles
    double *PP;                // [P * P] the reduced Gram (no prior), both triangles
    unsigned *grid_ctr;        // grid barrier counter (monotone over the launches of one chain)
    unsigned grid_target;      // counter value that releases this launch's barrier
    int64_t N;
    int P;
    int ntiles;                // ceil(N / 32)
    int draw_warps, stages;
    int serial_publish;        // A/B: publish from the last CTA alone
    int debug_nodraw;          // measurement aid: omega = 0.25 without drawing (Gram rate of the pipeline alone)
    StreamId id;
};

template <int kGramWarps>
__global__ void __launch_bounds__(kMaxThreads, 1)
k_logit_sweep(const __grid_constant__ CUtensorMap tmap, SweepArgs a, PeerPush px)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;             // 1024-byte aligned: the swizzle pattern is address based
    const int S = a.stages, D = a.draw_warps;
    const uint32_t tiles_sm = base;                                          // S x 16 KB
    const uint32_t wts_sm = tiles_sm + (uint32_t)S * (4 * kBoxBytes);         // S x 32 doubles
    const uint32_t bars = wts_sm + (uint32_t)S * (kTileRows * 8);             // full[S], wready[S], empty[S]
    const uint32_t beta_sm = bars + (uint32_t)S * 24;                         // beta, 64 doubles (zero past P)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    const int nboxes = (a.P + kBoxCols - 1) / kBoxCols;                        // column boxes that exist; the others stay zero

    // this CTA's run of tiles
    const int per = a.ntiles / gridDim.x, rem = a.ntiles % gridDim.x;
    const int t_begin = blockIdx.x * per + min((int)blockIdx.x, rem);
    const int t_count = per + ((int)blockIdx.x < rem ? 1 : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bars + 8 * s, 1);                   // full: the producer's arrive + the TMA bytes
            mbar_init(bars + 8 * (S + s), 1);             // wready: the draw warp's elected lane
            mbar_init(bars + 8 * (2 * S + s), kGramWarps);    // empty: one lane of every Gram warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // programmatic dependent launch: everything above is independent of the kernel before this one
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < 64)
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(beta_sm + 8 * threadIdx.x), "d"((int)threadIdx.x < a.P ? __ldcg(a.beta + threadIdx.x) : 0.0) : "memory");
    // boxes past P are never written by TMA: zero them once in every stage (generic-proxy stores, ordered
    // before any read by the barrier below; TMA never touches them)
    if (nboxes < 4) {
        for (int s = 0; s < S; ++s)
            for (int e = threadIdx.x; e < (4 - nboxes) * (kBoxBytes / 8); e += blockDim.x)
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(tiles_sm + s * 4 * kBoxBytes + nboxes * kBoxBytes + 8 * e), "d"(0.0) : "memory");
    }
    __syncthreads();

    double *fold = reinterpret_cast<double *>(smem_raw + (base - smem_u32(smem_raw)));
    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int k = 0; k < t_count; ++k) {
                const int s = k % S;
                if (k >= S) mbar_wait(bars + 8 * (2 * S + s), ((k / S) - 1) & 1);
                const uint32_t full = bars + 8 * s;
                mbar_expect_tx(full, (unsigned)(nboxes * kBoxBytes));
                const int row0 = (t_begin + k) * kTileRows;
                for (int b = 0; b < nboxes; ++b)
                    tma_load_2d(tiles_sm + s * 4 * kBoxBytes + b * kBoxBytes, &tmap, b * kBoxCols, row0, full);
            }
        }
    } else if (warp <= kGramWarps) {
        // ===== Gram warps =====
        const int gw = warp - 1, sub = gw & 3, grp = gw >> 2;
        const int nlo = 8 - sub;
        // byte offset of this lane's fragment element of column block j inside an (even) row of the stage
        auto coloff = [&](int j) {
            return (uint32_t)((j >> 1) * kBoxBytes + ((((j & 1) * 4 + (gid >> 1)) ^ (2 * tig)) << 4) + 8 * (gid & 1));
        };
        uint32_t boff[9], aoff[2] = {coloff(sub), coloff(7 - sub)};
#pragma unroll
        for (int p = 0; p < 9; ++p) boff[p] = coloff(p < nlo ? sub + p : (7 - sub) + (p - nlo));
        double c9[9][2] = {};
        for (int k = 0; k < t_count; ++k) {
            const int s = k % S;
            const unsigned par = (k / S) & 1;
            mbar_wait(bars + 8 * s, par);
            mbar_wait(bars + 8 * (S + s), par);
            sweep_gram_tile<kTileRows / (kGramWarps / 4)>(c9, tiles_sm + s * 4 * kBoxBytes, wts_sm + s * (kTileRows * 8), grp, tig, boff, aoff, nlo);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (2 * S + s));
        }
        // Every tile of this CTA has been multiplied, so every draw warp is past its last stage access and
        // no TMA write is outstanding: the ring's memory is free.  Fold the two row groups through it (named
        // barrier over the 8 Gram warps only -- the accumulators never live in the other roles' registers)
        // and store the CTA's partial tile.
        constexpr int kGroups = kGramWarps / 4;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kGramWarps) : "memory");
        if (grp > 0) {
            double *f = fold + (((grp - 1) * 4 + sub) * 32 + lane) * 19;
#pragma unroll
            for (int t = 0; t < 9; ++t) { f[2 * t] = c9[t][0]; f[2 * t + 1] = c9[t][1]; }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kGramWarps) : "memory");
        if (grp == 0) {
#pragma unroll
            for (int g = 1; g < kGroups; ++g) {
                const double *f = fold + (((g - 1) * 4 + sub) * 32 + lane) * 19;
#pragma unroll
                for (int t = 0; t < 9; ++t) { c9[t][0] += f[2 * t]; c9[t][1] += f[2 * t + 1]; }
            }
            double *out = a.part + (size_t)blockIdx.x * (kGramTile * kGramTile);
#pragma unroll
            for (int p = 0; p < 9; ++p) {
                const int ti = p < nlo ? sub : 7 - sub, tj = p < nlo ? sub + p : (7 - sub) + (p - nlo);
                const int row = 8 * ti + gid, col = 8 * tj + 2 * tig;
                out[row * kGramTile + col] = c9[p][0];
                out[row * kGramTile + col + 1] = c9[p][1];
            }
        }
    } else if (warp < 1 + kGramWarps + D) {
        // ===== draw warps =====
        const int dw = warp - 1 - kGramWarps;
        // the row of the tile this lane draws: the one whose psi the butterfly leaves in it
        const int myrow = 8 * (gid >> 1) + 2 * tig + (gid & 1);
        for (int k = dw; k < t_count; k += D) {
            const int s = k % S;
            // S == D: this warp sees every use of its stage, so the parity of use k / S is unambiguous (a warp
            // that skipped uses could not tell "use u has landed" from "use u - 1 has not")
            mbar_wait(bars + 8 * s, (k / S) & 1);
            const uint32_t tile = tiles_sm + s * 4 * kBoxBytes;
            double z = 0.0;
            double br[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) br[j] = lds_f64(beta_sm + 8 * (8 * j + gid));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int p0 = 0; p0 < 2; ++p0) {
                    const int r = 8 * q + 2 * tig + p0;
                    const uint32_t row = tile + 128 * r + 8 * (gid & 1);
                    const int x = 2 * tig + p0;
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        acc = fma(lds_f64(row + (j >> 1) * kBoxBytes + ((((j & 1) * 4 + (gid >> 1)) ^ x) << 4)), br[j], acc);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    if (gid == 2 * q + p0) z = acc;
                }
            }
            const int64_t i = (int64_t)(t_begin + k) * kTileRows + myrow;
            double om = 0.0;
            if (i < a.N) {
                const int ni = a.shape[i];
                if (a.debug_nodraw) om = 0.25;                           // 1: + no multiply, see sweep_gram_tile; 2: plain
                else if (ni != 0) {                                      // LogitWrapper.cpp:76-79: n == 0 -> 0
                    int remaining = ni < 1 ? 1 : ni;                     // NTHROW clamp, PolyaGamma.cpp:128-135
                    DevSetup st = dev_setup(z);
                    PhiloxSource src;
                    src.open(a.id.seed, a.id.obs0 + (uint64_t)i, a.id.call_id);
                    do {
                        double X;
                        while (!dev_propose(src, st, X)) {}
                        om += 0.25 * X;
                    } while (--remaining);
                }
                a.w_out[i] = om;
            }
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(wts_sm + s * (kTileRows * 8) + 8 * myrow), "d"(om) : "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (S + s));
        }
    }
    // ---- grid barrier (cooperative launch: every CTA is resident) ----
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.grid_ctr, 1u);
        unsigned spins = 0;
        while ((int)(ld_acquire_gpu(a.grid_ctr) - a.grid_target) < 0) {
            __nanosleep(32);
            if (++spins > (1u << 26)) __trap();
        }
    }
    __syncthreads();

    // ---- PP[a, b] = sum over the CTAs' partial tiles, fixed order; a CTA owns runs of 32 entries of the tile ----
    {
        double *red = fold;                                           // [warps][32]
        const int nw = blockDim.x >> 5;
        for (int e0 = blockIdx.x * 32; e0 < a.P * kGramTile; e0 += gridDim.x * 32) {
            const int e = e0 + lane;
            const int ra = e >> 6, cb = e & 63;                       // tile entry (row, col); stored where (row >> 3) <= (col >> 3)
            const bool want = ra <= cb && cb < a.P;
            double s = 0.0;
            if (want) {
                const double *src = a.part + e;
#pragma unroll 4
                for (int k = warp; k < (int)gridDim.x; k += nw) s += __ldcg(src + (size_t)k * (kGramTile * kGramTile));
            }
            __syncthreads();
            red[warp * 32 + lane] = s;
            __syncthreads();
            if (warp == 0 && want) {
                double v = 0.0;
                for (int k = 0; k < nw; ++k) v += red[k * 32 + lane];
                a.PP[ra + (size_t)a.P * cb] = v;
                a.PP[cb + (size_t)a.P * ra] = v;
            }
        }
    }
    // Sharded data: the finished sums go into every rank's window -- one CTA per destination.  (From the last CTA
    // alone, one window after the other, the kernel was 4 us longer at 2 ranks and 20 us longer at 8 than on a GPU
    // that runs the same shard alone.)  The grid counter makes a second round: every CTA adds one when its part of
    // PP is stored, CTA r waits for the round to complete, copies PP into rank r's window with coalesced 16-byte
    // stores, and sets its flag there behind a system-scope fence.  All CTAs are resident (cooperative launch), so
    // the wait cannot starve the CTAs it waits for.
    if (px.world > 1 && a.serial_publish) {
        peer_publish(px, a.PP, a.P * a.P);                   // A/B: the last CTA publishes to every rank (BL_K3_SERIAL_PUBLISH)
    } else if (px.world > 1) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(a.grid_ctr, 1u);
        }
        if ((int)blockIdx.x < px.world) {
            if (threadIdx.x == 0) {
                unsigned spins = 0;
                while ((int)(ld_acquire_gpu(a.grid_ctr) - (a.grid_target + gridDim.x)) < 0) {
                    __nanosleep(32);
                    if (++spins > (1u << 26)) __trap();
                }
            }
            __syncthreads();
            const int cnt = a.P * a.P, n2 = cnt >> 1;
            const double2 *src = reinterpret_cast<const double2 *>(a.PP);
            for (int r = blockIdx.x; r < px.world; r += gridDim.x) {
                double2 *dst = reinterpret_cast<double2 *>(px.slot[r]);
                for (int i = threadIdx.x; i < n2; i += blockDim.x) dst[i] = __ldcg(src + i);
                if ((cnt & 1) && threadIdx.x == 0) px.slot[r][cnt - 1] = __ldcg(a.PP + cnt - 1);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence_system();             // the slot stores are performed system-wide before the flags
                for (int r = blockIdx.x; r < px.world; r += gridDim.x) st_release_sys(px.flag[r], px.epoch);
            }
        }
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encode_fn()
{
    static EncodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiled)p;
    }();
    return fn;
}

}  // namespace

bool logit_sweep_ok(const double *tX, int P)
{
    if (getenv("BL_GIBBS_NO_K3")) return false;
    return P % 2 == 0 && P >= 2 && P <= 64 && (reinterpret_cast<uintptr_t>(tX) & 15) == 0;
}

int LogitSweep::init(const double *tX, int64_t N_, int P_, cudaStream_t st, std::string &err)
{
    N = N_; P = P_;
    EncodeTiled enc = encode_fn();
    if (!enc) { err = "cuTensorMapEncodeTiled is not available from this driver"; return 1; }
    static_assert(sizeof(map_storage) >= sizeof(CUtensorMap), "tensor map storage");
    CUtensorMap *map = reinterpret_cast<CUtensorMap *>(map_storage);
    const cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)P * 8};
    const cuuint32_t box[2] = {(cuuint32_t)kBoxCols, (cuuint32_t)kTileRows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(tX), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return 1; }
    ntiles = (int)((N + kTileRows - 1) / kTileRows);
    const char *e = getenv("BL_K3_GRAM_WARPS");
    gram_warps = e && atoi(e) == 16 ? 16 : 8;
    e = getenv("BL_K3_DRAW_WARPS");
    draw_warps = e ? atoi(e) : 11;
    draw_warps = std::max(1, std::min(draw_warps, kMaxThreads / 32 - 1 - gram_warps));
    stages = draw_warps;            // one stage per draw warp: tile k -> stage k mod S -> draw warp k mod D, the same warp every use
    smem = 1024 + (size_t)stages * (4 * kBoxBytes + kTileRows * 8 + 3 * 8) + 64 * 8;
    // the epilogue folds the row groups through the ring's memory: (groups - 1) x 4 slots x 32 lanes x 19 doubles
    smem = std::max(smem, (size_t)1024 + (size_t)(gram_warps / 4 - 1) * 4 * 32 * 19 * 8 + 1024);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // at least four tiles per CTA, at most one CTA per SM (cooperative launch: all resident)
    grid = std::max(1, std::min(sms, ntiles / 4));
    cudaError_t ce = cudaFuncSetAttribute(k_logit_sweep<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_logit_sweep<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce != cudaSuccess) { err = std::string("k_logit_sweep shared memory: ") + cudaGetErrorString(ce); return 1; }
    if ((ce = cudaMallocAsync((void **)&part, (size_t)grid * kGramTile * kGramTile * sizeof(double), st)) != cudaSuccess ||
        (ce = cudaMallocAsync((void **)&ctr, sizeof(unsigned), st)) != cudaSuccess ||
        (ce = cudaMemsetAsync(ctr, 0, sizeof(unsigned), st)) != cudaSuccess) {
        err = std::string("k_logit_sweep scratch: ") + cudaGetErrorString(ce);
        return 1;
    }
    stream = st;
    launches = 0;
    ctr_rounds = 0;
    return 0;
}

LogitSweep::~LogitSweep()
{
    if (part) cudaFreeAsync(part, stream);
    if (ctr) cudaFreeAsync(ctr, stream);
}

cudaError_t LogitSweep::launch(double *w_out, double *PP, const int *shape, const double *beta, StreamId id,
                               const PeerPush &px)
{
    SweepArgs a;
    a.w_out = w_out; a.shape = shape; a.beta = beta; a.part = part; a.PP = PP;
    a.grid_ctr = ctr;
    a.serial_publish = getenv("BL_K3_SERIAL_PUBLISH") ? 1 : 0;
    // two rounds of the grid counter per sharded launch (barrier, publish), one otherwise
    a.grid_target = (unsigned)(ctr_rounds + (uint64_t)grid);
    ctr_rounds += (uint64_t)grid * (px.world > 1 && !a.serial_publish ? 2 : 1);
    a.N = N; a.P = P; a.ntiles = ntiles; a.draw_warps = draw_warps; a.stages = stages; a.id = id;
    a.debug_nodraw = getenv("BL_K3_NODRAW") ? atoi(getenv("BL_K3_NODRAW")) : 0;
    ++launches;
    CUtensorMap *map = reinterpret_cast<CUtensorMap *>(map_storage);
    PeerPush pxc = px;
    void *args[] = {map, &a, &pxc};
    const int threads = 32 * (1 + gram_warps + draw_warps);
    const void *fn = gram_warps == 16 ? (const void *)k_logit_sweep<16> : (const void *)k_logit_sweep<8>;
    // Cooperative AND programmatic dependent launch: the CTAs are scheduled while the beta draw before them still
    // runs -- barrier set-up and the first TMA loads of X do not depend on beta; every thread passes
    // griddepcontrol.wait before beta is read or anything is written.  Drivers that refuse the combination get
    // the plain cooperative launch.
    static int pdl = getenv("BL_GIBBS_NO_PDL") ? 0 : 1;
    cudaError_t e = cudaErrorNotSupported;
    if (pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        e = cudaLaunchKernelExC(&cfg, fn, args);
        if (e != cudaSuccess) { (void)cudaGetLastError(); pdl = 0; }
    }
    if (e != cudaSuccess) e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, smem, stream);
    count_launch();
    return e;
}

}  // namespace bl
