// Device kernels of the Gibbs sweeps that consume PG draws.
//
// Reference statements (all fp64):
//   psi = X beta                    Logit.hpp:421,431; MultLogit.hpp:277,307     (gemm 'T')
//   X' Omega X (+ P0)               Logit.hpp:293-301, :325-332; MultLogit.hpp:246-248,252
//   X' v                            Logit.hpp:174-183 (kappa), MultLogit.hpp:249,253 (Omega c_j),
//                                   NBPG-logmean.R:28 (kappa + omega log d)
//   c_j = log sum_{k != j} exp(XB_k), eta_j = XB_j - c_j      MultLogit.hpp:288,293-299,313
//   beta draws                      Logit.hpp:291-320 (plain), :322-400 (constrained, the one
//                                   called), Normal.hpp:98-131 (mlogit)
//
// Data layout in HBM: tX is P x N column-major exactly as R hands it over
// (LogitWrapper.R:226), i.e. observation i's P covariates are contiguous at
// tX[i*P .. i*P+P): every pass below streams X once, row by row, fully coalesced.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "philox.cuh"
#include "specfun.cuh"

namespace bl {

// ---------------------------------------------------------------------------------
// psi_i = x_i . beta (+ off_i): one warp per row, lanes stride the row (coalesced),
// butterfly reduction.  HBM-bound: N*P*8 bytes.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_xbeta(double *__restrict__ psi, const double *__restrict__ tX, const double *__restrict__ beta,
        const double *__restrict__ off, double off_scale, double shift, int64_t N, int P)
{
    extern __shared__ double sbeta[];
    for (int p = threadIdx.x; p < P; p += blockDim.x) sbeta[p] = beta[p];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < N; i += warps) {
        const double *row = tX + i * P;
        double s = 0.0;
        for (int p = lane; p < P; p += 32) s = fma(row[p], sbeta[p], s);
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) psi[i] = (off ? s + off_scale * off[i] : s) + shift;
    }
}

// ---------------------------------------------------------------------------------
// Weighted Gram, SYRK-shaped: G += sum_i w_i x_i x_i' over a slab of rows.
// CTA = 256 threads as a 16x16 grid of 4x4 register tiles -> one 64x64 output tile
// (tile (bi,bj), bi <= bj, over blockIdx.y); row slabs over blockIdx.x.  Row chunks of
// kGramRows rows are staged in shared memory (x and w*x panels) so each X element is
// read from HBM once per output-tile row of the grid.  Per-CTA partial tiles go to
// `part` and are summed in a fixed order by k_gram_reduce: deterministic, no atomics.
// Bound by the FP64 pipe: 2*N*P^2 flops (P^2 + P when only the upper triangle counts).
// ---------------------------------------------------------------------------------
constexpr int kGramRows = 32;
constexpr int kGramTile = 64;

__global__ void __launch_bounds__(256)
k_gram_partial(double *__restrict__ part, const double *__restrict__ tX, const double *__restrict__ w,
               int64_t N, int P, int nt)
{
    __shared__ __align__(16) double sa[kGramRows][kGramTile + 2];   // w_i * x_i[bi panel]
    __shared__ __align__(16) double sb[kGramRows][kGramTile + 2];   // x_i[bj panel]
    // decode (bi, bj), bi <= bj, from blockIdx.y
    int t = blockIdx.y, bi = 0;
    while (t >= nt - bi) { t -= nt - bi; ++bi; }
    const int bj = bi + t;
    // thread -> 4x4 output block (ty, tx): rows 4ty.., cols 4tx...  On a diagonal tile only
    // the 136 blocks with ty <= tx are needed; they are packed into the first 136 threads
    // (4.25 warps) so the skipped half really frees FP64 issue slots.
    int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    bool live = true;
    if (bi == bj) {
        int r = threadIdx.x;
        live = r < 136;
        ty = 0;
        while (live && r >= 16 - ty) { r -= 16 - ty; ++ty; }
        tx = live ? ty + r : 0;
        if (!live) ty = 0;
    }
    double acc[4][4] = {};
    const int64_t slab = (N + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * slab;
    const int64_t r1 = r0 + slab < N ? r0 + slab : N;
    for (int64_t base = r0; base < r1; base += kGramRows) {
        __syncthreads();
        for (int e = threadIdx.x; e < kGramRows * kGramTile; e += 256) {
            int r = e >> 6, c = e & 63;
            int64_t i = base + r;
            double xa = 0.0, xb = 0.0;
            if (i < r1) {
                int ca = bi * kGramTile + c, cb = bj * kGramTile + c;
                double wi = w[i];
                if (ca < P) xa = tX[i * P + ca] * wi;
                if (cb < P) xb = tX[i * P + cb];
            }
            sa[r][c] = xa;
            sb[r][c] = xb;
        }
        __syncthreads();
        if (live) {
#pragma unroll 4
            for (int r = 0; r < kGramRows; ++r) {
                double a[4], b[4];
                const double2 a01 = *(const double2 *)&sa[r][4 * ty], a23 = *(const double2 *)&sa[r][4 * ty + 2];
                const double2 b01 = *(const double2 *)&sb[r][4 * tx], b23 = *(const double2 *)&sb[r][4 * tx + 2];
                a[0] = a01.x; a[1] = a01.y; a[2] = a23.x; a[3] = a23.y;
                b[0] = b01.x; b[1] = b01.y; b[2] = b23.x; b[3] = b23.y;
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], b[q], acc[p][q]);
            }
        }
    }
    double *out = part + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (kGramTile * kGramTile);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (live) out[(4 * ty + p) * kGramTile + 4 * tx + q] = acc[p][q];   // blocks below the diagonal are never read
}

// PP = P0 + sum over slabs of the partial tiles, mirrored to a full symmetric P x P
// column-major matrix.  One thread per upper-triangle element, fixed summation order.
__global__ void k_gram_reduce(double *__restrict__ PP, const double *__restrict__ P0,
                              const double *__restrict__ part, int P, int nt, int nslab)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P * P) return;
    int a = e % P, b = e / P;
    if (a > b) return;
    int bi = a / kGramTile, bj = b / kGramTile;
    int tile = 0;
    for (int k = 0; k < bi; ++k) tile += nt - k;
    tile += bj - bi;
    const double *src = part + (size_t)tile * nslab * (kGramTile * kGramTile)
                      + (a % kGramTile) * kGramTile + (b % kGramTile);
    double s = 0.0;
    for (int k = 0; k < nslab; ++k) s += src[(size_t)k * (kGramTile * kGramTile)];
    double v = s + (P0 ? P0[a + (size_t)P * b] : 0.0);
    PP[a + (size_t)P * b] = v;
    PP[b + (size_t)P * a] = v;
}

// out_p = sum_i x_i[p] * v_i  (X'v), v_i = c0*v0_i + c1*v1_i*v2_i  (v1/v2 optional).
// Per-CTA partial sums then a fixed-order reduce (k_xtv_reduce).
__global__ void __launch_bounds__(256)
k_xtv_partial(double *__restrict__ part, const double *__restrict__ tX, const double *__restrict__ v0,
              double c0, const double *__restrict__ v1, const double *__restrict__ v2, double c1,
              int64_t N, int P)
{
    extern __shared__ double sacc[];   // [warps][P]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int p = threadIdx.x; p < nw * P; p += blockDim.x) sacc[p] = 0.0;
    __syncthreads();
    const int64_t slab = (N + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * slab;
    const int64_t r1 = r0 + slab < N ? r0 + slab : N;
    for (int p = lane; p < P; p += 32) {
        double s = 0.0;
        for (int64_t i = r0 + warp; i < r1; i += nw) {
            double vi = (v0 ? c0 * v0[i] : 0.0) + (v1 ? c1 * v1[i] * (v2 ? v2[i] : 1.0) : 0.0);
            s = fma(tX[i * P + p], vi, s);
        }
        sacc[warp * P + p] = s;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < nw; ++k) s += sacc[k * P + p];
        part[(size_t)blockIdx.x * P + p] = s;
    }
}

__global__ void k_xtv_reduce(double *__restrict__ out, const double *__restrict__ add0,
                             const double *__restrict__ add1, const double *__restrict__ part,
                             int P, int nslab)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double s = 0.0;
    for (int k = 0; k < nslab; ++k) s += part[(size_t)k * P + p];
    out[p] = s + (add0 ? add0[p] : 0.0) + (add1 ? add1[p] : 0.0);
}

// kappa_i = n_i (y_i - 1/2)   (Logit.hpp:180-181);  NB: kappa_i = (y_i - d)/2
__global__ void k_kappa(double *__restrict__ kappa, const double *__restrict__ y,
                        const double *__restrict__ n, double d, int64_t N)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) kappa[i] = n ? n[i] * (y[i] - 0.5) : 0.5 * (y[i] - d);
}

// shape_i = (int) n_i  (Logit.hpp:287)  /  b_i = y_i + d  (NBPG-logmean.R:88)
__global__ void k_shape_int(int *__restrict__ out, const double *__restrict__ n, int64_t N)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = (int)n[i];
}
__global__ void k_shape_add(double *__restrict__ out, const double *__restrict__ y, double d, int64_t N)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) out[i] = y[i] + d;
}

// mlogit offsets for category j: A = sum_{k != j, k < J-1} exp(XB_k) + exp(0),
// c = log A, eta = XB_j - c.  XB is N x (J-1) column-major (the reference's J-th
// column is identically 0, MultLogit.hpp:275-277).  No max-subtraction, as there.
__global__ void k_mlogit_offsets(double *__restrict__ c, double *__restrict__ eta,
                                 const double *__restrict__ XB, int64_t N, int U, int j)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    // the reference sums columns of XB_no_j in order: categories 0..J-1 without j, the
    // all-zero last column included (exp(0) = 1 added last)
    double A = 0.0;
    for (int k = 0; k < U; ++k)
        if (k != j) A += exp(XB[i + (size_t)N * k]);
    A += 1.0;
    double cj = log(A);
    c[i] = cj;
    eta[i] = XB[i + (size_t)N * j] - cj;
}

}  // namespace bl
