// Single-CTA dense algebra and the three beta draws of the Gibbs sweeps.
//
// P is small (<= 256) and the draw sits on the critical path between two sweeps, so
// one CTA does the whole thing out of shared memory (global scratch when 2 P^2
// doubles do not fit): Cholesky, triangular solves, the draw.  It is replicated on
// every GPU with an identical Philox stream, so beta stays bit-identical across
// ranks without a broadcast.  Reference statements:
//   plain        beta ~ N(PP^-1 bP, PP^-1)                  Logit.hpp:291-320
//   constrained  coordinate-wise truncated normals, beta_j >= 0 for j < P-1
//                (the draw Logit::gibbs_block actually calls)     Logit.hpp:322-400
//   mvn          Normal::set_from_likelihood + draw               Normal.hpp:98-131
// Variates come from the stream (seed, obs 2^64-1, call t) in statement order:
// plain/mvn: P normals; constrained: per sweep P-1 uniforms (r.flat) then P
// truncated normals (r.tnorm: inverse CDF on one uniform, Robert's rejection samplers
// in the far tails -- the reference's own generator lives in its absent RNG library).
#pragma once

#include "philox.cuh"
#include "specfun.cuh"

namespace bl {

enum BetaDraw { kBetaConstrained = 0, kBetaPlain = 1, kBetaMvn = 2 };

// A = U'U in place (upper triangle of column-major A, ld = P); strict lower part zeroed.
// Returns false through *ok when a pivot is not positive.
__device__ inline void cta_chol_upper(double *A, int P, int ld, int *ok)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = 0; j < P; ++j) {
        __syncthreads();
        double d = A[j + (size_t)ld * j];
        if (!(d > 0.0)) { if (tid == 0) *ok = 0; __syncthreads(); return; }
        d = sqrt(d);
        __syncthreads();
        for (int k = j + tid; k < P; k += nt) A[j + (size_t)ld * k] = k == j ? d : A[j + (size_t)ld * k] / d;
        __syncthreads();
        // trailing update on a 16-wide thread grid (no integer division in the index math)
        for (int k = j + 1 + (tid >> 4); k < P; k += nt >> 4) {
            double ujk = A[j + (size_t)ld * k];
            for (int i = j + 1 + (tid & 15); i <= k; i += 16)
                A[i + (size_t)ld * k] = fma(-A[j + (size_t)ld * i], ujk, A[i + (size_t)ld * k]);
        }
    }
    __syncthreads();
    for (int e = tid; e < P * P; e += nt) {
        int i = e % P, k = e / P;
        if (i > k) A[i + (size_t)ld * k] = 0.0;
    }
    __syncthreads();
}

// A = LL' in place (lower triangle); strict upper part zeroed.
__device__ inline void cta_chol_lower(double *A, int P, int ld, int *ok)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = 0; j < P; ++j) {
        __syncthreads();
        double d = A[j + (size_t)ld * j];
        if (!(d > 0.0)) { if (tid == 0) *ok = 0; __syncthreads(); return; }
        d = sqrt(d);
        __syncthreads();
        for (int i = j + tid; i < P; i += nt) A[i + (size_t)ld * j] = i == j ? d : A[i + (size_t)ld * j] / d;
        __syncthreads();
        for (int k = j + 1 + (tid >> 4); k < P; k += nt >> 4) {
            double lkj = A[k + (size_t)ld * j];
            for (int i = k + (tid & 15); i < P; i += 16)
                A[i + (size_t)ld * k] = fma(-A[i + (size_t)ld * j], lkj, A[i + (size_t)ld * k]);
        }
    }
    __syncthreads();
    for (int e = tid; e < P * P; e += nt) {
        int i = e % P, k = e / P;
        if (i < k) A[i + (size_t)ld * k] = 0.0;
    }
    __syncthreads();
}

// x <- (U'U)^-1 x for ncol right-hand sides (columns of X, ld = P): one thread per
// column, no synchronisation inside.
__device__ inline void cta_solve_utu(const double *U, double *X, int P, int ld, int ncol)
{
    for (int c = threadIdx.x; c < ncol; c += blockDim.x) {
        double *x = X + (size_t)ld * c;
        for (int i = 0; i < P; ++i) {
            double s = x[i];
            for (int k = 0; k < i; ++k) s = fma(-U[k + (size_t)ld * i], x[k], s);
            x[i] = s / U[i + (size_t)ld * i];
        }
        for (int i = P - 1; i >= 0; --i) {
            double s = x[i];
            for (int k = i + 1; k < P; ++k) s = fma(-U[i + (size_t)ld * k], x[k], s);
            x[i] = s / U[i + (size_t)ld * i];
        }
    }
    __syncthreads();
}

// Warp-cooperative single right-hand side solves (lane-parallel dot products).
__device__ inline double warp_sum(double v)
{
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x <- U^-T x (forward), one warp.
__device__ inline void warp_solve_ut(const double *U, double *x, int P, int ld, int lane)
{
    for (int i = 0; i < P; ++i) {
        double s = 0.0;
        for (int k = lane; k < i; k += 32) s = fma(U[k + (size_t)ld * i], x[k], s);
        s = warp_sum(s);
        if (lane == 0) x[i] = (x[i] - s) / U[i + (size_t)ld * i];
        __syncwarp();
    }
}

// x <- U^-1 x (backward), one warp.
__device__ inline void warp_solve_u(const double *U, double *x, int P, int ld, int lane)
{
    for (int i = P - 1; i >= 0; --i) {
        double s = 0.0;
        for (int k = i + 1 + lane; k < P; k += 32) s = fma(U[i + (size_t)ld * k], x[k], s);
        s = warp_sum(s);
        if (lane == 0) x[i] = (x[i] - s) / U[i + (size_t)ld * i];
        __syncwarp();
    }
}

// x <- L^-1 x (forward, lower), one warp.
__device__ inline void warp_solve_l(const double *L, double *x, int P, int ld, int lane)
{
    for (int i = 0; i < P; ++i) {
        double s = 0.0;
        for (int k = lane; k < i; k += 32) s = fma(L[i + (size_t)ld * k], x[k], s);
        s = warp_sum(s);
        if (lane == 0) x[i] = (x[i] - s) / L[i + (size_t)ld * i];
        __syncwarp();
    }
}

// Truncated N(0,1) on (a, b); same construction as the oracle's pgo_tnorm (the reference's
// RNG::tnorm lives in its absent library).  Far tails (a >= 4, or b <= -4 mirrored) use the
// rejection samplers of Robert (1995): uniform proposal on a narrow interval ((U U)+),
// translated exponential otherwise ((E [U])+); elsewhere inverse CDF on one uniform,
// evaluated on the tail that keeps precision.
__device__ inline double tnorm_tail(PhiloxSource &s, double a, double b)
{
    if (b < INFINITY && (b - a) * a < 1.0) {
        for (;;) {
            double z = a + (b - a) * s.unif();
            if (s.unif() < exp(0.5 * (a * a - z * z))) return z;
        }
    }
    double astar = 0.5 * (a + sqrt(a * a + 4.0));
    for (;;) {
        double z = a + s.expon() / astar;
        if (z > b) continue;
        if (s.unif() < exp(-0.5 * (z - astar) * (z - astar))) return z;
    }
}

__device__ inline double tnorm_std(PhiloxSource &s, double a, double b)
{
    if (!(a < b)) return a;
    if (b <= -4.0) return -tnorm_tail(s, -b, -a);
    if (a >= 4.0) return tnorm_tail(s, a, b);
    double u = s.unif();
    double z;
    if (a >= 0.0 || (a > -INFINITY && -a < b)) {
        double qa = isinf(a) ? 1.0 : 0.5 * erfc(a * kSqrt1_2);
        double qb = isinf(b) ? 0.0 : 0.5 * erfc(b * kSqrt1_2);
        double q = qa - u * (qa - qb);
        z = q <= 0.0 ? INFINITY : q >= 1.0 ? -INFINITY : -normcdfinv(q);
    } else {
        double pa = isinf(a) ? 0.0 : 0.5 * erfc(-a * kSqrt1_2);
        double pb = isinf(b) ? 1.0 : 0.5 * erfc(-b * kSqrt1_2);
        double p = pa + u * (pb - pa);
        z = p <= 0.0 ? -INFINITY : p >= 1.0 ? INFINITY : normcdfinv(p);
    }
    if (z < a) z = a;
    if (z > b) z = b;
    return z;
}

// tnorm_std for a whole warp holding identical (a, b) and identical stream state: the two tail
// probabilities are evaluated side by side on lanes 0 and 1 and exchanged by shuffle, which halves
// the dependent-instruction chain of the inverse-CDF branch (the constrained beta draw runs P^2 of
// these back to back on one warp).  Same values as tnorm_std.
__device__ inline double tnorm_std_warp(PhiloxSource &s, double a, double b, int lane)
{
    if (!(a < b)) return a;
    if (b <= -4.0) return -tnorm_tail(s, -b, -a);
    if (a >= 4.0) return tnorm_tail(s, a, b);
    double u = s.unif();
    const bool upper = a >= 0.0 || (a > -INFINITY && -a < b);
    const double sg = upper ? 1.0 : -1.0;
    // lane 0: tail probability at a, lane 1: at b (upper tails Q if `upper`, lower tails P otherwise)
    double arg = lane == 0 ? a : b;
    double t = 0.0;
    if (lane < 2) t = isinf(arg) ? ((arg > 0) == upper ? 0.0 : 1.0) : 0.5 * erfc(sg * arg * kSqrt1_2);
    const double ta = __shfl_sync(0xffffffffu, t, 0), tb = __shfl_sync(0xffffffffu, t, 1);
    double z;
    if (upper) {
        double q = ta - u * (ta - tb);
        z = q <= 0.0 ? INFINITY : q >= 1.0 ? -INFINITY : -normcdfinv(q);
    } else {
        double p = ta + u * (tb - ta);
        z = p <= 0.0 ? -INFINITY : p >= 1.0 ? INFINITY : normcdfinv(p);
    }
    if (z < a) z = a;
    if (z > b) z = b;
    return z;
}

// k-th normal of the beta stream (seed, obs 2^64-1, call): a normal always takes three
// words, so the k-th one starts at word 3k -- counter-based generation lets every lane
// jump straight to its own.
__device__ inline double stream_normal_obs(uint64_t seed, uint64_t obs, uint32_t call, int k)
{
    PhiloxSource s;
    s.open(seed, obs, call);
    s.blk = (uint32_t)(3 * k) >> 2;
    for (int skip = (3 * k) & 3; skip > 0; --skip) s.word();
    return s.norm();
}

// Stream of the constrained draw's rejection normals (below): (seed, obs 2^64-3, call).
constexpr uint64_t kTnObs = 0xFFFFFFFFFFFFFFFDull;

// Rejection normal number m generated on the spot -- only when the precomputed ones have run out; out of line so that
// the Philox rounds and the Box-Muller transform stay out of the coordinate loop's instruction stream.
static __device__ __noinline__ double tn_normal_spot(uint64_t seed, uint32_t call, int m)
{
    return stream_normal_obs(seed, kTnObs, call, m);
}

__device__ inline double stream_normal(uint64_t seed, uint32_t call, int k)
{
    PhiloxSource s;
    s.open(seed, 0xFFFFFFFFFFFFFFFFull, call);
    s.blk = (uint32_t)(3 * k) >> 2;
    for (int skip = (3 * k) & 3; skip > 0; --skip) s.word();
    return s.norm();
}

// ---- fast path pieces for the plain draw -----------------------------------------------------
// Root-free factorisation A = R' D^-1 R (row j of R is row j of the running Schur complement,
// D = diag(R)); U = D^-1/2 R is the Cholesky factor and is never formed -- the solves below use
// R and rd = 1/diag(R) directly.  Blocked by panels of 8 rows so that the CTA meets at three
// barriers per panel instead of one per column (a column-at-a-time sweep spent ~1000 cycles per
// column, 32 us at P = 64, mostly waiting at barriers with one division on the critical path):
//   (a) warp 0 factorises the 8 x 8 diagonal block (the only sequential part: 8 reciprocals);
//   (b) one thread per column right of the block finishes the 8 panel rows of that column
//       (forward substitution with the block, no barrier inside);
//   (c) all threads apply the rank-8 update to the trailing matrix.
// 1 / d for a positive normal d: hardware seed (20 bits) and two Newton steps, ~6 dependent operations
// instead of the ~25 of an IEEE division -- eight of these sit on the critical path of every panel.  Within
// an ulp of 1 / d; every rank runs the same code, so the replicated draws still agree bit for bit.
__device__ __forceinline__ double rcp_newton(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// `normals` (optional): e[i] = stream_normal(seed, call, rev ? P - 1 - i : i) is produced by the warps that would
// otherwise idle while warp 0 factorises the first diagonal block.
__device__ __forceinline__ void cta_ldl_upper(double *A, double *rd, int P, int ld, int *ok, double *normals = nullptr,
                                              uint64_t seed = 0, uint32_t call = 0, bool rev = false)
{
    constexpr int NB = 8;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int k0 = 0; k0 < P; k0 += NB) {
        const int nb = P - k0 < NB ? P - k0 : NB;
        const int kend = k0 + nb;
        if (tid < 32) {
            // The 8 x 8 diagonal block lives in the warp's registers for its eight elimination steps: lane
            // (i = lane >> 3, k = lane & 7) holds rows i and i + 4 of column k; a step broadcasts the pivot and the
            // pivot row by shuffles (no shared-memory round trip, no warp barrier).  Rows / columns past the
            // matrix edge act as an identity block.
            const int i = tid >> 3, k = tid & 7;
            double a0 = (i < nb && k < nb) ? A[(k0 + i) + (size_t)ld * (k0 + k)] : (i == k ? 1.0 : 0.0);
            double a1 = (i + 4 < nb && k < nb) ? A[(k0 + i + 4) + (size_t)ld * (k0 + k)] : (i + 4 == k ? 1.0 : 0.0);
            bool bad = false;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int src = (j & 3) * 8;
                const double vj = j < 4 ? a0 : a1;                       // row j as held by lanes src .. src + 7
                const double d = __shfl_sync(0xffffffffu, vj, src + j);
                bad = bad || !(d > 0.0);
                const double inv = rcp_newton(d);
                if (tid == 0 && j < nb) rd[k0 + j] = inv;
                const double s = __shfl_sync(0xffffffffu, vj, src + k) * inv;
                const double rji = __shfl_sync(0xffffffffu, vj, src + i);
                const double rji4 = __shfl_sync(0xffffffffu, vj, src + i + 4);
                if (i > j && i <= k) a0 = fma(-rji, s, a0);
                if (i + 4 > j && i + 4 <= k) a1 = fma(-rji4, s, a1);
            }
            if (bad) { if (tid == 0) *ok = 0; }
            else {
                if (i < nb && k < nb && i <= k) A[(k0 + i) + (size_t)ld * (k0 + k)] = a0;
                if (i + 4 < nb && k < nb && i + 4 <= k) A[(k0 + i + 4) + (size_t)ld * (k0 + k)] = a1;
            }
        } else if (k0 == 0 && normals) {
            for (int m = tid - 32; m < P; m += nt - 32) normals[m] = stream_normal(seed, call, rev ? P - 1 - m : m);
        }
        __syncthreads();
        if (!*ok) return;
        for (int k = kend + tid; k < P; k += nt) {
            double r[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) r[j] = j < nb ? A[(k0 + j) + (size_t)ld * k] : 0.0;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                if (j < nb) {
                    double s = r[j] * rd[k0 + j];
#pragma unroll
                    for (int j2 = j + 1; j2 < NB; ++j2)
                        if (j2 < nb) r[j2] = fma(-A[(k0 + j) + (size_t)ld * (k0 + j2)], s, r[j2]);
                }
            }
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (j < nb) A[(k0 + j) + (size_t)ld * k] = r[j];
        }
        __syncthreads();
        for (int k = kend + (tid >> 4); k < P; k += nt >> 4) {
            double s[NB];
#pragma unroll
            for (int j = 0; j < NB; ++j) s[j] = j < nb ? A[(k0 + j) + (size_t)ld * k] * rd[k0 + j] : 0.0;
            for (int i = kend + (tid & 15); i <= k; i += 16) {
                double acc = A[i + (size_t)ld * k];
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if (j < nb) acc = fma(-A[(k0 + j) + (size_t)ld * i], s[j], acc);
                A[i + (size_t)ld * k] = acc;
            }
        }
        __syncthreads();
    }
}

// One warp, x held in registers (lane l owns x[l], x[l+32], ...; up to 8 per lane = P <= 256).
// Column-oriented substitution: the pivot value travels by one shuffle, every lane then
// updates its own entries -- no reduction on the critical path.
//   forward : y = U^-T b   with U = D^-1/2 R  ->  y_i = sqrt(rd_i) (b_i - sum_{k<i} R[k,i] sqrt(rd_k) y_k)
//   backward: x = U^-1 y                         x_i = sqrt(rd_i) (y_i - sum_{k>i} R[i,k] sqrt(rd_i) x_k) ...
// written below in terms of t = D^-1/2-scaled quantities so only R and rd are touched.
template <int KP>
__device__ __forceinline__ void warp_solve_plain_k(const double *R, const double *rd, const double *b,
                                                   const double *e, double *out, int P, int ld, int lane, bool rev)
{
    double x[KP];
#pragma unroll
    for (int q = 0; q < KP; ++q) x[q] = lane + 32 * q < P ? b[lane + 32 * q] : 0.0;
    // forward: t_i = rd_i * (b_i - sum_{k<i} R[k,i] t_k); lanes keep the running residuals
    for (int i = 0; i < P; ++i) {
        double own = x[0];
#pragma unroll
        for (int q = 1; q < KP; ++q) own = (i >> 5) == q ? x[q] : own;
        double rr[KP];
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            int m = lane + 32 * q;
            rr[q] = (m > i && m < P) ? R[i + (size_t)ld * m] : 0.0;      // issued before the shuffle lands
        }
        double gi = __shfl_sync(0xffffffffu, own, i & 31) * rd[i];
#pragma unroll
        for (int q = 0; q < KP; ++q) x[q] = fma(-rr[q], gi, x[q]);
    }
    // R x = r + e / sqrt(rd)  (see the derivation above), backward substitution
#pragma unroll
    for (int q = 0; q < KP; ++q) {
        int m = lane + 32 * q;
        if (m < P) x[q] = x[q] + e[m] / sqrt(rd[m]);
    }
    for (int i = P - 1; i >= 0; --i) {
        double own = x[0];
#pragma unroll
        for (int q = 1; q < KP; ++q) own = (i >> 5) == q ? x[q] : own;
        double rr[KP];
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            int m = lane + 32 * q;
            rr[q] = m < i ? R[m + (size_t)ld * i] : 0.0;
        }
        double xi = __shfl_sync(0xffffffffu, own, i & 31) * rd[i];
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            int m = lane + 32 * q;
            x[q] = m == i ? xi : fma(-rr[q], xi, x[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < KP; ++q) {
        int m = lane + 32 * q;
        if (m < P) out[rev ? P - 1 - m : m] = x[q];      // rev: solution of the index-reversed system
    }
}

__device__ inline void warp_solve_plain(const double *R, const double *rd, const double *b,
                                        const double *e, double *out, int P, int ld, int lane, bool rev = false)
{
    if (P <= 64) warp_solve_plain_k<2>(R, rd, b, e, out, P, ld, lane, rev);
    else if (P <= 128) warp_solve_plain_k<4>(R, rd, b, e, out, P, ld, lane, rev);
    else warp_solve_plain_k<8>(R, rd, b, e, out, P, ld, lane, rev);
}


// ---- plain / mvn draw for P <= 64: blocked factorisation with look-ahead ----------------------------------------
// The panel loop of cta_ldl_upper costs ~3 900 cycles per 8 columns at P = 64 (BL_BETA_CLOCKS: 31 000 cycles for the
// factorisation, 12 600 for the two substitutions), most of it one warp's diagonal-block recurrence with the other
// seven waiting, then everyone's trailing update with that warp's result.  Here:
//   * the matrix is padded to a multiple of 8 with an identity block, so every panel is a full 8 x 8;
//   * the right-hand side rides along as column Pp: the panel step applied to it IS the forward substitution;
//   * the trailing update runs on the FP64 tensor path, one 8 x 8 tile = two DMMA m8n8k4 (A = the panel's rows,
//     B = the same rows scaled by -1/d, kept in a side buffer by the panel step);
//   * look-ahead: warp 0 updates the NEXT diagonal tile first and factorises it while warps 1-7 finish the other
//     tiles, so the serial recurrence (8 pivots x ~185 cycles) overlaps the bulk of the update.
// Workspace W: F [LD x (Pp + 8)] column-major, LD = Pp + 1 (odd); Sn [8 x 76]; rd [64]; e [64].
constexpr int kFastLdS = 76;                  // row stride of Sn: 4 tig + gid distinct mod 16 -> conflict-free B fragments

__host__ __device__ inline size_t beta_fast_doubles(int P)           // F, Sn, rd
{
    const int Pp = (P + 7) & ~7;
    return (size_t)(Pp + 1) * (Pp + 8) + 8 * kFastLdS + 64;
}
// doubles in front of the scratch of the constrained draw's fast set-up (192 + Pp^2 doubles: cta_constrained_setup_fast)
__host__ __device__ inline size_t beta_tn_scratch_offset(int P)
{
    const size_t plain = 2 * (size_t)(P | 1) * P + 5 * (size_t)P + (size_t)P * P + 2 * (size_t)P, fast = beta_fast_doubles(P);
    return plain > fast ? plain : fast;
}
__host__ __device__ inline size_t beta_tn_doubles(int P)
{
    const int Pp = (P + 7) & ~7;
    return P <= 64 ? beta_tn_scratch_offset(P) + 192 + (size_t)(Pp + 4) * Pp
                   : 2 * (size_t)(P | 1) * P + 5 * (size_t)P + (size_t)P * P + 2 * (size_t)P;
}
// doubles of the CTA's workspace in front of the fast path's normals e[64]: the longer of the two layouts
__host__ __device__ inline size_t beta_fast_e_offset(int P)
{
    const size_t plain = 2 * (size_t)(P | 1) * P + 5 * (size_t)P, fast = beta_fast_doubles(P);
    return plain > fast ? plain : fast;
}

__device__ __forceinline__ void beta_dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 1 / d as rcp_newton, with ONE third-order step on the 20-bit hardware seed (error e^3 < 2^-58): three dependent
// operations instead of four -- it sits on the critical path of every pivot.
__device__ __forceinline__ double rcp_newton3(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// The 8 x 8 diagonal block at (k0, k0) in one warp's registers: lane (i = lane >> 3, k = lane & 7) holds rows i and
// i + 4 of column k; a step broadcasts the pivot and the pivot row by shuffles.  Writes the block's rows of R and rd.
__device__ __forceinline__ bool warp_ldl_diag8(double *F, double *rd, int k0, int LD, int lane)
{
    const int i = lane >> 3, k = lane & 7;
    double a0 = F[(k0 + i) + LD * (k0 + k)];
    double a1 = F[(k0 + i + 4) + LD * (k0 + k)];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int src = (j & 3) * 8;
        const double vj = j < 4 ? a0 : a1;
        const double d = __shfl_sync(0xffffffffu, vj, src + j);
        bad = bad || !(d > 0.0);
        const double inv = rcp_newton3(d);
        if (lane == 0) rd[k0 + j] = inv;
        const double s = __shfl_sync(0xffffffffu, vj, src + k) * inv;
        const double rji = __shfl_sync(0xffffffffu, vj, src + i);
        const double rji4 = __shfl_sync(0xffffffffu, vj, src + i + 4);
        if (i > j && i <= k) a0 = fma(-rji, s, a0);
        if (i + 4 > j && i + 4 <= k) a1 = fma(-rji4, s, a1);
    }
    if (i <= k) F[(k0 + i) + LD * (k0 + k)] = a0;
    if (i + 4 <= k) F[(k0 + i + 4) + LD * (k0 + k)] = a1;
    return !bad;
}

// Workspace split of the fast path (doubles from the start of the CTA's workspace); the normals e[64] live behind
// whichever of the two layouts is longer (k_beta_draw computes them before the matrix arrives).
struct BetaFast {
    double *F, *Sn, *rd;
    int Pp, LD, nblk;
    __device__ __forceinline__ BetaFast(double *W, int P)
    {
        Pp = (P + 7) & ~7; LD = Pp + 1; nblk = Pp >> 3;
        F = W; Sn = F + LD * (Pp + 8); rd = Sn + 8 * kFastLdS;
    }
};

// A (ld, P x P, full symmetric) and rhs, both in the CTA's workspace, -> the padded layout, which overlays them:
// through registers, element (i = tid & 63, column (tid >> 6) + 4 u).  256 threads.
__device__ __forceinline__ void beta_fast_relayout(double *W, const double *A, int ld, const double *rhs, int P)
{
    const BetaFast w(W, P);
    const int tid = threadIdx.x, i = tid & 63, cb = tid >> 6;
    double g[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        const int c = cb + 4 * u;
        g[u] = (i < P && c < P) ? A[i + (size_t)ld * c] : (i == c ? 1.0 : 0.0);
    }
    const double rr = tid < P ? rhs[tid] : 0.0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        const int c = cb + 4 * u;
        if (i < w.Pp && c < w.Pp) w.F[i + w.LD * c] = g[u];
    }
    if (tid < w.Pp) w.F[tid + w.LD * w.Pp] = rr;
    if (tid < 56) w.Sn[(tid / 7) * kFastLdS + w.Pp + 1 + tid % 7] = 0.0;       // the rhs tile's seven unused columns
    __syncthreads();
}

// The factorisation and both substitutions on the padded layout (F, rhs column and Sn padding in place; the caller
// has synchronised).  256 threads: warps 0-6 factorise (named barrier 1 over their 224 threads), warp 7 produces the
// P normals of the draw meanwhile (~5 000 cycles of one thread's latency, needed only by the backward substitution).
// `rev`: the system is the index-reversed one (mvn draw).
//   kDraw = false: the solve alone (no normals, R left unscaled for the caller) -- the set-up of the constrained draw.
template <bool kDraw = true>
__device__ __forceinline__ void cta_plain_fast(double *W, double *e, double *beta_out, int P, bool rev, int *ok,
                                               uint64_t seed, uint32_t call)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const BetaFast w(W, P);
    double *F = w.F, *Sn = w.Sn, *rd = w.rd;
    const int Pp = w.Pp, LD = w.LD, nblk = w.nblk;
#ifdef BL_BETA_CLOCKS
    long long ck[40]; int nck = 0;
    ck[nck++] = clock64();
#define BL_CK() ck[nck++] = clock64()
#else
#define BL_CK()
#endif
#define BL_BAR7() asm volatile("bar.sync 1, 224;" ::: "memory")
    if (warp == 7) {
        if (kDraw)
            for (int m = lane; m < Pp; m += 32) e[m] = m < P ? stream_normal(seed, call, rev ? P - 1 - m : m) : 0.0;
    } else {
        if (warp == 0 && !warp_ldl_diag8(F, rd, 0, LD, lane)) *ok = 0;
        BL_CK();
        BL_BAR7();
        const int gid = lane >> 2, tig = lane & 3;
        for (int p = 0; p < nblk; ++p) {
            const int k0 = 8 * p, kend = k0 + 8;
            // panel rows of every column right of the block, the rhs column (index Pp) included: one thread per column
            {
                const int k = kend + tid;
                if (k <= Pp) {
                    double r[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = F[(k0 + j) + LD * k];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double s = r[j] * rd[k0 + j];
                        Sn[j * kFastLdS + k] = -s;
#pragma unroll
                        for (int j2 = j + 1; j2 < 8; ++j2) r[j2] = fma(-F[(k0 + j) + LD * (k0 + j2)], s, r[j2]);
                    }
#pragma unroll
                    for (int j = 1; j < 8; ++j) F[(k0 + j) + LD * k] = r[j];
                }
            }
            BL_CK();
            BL_BAR7();
            if (!*ok) break;
            if (warp == 0) {
                if (p + 1 < nblk) {
                    // look-ahead: the next diagonal tile, updated in the factorisation's own register layout (two
                    // chains of four FMAs per entry), then factorised while warps 1-6 update the other tiles
                    const int i = lane >> 3, k = lane & 7;
                    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const double s0 = Sn[j * kFastLdS + kend + k], s1 = Sn[(j + 1) * kFastLdS + kend + k];
                        a0 = fma(F[(k0 + j) + LD * (kend + i)], s0, a0);
                        b0 = fma(F[(k0 + j + 1) + LD * (kend + i)], s1, b0);
                        a1 = fma(F[(k0 + j) + LD * (kend + i + 4)], s0, a1);
                        b1 = fma(F[(k0 + j + 1) + LD * (kend + i + 4)], s1, b1);
                    }
                    F[(kend + i) + LD * (kend + k)] += a0 + b0;
                    F[(kend + i + 4) + LD * (kend + k)] += a1 + b1;
                    __syncwarp();
                    BL_CK();
                    if (!warp_ldl_diag8(F, rd, kend, LD, lane)) *ok = 0;
                }
            } else {
                // The block's columns are final and no later step reads them: scale them by their pivots'
                // reciprocals for the backward substitution (a step there is then one shuffle and one FMA:
                // x_m -= (R[m,i] rd_i) x_i(raw), with x_i = rd_i x_i(raw) off the critical path).
                if (kDraw)
                    for (int el = tid - 32; el < 8 * 64; el += 192) {
                        const int col = k0 + (el >> 6), m = el & 63;
                        if (m < col) F[m + LD * col] *= rd[col];
                    }
                // trailing tiles (ti, tk), p < ti <= tk <= nblk (tk = nblk: the rhs column's tile), row block by
                // row block, dealt round-robin to warps 1-6; tile 0 is warp 0's
                const int nrem = nblk - (p + 1);
                const int T = nrem * (nrem + 1) / 2 + nrem;
                int a = 0, base = 0;                                           // base = first tile index of row block a
                for (int q = warp; q < T; q += 6) {
                    while (q >= base + (nrem - a + 1)) { base += nrem - a + 1; ++a; }
                    const int ti = p + 1 + a, tk = ti + (q - base);
                    double *c = F + (8 * ti + gid) + LD * (8 * tk + 2 * tig);
                    double c0 = c[0], c1 = c[LD];
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        beta_dmma884(c0, c1, F[(k0 + 4 * h + tig) + LD * (8 * ti + gid)], Sn[(4 * h + tig) * kFastLdS + 8 * tk + gid]);
                    c[0] = c0; c[LD] = c1;
                }
            }
            BL_CK();
            if (p + 1 < nblk) BL_BAR7();
        }
    }
    __syncthreads();
    BL_CK();
    if (!*ok) return;
    // column Pp holds y = the forward substitution.  Backward: R x = y + e / sqrt(rd), x in one warp's registers.
    if (warp == 0) {
        double x[2], rdl[2];
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int m = lane + 32 * qq;
            rdl[qq] = m < Pp ? rd[m] : 1.0;
            x[qq] = m < Pp ? F[m + LD * Pp] + (kDraw ? e[m] / sqrt(rdl[qq]) : 0.0) : 0.0;
        }
        for (int i = Pp - 1; i >= 0; --i) {
            const double own = (i >> 5) ? x[1] : x[0];
            double rr[2];
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
                const int m = lane + 32 * qq;
                rr[qq] = m < i ? F[m + LD * i] : 0.0;
            }
            double xi = __shfl_sync(0xffffffffu, own, i & 31);
            if (!kDraw) xi *= rd[i];                                            // unscaled columns
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) x[qq] = fma(-rr[qq], xi, x[qq]);     // rr is zero at and below the diagonal
        }
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
            const int m = lane + 32 * qq;
            if (m < P) beta_out[rev ? P - 1 - m : m] = x[qq] * rdl[qq];
        }
    }
#ifdef BL_BETA_CLOCKS
    BL_CK();
    if (tid == (call == 3 ? 0 : 32) && (call == 3 || call == 4)) {
        printf("[beta fast clocks] tid %d:", tid);
        for (int q = 1; q < nck; ++q) printf(" %lld", ck[q] - ck[q - 1]);
        printf("\n");
    }
#endif
    __syncthreads();
#undef BL_BAR7
}


// Set-up of the constrained draw for P <= 64 (256 threads): what Logit.hpp:338-365 obtains from two Cholesky
// factorisations, P + 1 pairs of triangular solves and one more triangular solve -- mP = PP^-1 bP,
// L = chol_lower(PP^-1), z = L^-1 (beta_prev - mP), 98 000 + 126 000 + 75 000 + 157 000 cycles as written in
// cta_chol_* / cta_solve_utu / warp_solve_* -- from ONE blocked factorisation of the index-reversed matrix:
// with J the reversal and J PP J = U'U,   PP^-1 = (J U^-1 J)(J U^-1 J)'   and J U^-1 J is lower triangular with a
// positive diagonal, so it IS L.  Hence
//   mP = J (J PP J)^-1 J bP          the blocked solve (cta_plain_fast<false>),
//   L  = J U^-1 J                    an explicit triangular inverse, four threads per column,
//   z  = L^-1 d = J U (J d)          a triangular matrix-vector product, no solve.
// On entry A (ld) holds PP and rhs holds bP; on exit B (ld) holds L (strict upper part zero), A (ld) holds 1 / L
// with zeros where no constraint applies, mP and z are filled.  X (P x P doubles) is scratch -- the rejection
// normals' buffer, which is loaded afterwards.
// `scratch` (192 + (Pp + 4) Pp doubles) lies behind BOTH workspace layouts (beta_tn_scratch_offset): for small P the vectors
// of the plain layout sit inside the padded matrix.
__device__ __forceinline__ void cta_constrained_setup_fast(double *A, double *B, double *mP_out, double *z_out,
                                                           double *rhs, const double *beta_prev, double *scratch,
                                                           int P, int ld, int *ok, uint64_t seed, uint32_t call)
{
    const int tid = threadIdx.x;
    const int PP2 = P * P;
    double *mP = scratch, *z = scratch + 64, *rdv = scratch + 128, *X = scratch + 192;
#ifdef BL_BETA_CLOCKS
    long long sc[6]; sc[0] = clock64();
#endif
    // index reversal in place (as the mvn draw)
    for (int k = tid; k < PP2 / 2; k += blockDim.x) {
        const int k2 = PP2 - 1 - k;
        double *a = A + (k % P) + (size_t)ld * (k / P), *b = A + (k2 % P) + (size_t)ld * (k2 / P);
        const double t = *a; *a = *b; *b = t;
    }
    for (int k = tid; k < P / 2; k += blockDim.x) { const double t = rhs[k]; rhs[k] = rhs[P - 1 - k]; rhs[P - 1 - k] = t; }
    __syncthreads();
    beta_fast_relayout(A, A, ld, rhs, P);
#ifdef BL_BETA_CLOCKS
    sc[1] = clock64();
#endif
    cta_plain_fast<false>(A, nullptr, mP, P, true, ok, seed, call);          // mP in the original order
    if (!*ok) return;
#ifdef BL_BETA_CLOCKS
    sc[2] = clock64();
#endif
    const BetaFast w(A, P);
    const int Pp = w.Pp, LD = w.LD;
    const double *R = w.F;
    if (tid < Pp) rdv[tid] = w.rd[tid];
    __syncthreads();
    // X = U^-1 (upper), U = D^-1/2 R:  x_c = sqrt(rd_c),  x_k = -rd_k sum_{i > k} R[k,i] x_i.  Column c of X belongs to
    // threads 4 c .. 4 c + 3, which keep its running sums r_m = sum_{i > m} R[m,i] x_i for the rows m = t mod 4 in the
    // column's own storage; x_k travels from its owner to the other three by shuffle.
    {
        const int c = tid >> 2, sub = tid & 3;
        double *xc = X + (size_t)(Pp + 4) * c;                                // leading dimension Pp + 4: the quads of a half-warp hit distinct banks
        if (c < Pp)
            for (int m = sub; m < Pp; m += 4) xc[m] = 0.0;
        __syncwarp();
        for (int k = Pp - 1; k >= 0; --k) {
            // every thread of a quad evaluates x_k from the row owner's sum
            double xk = 0.0;
            const int owner = (tid & ~3) | (k & 3);
            double rk = (c < Pp && k <= c && sub == (k & 3)) ? xc[k] : 0.0;
            rk = __shfl_sync(0xffffffffu, rk, owner & 31);
            if (c < Pp && k <= c) {
                xk = k == c ? sqrt(rdv[c]) : -rdv[k] * rk;
                if (sub == (k & 3)) xc[k] = xk;
                int m = sub;
                for (; m + 12 < k; m += 16) {                                    // four rows at a time: their loads go out together
                    const double r0 = R[m + LD * k], r1 = R[m + 4 + LD * k], r2 = R[m + 8 + LD * k], r3 = R[m + 12 + LD * k];
                    const double x0 = xc[m], x1 = xc[m + 4], x2 = xc[m + 8], x3 = xc[m + 12];
                    xc[m] = fma(r0, xk, x0); xc[m + 4] = fma(r1, xk, x1); xc[m + 8] = fma(r2, xk, x2); xc[m + 12] = fma(r3, xk, x3);
                }
                for (; m < k; m += 4) xc[m] = fma(R[m + LD * k], xk, xc[m]);
            }
        }
    }
    __syncthreads();
#ifdef BL_BETA_CLOCKS
    sc[3] = clock64();
#endif
    // z = J U (J d), d = beta_prev - mP: row i of U against the reversed d; thread i (four partial sums)
    if (tid < P) {
        const int i = tid;
        const double sq = sqrt(rdv[i]);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = i + 1;
        for (; k + 3 < P; k += 4) {
            s0 = fma(R[i + LD * k], beta_prev[P - 1 - k] - mP[P - 1 - k], s0);
            s1 = fma(R[i + LD * (k + 1)], beta_prev[P - 2 - k] - mP[P - 2 - k], s1);
            s2 = fma(R[i + LD * (k + 2)], beta_prev[P - 3 - k] - mP[P - 3 - k], s2);
            s3 = fma(R[i + LD * (k + 3)], beta_prev[P - 4 - k] - mP[P - 4 - k], s3);
        }
        for (; k < P; ++k) s0 = fma(R[i + LD * k], beta_prev[P - 1 - k] - mP[P - 1 - k], s0);
        // U[i,i] = 1 / sqrt(rd_i), U[i,k] = sqrt(rd_i) R[i,k]
        z[P - 1 - i] = (beta_prev[P - 1 - i] - mP[P - 1 - i]) / sq + sq * ((s0 + s1) + (s2 + s3));
    }
    __syncthreads();
#ifdef BL_BETA_CLOCKS
    sc[4] = clock64();
#endif
    // L[j,c] = X[P-1-j, P-1-c] into B, 1 / L into A (both ld); element (j = tid & 63, column (tid >> 6) + 4 u)
    {
        const int j = tid & 63, cb = tid >> 6;
        double g[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int c = cb + 4 * u;
            g[u] = (j < P && c < P && j >= c) ? X[(P - 1 - j) + (size_t)(Pp + 4) * (P - 1 - c)] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int c = cb + 4 * u;
            if (j < P && c < P) {
                B[j + (size_t)ld * c] = g[u];
                A[j + (size_t)ld * c] = (j >= c && j < P - 1 && g[u] != 0.0) ? 1.0 / g[u] : 0.0;
            }
        }
        if (tid < P) { mP_out[tid] = mP[tid]; z_out[tid] = z[tid]; }
    }
    __syncthreads();
#ifdef BL_BETA_CLOCKS
    if (tid == 0 && call == 3) printf("[beta constrained set-up clocks] reverse + layout %lld solve %lld inverse %lld z %lld L, 1/L %lld\n", sc[1] - sc[0], sc[2] - sc[1], sc[3] - sc[2], sc[4] - sc[3], clock64() - sc[4]);
#endif
}

// Exact max / min of a double over the warp in two 32-bit hardware reductions (redux.sync) on an
// order-preserving integer key, instead of a five-stage butterfly of 64-bit shuffles and compares
// (the sweeps below do one of each per coordinate, on the critical path).
__device__ __forceinline__ unsigned long long ordered_key(double v)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_value(unsigned long long k)
{
    const unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)u);
}
__device__ __forceinline__ double warp_max_f64(double v)
{
    const unsigned long long k = ordered_key(v);
    const unsigned hi = (unsigned)(k >> 32);
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? (unsigned)k : 0u);
    return ordered_value(((unsigned long long)mhi << 32) | mlo);
}
__device__ __forceinline__ double warp_min_f64(double v)
{
    const unsigned long long k = ordered_key(v);
    const unsigned hi = (unsigned)(k >> 32);
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? (unsigned)k : 0xFFFFFFFFu);
    return ordered_value(((unsigned long long)mhi << 32) | mlo);
}

// The coordinate-wise constrained draw, Logit.hpp:366-399, on one warp with beta and z held in
// registers (lane l owns entries l, l + 32, ...), 1 / L precomputed (iL), and the truncated normal
// split across lanes: what sits between two consecutive updates is one shuffle (z1), one FMA per
// owned entry, a 5-stage min/max butterfly, the truncated normal, and one FMA per owned entry.
// `is` is the permutation scratch (shared memory); all lanes draw every variate so their 32 copies
// of the stream stay in step.
//
// The truncated normal itself.  Inverse CDF costs an erfc pair and an inverse normal CDF of dependent FP64 latency
// (~2 200 cycles on one lane) per coordinate, P^2 times per beta draw.  But in a well-identified model nearly every
// window (cmin, cmax) is wide -- the constraint beta_j >= 0 binds only for coefficients near zero -- and for a wide
// window plain rejection from N(0,1) is exact and almost always accepts at the first try.  So: one normal from the
// rejection stream (seed, obs 2^64-3, call) -- normal number m at words 3m..3m+2, PRECOMPUTED in parallel by the whole
// CTA before the sweeps (nbuf; past its end they are generated on the spot) -- is ALWAYS tried first; if it misses and
// cmin < 1, cmax > -1 and cmax - cmin >= 1/2, up to three more; otherwise, or after four misses, the inverse-CDF /
// tail sampler on the beta stream as before.  Every branch returns an exact truncated normal, so the
// draw's distribution is unchanged; the oracle (draw_beta_constrained) mirrors the rule variate for variate.
// One sweep's random order (Logit.hpp:367-374): is[i] <-> is[r.flat(i, P)] for i = 0 .. P-2, one warp.
// r.flat(i, P) takes one stream word each, words a0 .. a0 + P - 2 of the beta stream: lane l forms the swap targets of
// i = l, l + 32, ... straight from the counter (the block that holds its word) and lane 0 is left with the swaps
// alone -- drawn one after the other by every lane this was a third of the whole constrained draw (BL_BETA_CLOCKS:
// 15 500 of 49 000 cycles per sweep at P = 64).  `src` only supplies the stream's key and counter words.
__device__ __forceinline__ void warp_sweep_permutation_at(int *is, int *tt, int a0, const PhiloxSource &src, int P, int lane)
{
    for (int i = lane; i < P - 1; i += 32) {
        const int a = a0 + i;
        const uint4 b = philox_block_ool(src.c0, src.c1, (uint32_t)(a >> 2), src.c3, src.key);
        const uint32_t wv = (a & 3) == 0 ? b.x : (a & 3) == 1 ? b.y : (a & 3) == 2 ? b.z : b.w;
        const double f = (double)i + ((double)P - (double)i) * word_to_unif(wv);
        unsigned t = (unsigned)f;
        if (t > (unsigned)(P - 1)) t = (unsigned)(P - 1);
        tt[i] = (int)t;
    }
    __syncwarp();
    if (lane == 0) {
        // the swaps themselves are sequential (a target may have been moved by an earlier swap); their
        // targets are fetched eight at a time so that only the two loads of a swap wait on each other
        int i = 0;
        for (; i + 8 <= P - 1; i += 8) {
            int t8[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) t8[r] = tt[i + r];
#pragma unroll
            for (int r = 0; r < 8; ++r) { const int a = is[i + r], b = is[t8[r]]; is[i + r] = b; is[t8[r]] = a; }
        }
        for (; i < P - 1; ++i) { const int t = tt[i]; const int a = is[i], b = is[t]; is[i] = b; is[t] = a; }
    }
    __syncwarp();
}
// absolute index of the stream's next word / the stream moved to an absolute word index
__device__ __forceinline__ int philox_tell(const PhiloxSource &src) { return 4 * ((int)src.blk - 1) + src.pos; }
__device__ __forceinline__ void philox_seek(PhiloxSource &src, int a)
{
    src.buf = philox_block_ool(src.c0, src.c1, (uint32_t)(a >> 2), src.c3, src.key);
    src.blk = (uint32_t)(a >> 2) + 1u;
    src.pos = a & 3;
}
__device__ __forceinline__ void warp_sweep_permutation(int *is, PhiloxSource &src, int P, int lane)
{
    const int a0 = philox_tell(src);
    warp_sweep_permutation_at(is, is + P, a0, src, P, lane);
    philox_seek(src, a0 + P - 1);
}

// Coordinates i0 .. i1-1 of one sweep (in the order is[]), one warp, beta and z in registers (lane l owns entries l,
// l + 32, ...).  Returns with beta, z, mnorm and the stream advanced.
template <int KP>
__device__ __forceinline__ void warp_seq_range(const double *L, const double *iL, const int *is, int i0, int i1,
                                               double (&beta)[KP], double (&z)[KP], int &mnorm, PhiloxSource &src,
                                               int P, int ld, int lane, const double *nbuf, int nbuf_len,
                                               uint64_t seed, uint32_t call, int &n_fall, int &n_rej)
{
    if (i0 >= i1) return;
    // two register sets (current / next coordinate) used alternately: no copies at the end of an iteration
    struct Col { double lu[KP], il[KP]; bool pos[KP], neg[KP]; double z1; int c; };
    auto fetch = [&](Col &k, int c) {
        k.c = c;
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            const int j = lane + 32 * q;
            k.lu[q] = j < P ? L[j + (size_t)ld * c] : 0.0;               // L is zero above its diagonal
            k.il[q] = j < P ? iL[j + (size_t)ld * c] : 0.0;              // zero where no constraint applies
            k.pos[q] = k.il[q] > 0.0;
            k.neg[q] = k.il[q] < 0.0;
        }
        double zc = z[0];
#pragma unroll
        for (int q = 1; q < KP; ++q) zc = (c >> 5) == q ? z[q] : zc;
        k.z1 = __shfl_sync(0xffffffffu, zc, c & 31);
    };
    double Z0 = mnorm < nbuf_len ? nbuf[mnorm] : tn_normal_spot(seed, call, mnorm);
    auto step = [&](const Col &k, Col &kn, int i) {
        // the next coordinate: index, column, signs, its own current value (stale and unused when i + 1 == P),
        // and its normal assuming a hit
        fetch(kn, i + 1 < i1 ? is[i + 1] : k.c);
        const double Zn = mnorm + 1 < nbuf_len ? nbuf[mnorm + 1] : 0.0;
        const double z1 = k.z1, u = Z0 - z1;
        double bnew[KP];
        bool inside = true;
#pragma unroll
        for (int q = 0; q < KP; ++q) {
            const double v = fma(beta[q], k.il[q], u);
            inside = inside & ((v > 0.0) | !k.pos[q]) & ((v < 0.0) | !k.neg[q]);     // bitwise: no short-circuit branches
            bnew[q] = fma(k.lu[q], u, beta[q]);
        }
        ++mnorm;
        double z2 = Z0;
        const bool hit = __all_sync(0xffffffffu, inside);
        if (__builtin_expect(hit, 1)) {
#pragma unroll
            for (int q = 0; q < KP; ++q) beta[q] = bnew[q];
        } else {
            ++n_fall;
            double cmin = -INFINITY, cmax = INFINITY;
#pragma unroll
            for (int q = 0; q < KP; ++q) {
                const double c1 = fma(-beta[q], k.il[q], z1);
                if (k.pos[q] && c1 > cmin) cmin = c1;
                else if (k.neg[q] && c1 < cmax) cmax = c1;
            }
            cmin = warp_max_f64(cmin);
            cmax = warp_min_f64(cmax);
            bool got = false;
            if (cmin < cmax && cmin < 1.0 && cmax > -1.0 && cmax - cmin >= 0.5) {
                for (int tr = 1; tr < 4 && !got; ++tr) {
                    const double Z = mnorm < nbuf_len ? nbuf[mnorm] : tn_normal_spot(seed, call, mnorm);
                    ++mnorm;
                    if (Z > cmin && Z < cmax) { z2 = Z; got = true; }
                }
            }
            if (!got) ++n_rej;
            if (!got) z2 = tnorm_std_warp(src, cmin, cmax, lane);
            const double dz = z2 - z1;
#pragma unroll
            for (int q = 0; q < KP; ++q) beta[q] = fma(k.lu[q], dz, beta[q]);
        }
#pragma unroll
        for (int q = 0; q < KP; ++q)
            if (lane + 32 * q == k.c) z[q] = z2;
        // the next coordinate's normal: the prefetched one after a hit (mnorm advanced by exactly one)
        Z0 = hit ? Zn : nbuf[mnorm < nbuf_len ? mnorm : 0];
        if (__builtin_expect(mnorm >= nbuf_len && i + 1 < i1, 0)) Z0 = tn_normal_spot(seed, call, mnorm);
    };
    Col ka, kb;
    fetch(ka, is[i0]);
    int i = i0;
    for (; i + 1 < i1; i += 2) {
        step(ka, kb, i);
        step(kb, ka, i + 1);
    }
    if (i < i1) step(ka, kb, i);
}

template <int KP>
__device__ __forceinline__ void warp_constrained_sweeps(const double *L, const double *iL, const double *z_in,
                                                        const double *beta_prev, double *beta_out, int *is,
                                                        PhiloxSource &src, int P, int ld, int lane,
                                                        const double *nbuf, int nbuf_len, uint64_t seed, uint32_t call)
{
    int mnorm = 0;                         // normals of the rejection stream consumed so far (warp-uniform)
    int n_fall = 0, n_rej = 0;          // first normals that missed, inverse-CDF draws (BL_BETA_CLOCKS prints them)
#ifdef BL_BETA_CLOCKS
    long long t_perm = 0, t_coord = 0;
#endif
    double beta[KP], z[KP];
#pragma unroll
    for (int q = 0; q < KP; ++q) {
        int m = lane + 32 * q;
        beta[q] = m < P ? beta_prev[m] : 0.0;
        z[q] = m < P ? z_in[m] : 0.0;
    }
    for (int i = lane; i < P; i += 32) is[i] = i;
    __syncwarp();
    for (int k = 0; k < P; ++k) {
#ifdef BL_BETA_CLOCKS
        long long tp0 = clock64();
#endif
        warp_sweep_permutation(is, src, P, lane);
        // Software-pipelined over the coordinates.  Everything of coordinate i + 1 that does not depend on the draw
        // of coordinate i is fetched while coordinate i is decided: its index, its column of L and 1 / L, the signs
        // of that column, its own current value z1 (each index occurs once per sweep, so z[c'] is not touched by
        // coordinate c) and -- assuming the common outcome, first normal accepted -- its rejection normal.  The
        // first normal Z0 is tried against the constraints lane by lane in a branch-free form,
        //     Z0 inside lane j's bound  <=>  sign(L_jc) (beta_j / L_jc + (Z0 - z1)) > 0,
        // one FMA and two compares per owned entry and ONE warp vote (Z0 lies in the window iff it clears every
        // lane's own bounds); the updated beta = beta + L_c (Z0 - z1) is formed beside the test and kept when the
        // vote passes.  The critical path per coordinate is then FMA -> compare -> vote -> select (~100 cycles; the
        // window-forming version with its two exact max / min reductions spent ~520).  The window itself is only
        // formed on a miss.
#ifdef BL_BETA_CLOCKS
        long long tp1 = clock64(); t_perm += tp1 - tp0;
#endif
        warp_seq_range<KP>(L, iL, is, 0, P, beta, z, mnorm, src, P, ld, lane, nbuf, nbuf_len, seed, call, n_fall, n_rej);
#ifdef BL_BETA_CLOCKS
        t_coord += clock64() - tp1;
#endif
    }
#ifdef BL_BETA_CLOCKS
    if (lane == 0 && call == 3) printf("[beta constrained clocks] permutation %lld coordinates %lld (first normal missed %d times, %d inverse-CDF draws)\n", t_perm, t_coord, n_fall, n_rej);
#endif
#pragma unroll
    for (int q = 0; q < KP; ++q) {
        int m = lane + 32 * q;
        if (m < P) beta_out[m] = beta[q];
    }
}

// The P sweeps of the constrained draw on the whole CTA (P <= 64, 256 threads), SPECULATING on the common outcome:
// in a well-identified model nearly every coordinate accepts its first rejection normal, and then a sweep is nothing
// but the chain  beta <- beta + L_c (Z - z_c)  over its P coordinates with one test per coordinate that always
// passes.  So, per sweep (warps 0-6, named barrier 1 over their 224 threads):
//   * u_i = Z_i - z_{c_i} for all remaining coordinates at once (a coordinate occurs once per sweep, so its z is
//     the committed one; a hit consumes exactly one normal, so Z_i is normal number mnorm + i - i_s);
//   * warp w takes a block of the remaining coordinates: it replays the chain of the coordinates before its block
//     -- the same FMAs in the same order as the sequential loop, so the same bits -- and then tests and applies its
//     own; the first miss of the sweep is found with a shared-memory minimum;
//   * no miss: the last block's chain is the sweep's result.  A miss at i_f: warp 0 replays the chain up to i_f,
//     decides coordinate i_f the long way (window, further tries, inverse CDF -- warp_seq_range on that one
//     coordinate) and the speculation restarts behind it.  After three misses in one sweep the rest of the sweep
//     (and the whole next sweep, if this one had more than six) runs in the sequential loop: models whose
//     constraints bind pay one wasted pass per sweep.
// Warp 7 meanwhile prepares the NEXT sweep's order (the 63 sequential swaps are ~4 300 cycles, as long as a sweep's
// passes), from the stream position the beta stream will have if this sweep needs no inverse-CDF draw; warp 0
// checks that at the end of the sweep and redoes the order from the true position otherwise.
// Identical variates, identical operation order, hence bit-identical to warp_constrained_sweeps (BL_BETA_NO_SPEC
// selects that one; test_constrained_draw_speculation_is_bit_identical).
__device__ __forceinline__ void cta_constrained_sweeps_spec(const double *L, const double *iL, const double *z_in,
                                                            const double *beta_prev, double *beta_out, int *is,
                                                            double *work, PhiloxSource &src, int P, int ld,
                                                            const double *nbuf, int nbuf_len, uint64_t seed, uint32_t call)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double *sb = work, *sz = work + 64, *su = work + 128, *sbf = work + 192;       // committed beta, z; u; a pass's result
    int *ctl = reinterpret_cast<int *>(work + 256);      // [0] first miss, [1] mnorm, [2] i_s, [3] mode, [4] stream position, [5] miss resolved by its finder
    int *isv[2] = {is, is + 2 * P};                      // this sweep's order / the next one's; swap targets behind each
    int n_fall = 0, n_rej = 0;
#define BL_BAR7() asm volatile("bar.sync 1, 224;" ::: "memory")
#ifdef BL_BETA_CLOCKS
    long long t_pass = 0, t_miss = 0, t_redo = 0; int n_pass = 0, n_redo = 0;
#endif
    if (tid < 64) {
        sb[tid] = tid < P ? beta_prev[tid] : 0.0;
        sz[tid] = tid < P ? z_in[tid] : 0.0;
        if (tid < P) is[tid] = tid;
    }
    if (tid == 0) { ctl[1] = 0; ctl[3] = 0; }
    __syncthreads();
    if (warp == 0) {
        warp_sweep_permutation(isv[0], src, P, lane);
        if (lane == 0) { ctl[2] = 0; ctl[4] = philox_tell(src); }
    }
    __syncthreads();
    for (int k = 0; k < P; ++k) {
        const int *cur = isv[k & 1];
        int *nxt = isv[(k & 1) ^ 1];
        if (warp == 7) {
            if (k + 1 < P) {
                for (int i = lane; i < P; i += 32) nxt[i] = cur[i];
                __syncwarp();
                warp_sweep_permutation_at(nxt, nxt + P, ctl[4], src, P, lane);
            }
        } else {
            int misses = 0;
            for (;;) {
                const int i_s = ctl[2], mnorm = ctl[1];
                if (i_s >= P) break;
                const bool sequential = ctl[3] != 0 || misses >= 3 || mnorm + (P - i_s) > nbuf_len;
                if (sequential) {
                    // the rest of this sweep in the one-warp loop
                    if (warp == 0) {
                        double beta[2], z[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) { beta[q] = sb[lane + 32 * q]; z[q] = sz[lane + 32 * q]; }
                        int mn = mnorm, nf = 0;
                        warp_seq_range<2>(L, iL, cur, i_s, P, beta, z, mn, src, P, ld, lane, nbuf, nbuf_len, seed, call, nf, n_rej);
#pragma unroll
                        for (int q = 0; q < 2; ++q) { sb[lane + 32 * q] = beta[q]; sz[lane + 32 * q] = z[q]; }
                        if (lane == 0) { ctl[1] = mn; ctl[2] = P; ctl[3] = (misses + nf) > 6 ? 1 : 0; }
                        n_fall += nf;
                    }
                    BL_BAR7();
                    break;
                }
                // ---- one speculative pass over the coordinates [i_s, P) ----
#ifdef BL_BETA_CLOCKS
                long long q1 = clock64(); ++n_pass;
#endif
                if (tid < P - i_s) su[i_s + tid] = nbuf[mnorm + tid] - sz[cur[i_s + tid]];
                if (tid == 0) { ctl[0] = P; ctl[5] = 0; }
                BL_BAR7();
                double b0 = 0.0, b1 = 0.0;                 // this warp's chain: beta in front of coordinate miss_at on a miss
                int miss_at = -1;
                const bool v0 = lane < P, v1 = lane + 32 < P;
                {
                    const int blk = (P - i_s + 6) / 7;
                    const int a = i_s + blk * warp, b = min(a + blk, P);
                    if (a < P) {
                        b0 = sb[lane]; b1 = sb[lane + 32];
                        // the chain of the coordinates before this warp's block, four columns' loads in flight
                        int i = i_s;
                        for (; i + 4 <= a; i += 4) {
                            int c[4]; double u[4], l0[4], l1[4];
#pragma unroll
                            for (int r = 0; r < 4; ++r) { c[r] = cur[i + r]; u[r] = su[i + r]; }
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                l0[r] = v0 ? L[lane + (size_t)ld * c[r]] : 0.0;
                                l1[r] = v1 ? L[lane + 32 + (size_t)ld * c[r]] : 0.0;
                            }
#pragma unroll
                            for (int r = 0; r < 4; ++r) { b0 = fma(l0[r], u[r], b0); b1 = fma(l1[r], u[r], b1); }
                        }
                        for (; i < a; ++i) {
                            const int c = cur[i];
                            const double u = su[i];
                            b0 = fma(v0 ? L[lane + (size_t)ld * c] : 0.0, u, b0);
                            b1 = fma(v1 ? L[lane + 32 + (size_t)ld * c] : 0.0, u, b1);
                        }
                        // this warp's own coordinates: test, then apply
                        bool clean = true;
                        for (i = a; i < b; ++i) {
                            const int c = cur[i];
                            const double u = su[i];
                            const double il0 = v0 ? iL[lane + (size_t)ld * c] : 0.0, il1 = v1 ? iL[lane + 32 + (size_t)ld * c] : 0.0;
                            const double l0 = v0 ? L[lane + (size_t)ld * c] : 0.0, l1 = v1 ? L[lane + 32 + (size_t)ld * c] : 0.0;
                            const double w0 = fma(b0, il0, u), w1 = fma(b1, il1, u);
                            const bool inside = ((w0 > 0.0) | !(il0 > 0.0)) & ((w0 < 0.0) | !(il0 < 0.0)) &
                                                ((w1 > 0.0) | !(il1 > 0.0)) & ((w1 < 0.0) | !(il1 < 0.0));
                            if (!__all_sync(0xffffffffu, inside)) {
                                if (lane == 0) atomicMin(&ctl[0], i);
                                clean = false;
                                miss_at = i;
                                break;
                            }
                            b0 = fma(l0, u, b0);
                            b1 = fma(l1, u, b1);
                        }
                        if (clean && b == P) { sbf[lane] = b0; sbf[lane + 32] = b1; }
                    }
                }
                BL_BAR7();
#ifdef BL_BETA_CLOCKS
                long long q2 = clock64(); t_pass += q2 - q1;
#endif
                const int i_f = ctl[0];
                if (i_f >= P) {
                    // every remaining coordinate accepted its first normal: the last block's chain is the new beta
                    if (tid < 64) sb[tid] = sbf[tid];
                    if (tid < P - i_s) sz[cur[i_s + tid]] = nbuf[mnorm + tid];
                    BL_BAR7();                                                    // ctl is still being read above
                    if (tid == 0) { ctl[1] = mnorm + (P - i_s); ctl[2] = P; }
                    BL_BAR7();
                    break;
                }
                // A miss at i_f.  The warp that found it holds beta in front of that coordinate: it forms the window and
                // takes the further tries (the same normals, in the same order, as the sequential loop would); when one
                // of them lands it commits the prefix and the coordinate itself.  Otherwise (a narrow or far window, three
                // more misses, normals exhausted) nothing is committed and warp 0, which owns the beta stream, replays
                // the accepted prefix and decides the coordinate the long way.
                if (miss_at == i_f) {
                    const int c = cur[i_f];
                    const double z1 = sz[c];
                    const double il0 = v0 ? iL[lane + (size_t)ld * c] : 0.0, il1 = v1 ? iL[lane + 32 + (size_t)ld * c] : 0.0;
                    double cmin = -INFINITY, cmax = INFINITY;
                    const double c10 = fma(-b0, il0, z1), c11 = fma(-b1, il1, z1);
                    if (il0 > 0.0 && c10 > cmin) cmin = c10;
                    else if (il0 < 0.0 && c10 < cmax) cmax = c10;
                    if (il1 > 0.0 && c11 > cmin) cmin = c11;
                    else if (il1 < 0.0 && c11 < cmax) cmax = c11;
                    cmin = warp_max_f64(cmin);
                    cmax = warp_min_f64(cmax);
                    int mn = mnorm + (i_f - i_s) + 1;                              // the first normal is spent
                    bool got = false;
                    double z2 = 0.0;
                    if (cmin < cmax && cmin < 1.0 && cmax > -1.0 && cmax - cmin >= 0.5) {
                        for (int tr = 1; tr < 4 && !got && mn < nbuf_len; ++tr) {
                            const double Z = nbuf[mn];
                            ++mn;
                            if (Z > cmin && Z < cmax) { z2 = Z; got = true; }
                        }
                    }
                    if (got) {
                        const double dz = z2 - z1;
                        sb[lane] = fma(v0 ? L[lane + (size_t)ld * c] : 0.0, dz, b0);
                        sb[lane + 32] = fma(v1 ? L[lane + 32 + (size_t)ld * c] : 0.0, dz, b1);
                        for (int i = i_s + lane; i < i_f; i += 32) sz[cur[i]] = nbuf[mnorm + (i - i_s)];
                        if (lane == 0) { sz[c] = z2; ctl[1] = mn; ctl[2] = i_f + 1; ctl[5] = 1; }
                    }
                }
                BL_BAR7();
                if (warp == 0 && ctl[5] == 0) {
                    double beta[2], z[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) beta[q] = sb[lane + 32 * q];
                    for (int i = i_s; i < i_f; ++i) {
                        const int c = cur[i];
                        const double u = su[i];
                        beta[0] = fma(lane < P ? L[lane + (size_t)ld * c] : 0.0, u, beta[0]);
                        beta[1] = fma(lane + 32 < P ? L[lane + 32 + (size_t)ld * c] : 0.0, u, beta[1]);
                    }
                    __syncwarp();
                    for (int i = i_s + lane; i < i_f; i += 32) sz[cur[i]] = nbuf[mnorm + (i - i_s)];
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 2; ++q) z[q] = sz[lane + 32 * q];
                    int mn = mnorm + (i_f - i_s), nf = 0;
                    warp_seq_range<2>(L, iL, cur, i_f, i_f + 1, beta, z, mn, src, P, ld, lane, nbuf, nbuf_len, seed, call, nf, n_rej);
#pragma unroll
                    for (int q = 0; q < 2; ++q) { sb[lane + 32 * q] = beta[q]; sz[lane + 32 * q] = z[q]; }
                    if (lane == 0) { ctl[1] = mn; ctl[2] = i_f + 1; }
                    n_fall += nf;
                }
                ++misses;
                BL_BAR7();
#ifdef BL_BETA_CLOCKS
                t_miss += clock64() - q2;
#endif
            }
        }
        __syncthreads();
        // the next sweep: warp 7's order stands if the beta stream is where it was assumed to be
        if (warp == 0 && k + 1 < P) {
#ifdef BL_BETA_CLOCKS
            long long q3 = clock64();
#endif
            const int a_true = philox_tell(src), a_assumed = ctl[4];
            if (a_true == a_assumed) {
                philox_seek(src, a_assumed + P - 1);
            } else {
                for (int i = lane; i < P; i += 32) nxt[i] = cur[i];
                __syncwarp();
                warp_sweep_permutation(nxt, src, P, lane);
#ifdef BL_BETA_CLOCKS
                ++n_redo;
#endif
            }
            if (lane == 0) { ctl[2] = 0; ctl[4] = philox_tell(src); }
#ifdef BL_BETA_CLOCKS
            t_redo += clock64() - q3;
#endif
        }
        __syncthreads();
    }
    if (tid < P) beta_out[tid] = sb[tid];
#ifdef BL_BETA_CLOCKS
    if (tid == 0 && call == 3) printf("[beta constrained spec] %d passes %lld, misses %lld, order check %lld cycles (%d redone); first normal missed %d times, %d inverse-CDF draws\n", n_pass, t_pass, t_miss, t_redo, n_redo, n_fall, n_rej);
#endif
    __syncthreads();
#undef BL_BAR7
}

// One beta draw.  Workspace (all column-major, ld = P):
//   A  [P*P]  in: PP (posterior precision, full symmetric)   -> U
//   B  [P*P]  scratch: S = PP^-1 -> L                        (constrained, mvn)
//   v  [4*P]  scratch vectors
// rhs = bP (precision-weighted mean), beta_prev (constrained only), beta_out.
template <bool kPlainOnly = false>
__device__ __forceinline__ void cta_beta_draw(int mode, double *A, double *B, double *v, const double *rhs,
                                     const double *beta_prev, double *beta_out, int P, int ld,
                                     uint64_t seed, uint32_t call, int *status, double *nbuf = nullptr, int nbuf_len = 0,
                                     double *efast = nullptr, const double *tn_pre = nullptr, int tn_flags = 0)
{
    const bool tn_fast = (tn_flags & 1) != 0;          // bit 0: fast set-up (+ speculative sweeps unless bit 1)
    __shared__ int ok;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) ok = 1;
    __syncthreads();
    double *mP = v, *z = v + P, *e = v + 2 * P;
    PhiloxSource src;
    src.open(seed, 0xFFFFFFFFFFFFFFFFull, call);

    if (mode == kBetaPlain || mode == kBetaMvn) {
        // plain: beta = PP^-1 bP + U^-1 eps = U^-1 (U^-T bP + eps), U'U = PP        (Logit.hpp:303-319)
        // mvn  : beta = PP^-1 b1 + L eps with L = chol(PP^-1) lower                 (Normal.hpp:98-131).
        //   The reference inverts PP and factorises the inverse.  With J the index reversal and
        //   J PP J = U'U (the same upper factorisation, of the reversed matrix), PP^-1 =
        //   (J U^-1 J)(J U^-1 J)' and J U^-1 J is lower triangular with a positive diagonal, i.e. it IS
        //   the Cholesky factor L (uniqueness).  So L eps = J U^-1 (J eps) and PP^-1 b1 = J U^-1 U^-T (J b1):
        //   the mvn draw is the plain draw of the reversed system with the normals reversed -- one
        //   factorisation and two substitutions instead of chol + P solves + a second chol.
        const bool rev = mode == kBetaMvn;
        if (rev) {
            double *rw = const_cast<double *>(rhs);           // the CTA's own workspace (k_beta_draw)
            const int PP2 = P * P;
            for (int k = tid; k < PP2 / 2; k += blockDim.x) {
                const int k2 = PP2 - 1 - k;
                double *a = A + (k % P) + (size_t)ld * (k / P), *b = A + (k2 % P) + (size_t)ld * (k2 / P);
                const double t = *a; *a = *b; *b = t;
            }
            for (int k = tid; k < P / 2; k += blockDim.x) { const double t = rw[k]; rw[k] = rw[P - 1 - k]; rw[P - 1 - k] = t; }
            __syncthreads();
        }
        if (efast) {
#ifdef BL_BETA_CLOCKS
            long long f0 = clock64();
#endif
            beta_fast_relayout(A, A, ld, rhs, P);
            cta_plain_fast(A, efast, beta_out, P, rev, &ok, seed, call);
            if (!ok && tid == 0) *status = 1;
#ifdef BL_BETA_CLOCKS
            if (tid == 0 && call == 3) printf("[beta clocks] fast path %lld\n", clock64() - f0);
#endif
            return;
        }
        double *rd = z;
#ifdef BL_BETA_CLOCKS
        long long c0 = clock64();
#endif
#ifdef BL_BETA_CLOCKS
        long long c1 = clock64();
#endif
        cta_ldl_upper(A, rd, P, ld, &ok, e, seed, call, rev);
#ifdef BL_BETA_CLOCKS
        long long c2 = clock64();
#endif
        if (!ok) { if (tid == 0) *status = 1; return; }
        if (tid < 32) warp_solve_plain(A, rd, rhs, e, beta_out, P, ld, lane, rev);
        __syncthreads();
#ifdef BL_BETA_CLOCKS
        if (tid == 0 && call == 3) printf("[beta clocks] normals %lld ldl %lld solve %lld\n", c1 - c0, c2 - c1, clock64() - c2);
#endif
        return;
    }

    if (kPlainOnly) return;                // this instantiation never sees the constrained draw
    const double *L = B;
    double *iL = A;
    // the scratch of the fast set-up is free again when the sweeps start: the speculative sweeps take it
    double *spec_work = (P <= 64 && blockDim.x == 256 && tn_fast && !(tn_flags & 2)) ? A + beta_tn_scratch_offset(P) : nullptr;
    if (P <= 64 && blockDim.x == 256 && tn_fast) {
        // (mP, L, 1 / L, z) from one blocked factorisation; the rejection normals arrive precomputed (tn_pre) or are
        // generated here, after the scratch they share with the inverse is free
        cta_constrained_setup_fast(A, B, mP, z, const_cast<double *>(rhs), beta_prev, A + beta_tn_scratch_offset(P), P, ld, &ok, seed, call);
        if (!ok) { if (tid == 0) *status = 1; return; }
        if (tn_pre) for (int m = tid; m < nbuf_len; m += blockDim.x) nbuf[m] = __ldcg(tn_pre + m);
        else for (int m = tid; m < nbuf_len; m += blockDim.x) nbuf[m] = stream_normal_obs(seed, kTnObs, call, m);
        __syncthreads();
    } else {
        cta_chol_upper(A, P, ld, &ok);
        if (!ok) { if (tid == 0) *status = 1; return; }

        // S = PP^-1 by solving against the identity (Logit.hpp:338-347; Normal.hpp:106-107)
        for (int k = tid; k < P * P; k += blockDim.x) B[k % P + (size_t)ld * (k / P)] = (k % P == k / P) ? 1.0 : 0.0;
        __syncthreads();
        cta_solve_utu(A, B, P, ld, P);

        // constrained coordinate-wise draw (Logit.hpp:349-399)
        cta_chol_lower(B, P, ld, &ok);
        if (!ok) { if (tid == 0) *status = 3; return; }
        if (tid < 32) {
            for (int i = lane; i < P; i += 32) mP[i] = rhs[i];
            __syncwarp();
            warp_solve_ut(A, mP, P, ld, lane);
            warp_solve_u(A, mP, P, ld, lane);
            for (int i = lane; i < P; i += 32) z[i] = beta_prev[i] - mP[i];
            __syncwarp();
            warp_solve_l(L, z, P, ld, lane);
        }
        __syncthreads();
        // U is no longer needed: its storage takes 1 / L for the sweeps -- entries on and below the diagonal of the
        // rows that carry a constraint; zero above the diagonal and in the last row (the free coefficient), so that
        // the sweeps read a column without masks and take the constraint's direction from the sign of 1 / L alone
        for (int e2 = tid; e2 < P * P; e2 += blockDim.x) {
            int j = e2 % P, c = e2 / P;
            const double l = L[j + (size_t)ld * c];
            iL[j + (size_t)ld * c] = (j >= c && j < P - 1 && l != 0.0) ? 1.0 / l : 0.0;      // L_jc = 0: no constraint from row j
        }
        // the rejection normals of the sweeps (see warp_constrained_sweeps), all threads
        for (int m = tid; m < nbuf_len; m += blockDim.x) nbuf[m] = tn_pre ? __ldcg(tn_pre + m) : stream_normal_obs(seed, kTnObs, call, m);
        __syncthreads();
    }
    if (spec_work) {
        cta_constrained_sweeps_spec(L, iL, z, beta_prev, beta_out, (int *)e, spec_work, src, P, ld, nbuf, nbuf_len, seed, call);
        return;
    }
    if (tid < 32) {
        int *is = (int *)e;                  // the permutation lives in the e[] scratch as ints
        if (P <= 64) warp_constrained_sweeps<2>(L, iL, z, beta_prev, beta_out, is, src, P, ld, lane, nbuf, nbuf_len, seed, call);
        else if (P <= 128) warp_constrained_sweeps<4>(L, iL, z, beta_prev, beta_out, is, src, P, ld, lane, nbuf, nbuf_len, seed, call);
        else warp_constrained_sweeps<8>(L, iL, z, beta_prev, beta_out, is, src, P, ld, lane, nbuf, nbuf_len, seed, call);
    }
    __syncthreads();
}

}  // namespace bl
