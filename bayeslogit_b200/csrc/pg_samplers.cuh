// Per-lane Polya-Gamma samplers (fp64), templated on the variate source so the
// very same code runs from the Philox stream and from an injected tape.
//
// What each routine computes, and the reference statement it has to agree with:
//   devroye_one      PG(1,z)            PolyaGamma.cpp:151-202 (+ :41-55, :65-80, :82-115)
//   devroye_sum      sum of n PG(1,z)   PolyaGamma.cpp:126-140
//   gamma_sum        truncated sum      PolyaGamma.cpp:142-149, :19-39
//   alt_draw         PG(h,z), h>=1      PolyaGammaAlt.cpp:114-203, :205-225   (pg_alt.cuh)
//   v_eval           y -> v             InvertY.cpp:57-99                      (pg_sp.cuh)
//   sp_draw          PG(n,z), n large   PolyaGammaSP.cpp:169-264               (pg_sp.cuh)
//   pg_m1 / pg_m2    exact moments      PolyaGamma.cpp:208-239
//   hybrid           regime dispatch    LogitWrapper.cpp:140-162
// Composite variates (the reference takes them from its absent RNG library) follow
// the in-tree R statements: Ch.R:83-114 (left-truncated gamma), Ch.R:403-413
// (inverse Gaussian), SPSample.R:534-550 (right-truncated inverse chi^2 through a
// one-sided truncated normal, Robert 1995).
//
// Differences from the reference that do NOT change any decision or value:
//   * the right-piece proposal mass depends on Z only and is computed once per
//     draw instead of once per proposal (PolyaGamma.cpp:170 recomputes it);
//   * constant logarithms are folded.
#pragma once

#include "philox.cuh"
#include "specfun.cuh"

#define PG_TABLE_QUAL static __device__ __constant__
#include "pg_tables.h"

namespace bl {

constexpr double kTrunc = 0.64;  // PolyaGamma.h:37

// Out-of-line libm entry points for the big samplers (alternate, saddle point).  Inlined,
// every log/exp/tan/... call site carries its own 40-200 instruction body and the saddle-point
// kernel grows to ~140 KB of SASS: its profile (profiles/r1_05_*) showed 15 of 20 stall cycles
// per instruction waiting on instruction fetch.  One shared body per function keeps the kernels
// near the instruction cache; results are unchanged (same libdevice code).
namespace ool {
static __device__ __noinline__ double log_(double x) { return ::log(x); }
static __device__ __noinline__ double exp_(double x) { return ::exp(x); }
static __device__ __noinline__ double tan_(double x) { return ::tan(x); }
static __device__ __noinline__ double tanh_(double x) { return ::tanh(x); }
static __device__ __noinline__ double cos_(double x) { return ::cos(x); }
static __device__ __noinline__ double cosh_(double x) { return ::cosh(x); }
static __device__ __noinline__ double atan_(double x) { return ::atan(x); }
static __device__ __noinline__ double lgamma_(double x) { return ::lgamma(x); }
static __device__ __noinline__ double tgamma_(double x) { return ::tgamma(x); }
}  // namespace ool

// ----------------------------------------------------------------------------
// Devroye PG(1,z)
// ----------------------------------------------------------------------------

__device__ __forceinline__ double dev_coef(int n, double x)
{
    double K = (n + 0.5) * kPi;
    if (x > kTrunc) return K * exp(-0.5 * K * K * x);
    if (x > 0) {
        double e = -1.5 * (log(0.5 * kPi) + log(x)) + log(K) - 2.0 * (n + 0.5) * (n + 0.5) / x;
        return exp(e);
    }
    return 0.0;
}

__device__ __forceinline__ double dev_right_mass(double Z)
{
    const double t = kTrunc;
    double fz = 0.125 * kPi * kPi + 0.5 * Z * Z;
    double b = sqrt(1.0 / t) * (t * Z - 1);
    double a = sqrt(1.0 / t) * (t * Z + 1) * -1.0;
    double x0 = log(fz) + fz * t;
    double xb = x0 - Z + log_p_norm(b);
    double xa = x0 + Z + log_p_norm(a);
    double qdivp = 4 / kPi * (exp(xb) + exp(xa));
    return 1.0 / (1.0 + qdivp);
}

// The three proposal formulas, written with explicitly rounded fp64 operations
// (no FMA contraction) so that, given the same variates, X carries the same bits
// as the reference's x86-64 arithmetic and as every other code path of this engine.
//   right piece        X = t + E/fz                         PolyaGamma.cpp:171
//   inverse chi^2 pair X = t / (1 + E1 t)^2                 PolyaGamma.cpp:98-99
//   inverse Gaussian   X = mu + mu/2 muY - mu/2 sqrt(4 muY + muY^2), maybe mu^2/X   :107-111
// fz = pi^2/8 + Z^2/2 (PolyaGamma.cpp:157), unfused like the reference's x86-64 build
__device__ __forceinline__ double dev_fz(double Z)
{
    return __dadd_rn(0.125 * kPi * kPi, __dmul_rn(__dmul_rn(0.5, Z), Z));
}

__device__ __forceinline__ double dev_x_right(double E, double fz)
{
    return __dadd_rn(kTrunc, __ddiv_rn(E, fz));
}

__device__ __forceinline__ double dev_x_pair(double E1)
{
    double X = __dadd_rn(1.0, __dmul_rn(E1, kTrunc));
    return __ddiv_rn(kTrunc, __dmul_rn(X, X));
}

__device__ __forceinline__ double dev_x_ig(double N, double mu)
{
    double Y = __dmul_rn(N, N);
    double half_mu = __dmul_rn(0.5, mu);
    double mu_Y = __dmul_rn(mu, Y);
    double rad = __dsqrt_rn(__dadd_rn(__dmul_rn(4.0, mu_Y), __dmul_rn(mu_Y, mu_Y)));
    return __dadd_rn(__dadd_rn(mu, __dmul_rn(half_mu, mu_Y)), -__dmul_rn(half_mu, rad));
}

__device__ __forceinline__ double dev_ig_flip_threshold(double X, double mu)
{
    return __ddiv_rn(mu, __dadd_rn(mu, X));
}

__device__ __forceinline__ double dev_x_ig_flip(double X, double mu)
{
    return __ddiv_rn(__dmul_rn(mu, mu), X);
}

template <class Src>
__device__ __forceinline__ double dev_trunc_igauss(Src &s, double Z)
{
    const double t = kTrunc;
    double X = t + 1.0;
    if (1.0 / kTrunc > Z) {
        double alpha = 0.0;
        while (s.unif() > alpha) {
            double E1 = s.expon();
            double E2 = s.expon();
            while (E1 * E1 > 2 * E2 / t) {
                E1 = s.expon();
                E2 = s.expon();
            }
            X = dev_x_pair(E1);
            alpha = exp(-0.5 * Z * Z * X);
        }
    } else {
        double mu = 1.0 / Z;
        while (X > t) {
            X = dev_x_ig(s.norm(), mu);
            if (s.unif() > dev_ig_flip_threshold(X, mu)) X = dev_x_ig_flip(X, mu);
        }
    }
    return X;
}

template <class Src>
__device__ __forceinline__ double devroye_one(Src &s, double Z, double fz, double right_mass)
{
    for (;;) {
        double X;
        if (s.unif() < right_mass)
            X = dev_x_right(s.expon(), fz);
        else
            X = dev_trunc_igauss(s, Z);
        double S = dev_coef(0, X);
        double Y = s.unif() * S;
        int n = 0;
        for (;;) {
            ++n;
            if (n & 1) {
                S = S - dev_coef(n, X);
                if (Y <= S) return 0.25 * X;
            } else {
                S = S + dev_coef(n, X);
                if (Y > S) break;
            }
        }
    }
}

template <class Src>
__device__ __forceinline__ double devroye_sum(Src &s, int n, double z)
{
    if (n < 1) n = 1;  // the package builds with -DNTHROW: clamp, PolyaGamma.cpp:128-135
    double Z = fabs(z) * 0.5;
    double fz = dev_fz(Z);
    double pr = dev_right_mass(Z);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += devroye_one(s, Z, fz, pr);
    return sum;
}

// ----------------------------------------------------------------------------
// Truncated sum of gammas
// ----------------------------------------------------------------------------

template <class Src>
__device__ __forceinline__ double gamma_sum(Src &s, double b, double z, int T)
{
    if (T < 1) T = 1;
    double x = 0.0;
    double kappa = z * z;
    for (int k = 0; k < T; ++k) {
        double d = (double)k + 0.5;
        double bk = (4 * kPi * kPi) * d * d;
        x += s.gamma(b) / (bk + kappa);
    }
    return 2.0 * x;
}

// ----------------------------------------------------------------------------
// Exact moments
// ----------------------------------------------------------------------------

__device__ __forceinline__ double jj_m1(double b, double z)
{
    z = fabs(z);
    if (z > 1e-12) return b * tanh(z) / z;
    return b * (1 - (1.0 / 3) * pow(z, 2) + (2.0 / 15) * pow(z, 4) - (17.0 / 315) * pow(z, 6));
}

__device__ __forceinline__ double jj_m2(double b, double z)
{
    z = fabs(z);
    if (z > 1e-12) {
        double tz = tanh(z) / z;
        return (b + 1) * b * (tz * tz) + b * ((tanh(z) - z) / (z * z * z));
    }
    double p = 1 - (1.0 / 3) * pow(z, 2) + (2.0 / 15) * pow(z, 4) - (17.0 / 315) * pow(z, 6);
    return (b + 1) * b * (p * p) + b * ((-1.0 / 3) + (2.0 / 15) * pow(z, 2) - (17.0 / 315) * pow(z, 4));
}

__device__ __forceinline__ double pg_m1(double b, double z) { return jj_m1(b, 0.5 * z) * 0.25; }
__device__ __forceinline__ double pg_m2(double b, double z) { return jj_m2(b, 0.5 * z) * 0.0625; }

// ----------------------------------------------------------------------------
// Composite variates
// ----------------------------------------------------------------------------

template <class Src>
__device__ __forceinline__ double igauss(Src &s, double mu, double lambda)
{
    double nu = s.norm();
    double y = nu * nu;
    double x = mu + 0.5 * mu * mu * y / lambda
             - 0.5 * mu / lambda * sqrt(4.0 * mu * lambda * y + (mu * y) * (mu * y));
    if (s.unif() > mu / (mu + x)) x = mu * mu / x;
    return x;
}

template <class Src>
__device__ __noinline__ double ltgamma(Src &s, double shape, double rate, double trunc)
{
    double a = shape;
    double b = rate * trunc;
    if (trunc <= 0.0 || shape < 1.0) return 0.0;
    if (shape == 1.0) return s.expon() / rate + trunc;
    double d1 = b - a;
    double d3 = a - 1.0;
    double c0 = 0.5 * (d1 + sqrt(d1 * d1 + 4.0 * b)) / b;
    double l_M = d3 * ool::log_(d3 / (1.0 - c0)) - d3;
    double x;
    for (;;) {
        x = b + s.expon() / c0;
        double u = s.unif();
        double l_rho = d3 * ool::log_(x) - x * (1.0 - c0);
        if (ool::log_(u) <= l_rho - l_M) break;
    }
    return trunc * (x / b);
}

template <class Src>
__device__ __forceinline__ double tnorm_left(Src &s, double left)
{
    if (left < 0.0) {
        for (;;) {
            double z = s.norm();
            if (z > left) return z;
        }
    }
    double astar = 0.5 * (left + sqrt(left * left + 4.0));
    for (;;) {
        double z = s.expon() / astar + left;
        double rho = ool::exp_(-0.5 * (z - astar) * (z - astar));
        if (s.unif() < rho) return z;
    }
}

template <class Src>
__device__ __forceinline__ double rtinvchi2(Src &s, double scale, double trunc)
{
    double R = trunc / scale;
    double z = tnorm_left(s, 1.0 / sqrt(R));
    return scale / (z * z);
}

}  // namespace bl

#include "pg_alt.cuh"  // alternate sampler, 1 <= shape, in chunks of at most 4
#include "pg_sp.cuh"   // y(v) inversion and the saddle-point sampler

namespace bl {

// ----------------------------------------------------------------------------
// Regime dispatch
// ----------------------------------------------------------------------------

enum Regime { kRegZero = 0, kRegGamma = 1, kRegDevroye = 2, kRegAlt = 3, kRegSP = 4, kRegNormal = 5 };

__device__ __forceinline__ int regime_of(double b)
{
    if (b > 170) return kRegNormal;
    if (b > 13) return kRegSP;
    if (b == 1 || b == 2) return kRegDevroye;
    if (b > 1) return kRegAlt;
    if (b > 0) return kRegGamma;
    return kRegZero;
}

template <class Src>
__device__ double devroye_sum_fast(Src &s, int n, double z);   // pg_devroye_fast.cuh

template <class Src>
__device__ double hybrid_draw(Src &s, double b, double z, int &aux)
{
    aux = 0;
    switch (regime_of(b)) {
    case kRegNormal: {
        double m = pg_m1(b, z);
        double v = pg_m2(b, z) - m * m;
        return m + sqrt(v) * s.norm();
    }
    case kRegSP: {
        double d;
        aux = sp_draw(s, d, b, z);
        return d;
    }
    case kRegDevroye:
        return devroye_sum_fast(s, (int)b, z);
    case kRegAlt:
        return alt_draw(s, b, z);
    case kRegGamma:
        return gamma_sum(s, b, z, 200);
    default:
        return 0.0;
    }
}

}  // namespace bl
