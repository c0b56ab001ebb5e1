// Device special functions the PG samplers need from the reference's (absent)
// RNG library: RNG::p_norm, RNG::p_gamma_rate, RNG::p_igauss, RNG::Gamma.
// Call sites in the reference: PolyaGamma.cpp:74-75; PolyaGammaAlt.cpp:56,66,73,103;
// PolyaGammaSP.cpp:218,222.  fp64 throughout.
#pragma once

#include <cmath>

namespace bl {

constexpr double kPi = 3.141592653589793238462643383279502884197;  // PolyaGamma.h:34
constexpr double kSqrt1_2 = 0.70710678118654752440;

// Phi(x)
__device__ __forceinline__ double p_norm(double x) { return 0.5 * erfc(-x * kSqrt1_2); }

// log Phi(x).  Lower tail through the scaled complementary error function so the
// far tail (x ~ -1e3, reached for |z| ~ 1e3 in mass_texpon) neither underflows
// nor loses relative accuracy.
static __device__ __noinline__ double log_p_norm(double x)
{
    if (x > 0.0) return log1p(-0.5 * erfc(x * kSqrt1_2));
    double t = -x * kSqrt1_2;
    return log(0.5 * erfcx(t)) - t * t;
}

/* log of x^a e^-x / Gamma(a).  For a >= 10 the three ~a-sized terms of the direct
 * form cancel to O(1) and lose ~log10(a) digits, so the prefix is rearranged
 * around Stirling's series: -a (mu - log1p mu) - S(a) + log sqrt(a/2pi),
 * mu = (x-a)/a, S(a) = 1/(12a) - 1/(360a^3) + ... (Temme 1979; the same
 * rearrangement Boost.Math calls regularised_gamma_prefix). */
__device__ inline double gamma_log_prefix(double a, double x)
{
    if (a < 10.0) return a * log(x) - x - lgamma(a);
    double mu = (x - a) / a;
    double phi = mu - log1p(mu);
    double ia = 1.0 / a, ia2 = ia * ia;
    double S = ia * (1.0 / 12 - ia2 * (1.0 / 360 - ia2 * (1.0 / 1260 - ia2 * (1.0 / 1680
             - ia2 * (1.0 / 1188 - ia2 * (691.0 / 360360 - ia2 * (1.0 / 156)))))));
    return -a * phi - S + 0.5 * log(a / (2.0 * 3.14159265358979323846));
}

// Regularised lower incomplete gamma P(a, x): power series for x < a+1, modified
// Lentz continued fraction for Q otherwise (returned as 1-Q).  The samplers only
// ever use 1.0 - P, so absolute accuracy is what matters.
static __device__ __noinline__ double p_gamma_lower(double a, double x)
{
    if (x <= 0.0) return 0.0;
    if (isinf(x)) return 1.0;
    double lpre = gamma_log_prefix(a, x);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 2000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return sum * exp(lpre);
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a;
    double c = 1.0 / tiny;
    double d = 1.0 / b;
    double h = d;
    for (int i = 1; i < 2000; ++i) {
        double an = -(double)i * ((double)i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return 1.0 - exp(lpre) * h;
}

// Gamma(a, x) e^x x^-a = 1 / (x+1-a - 1(1-a)/(x+3-a - 2(2-a)/(x+5-a - ...))), x >= a + 1: the same
// Legendre continued fraction as above, evaluated by the forward (Wallis) recurrence
// A_k = b_k A_{k-1} + a_k A_{k-2} (B alike), so a term costs a dozen FMAs and no division.
// Convergence is read off the determinant A_k B_{k-1} - A_{k-1} B_k = prod(-a_j), which is
// carried as a product (no cancellation): |f_k - f_{k-1}| <= 1e-16 |f_k|.
__device__ __forceinline__ double upper_gamma_cf(double a, double x)
{
    double b = x + 1.0 - a;
    double A0 = 0.0, B0 = 1.0;      // k - 1
    double A1 = 1.0, B1 = b;        // k
    double det = 1.0, di = 0.0;
    for (int i = 1; i < 2000; ++i) {
        di += 1.0;
        double an = -di * (di - a);
        b += 2.0;
        double A2 = fma(b, A1, an * A0);
        double B2 = fma(b, B1, an * B0);
        det *= -an;
        A0 = A1; B0 = B1; A1 = A2; B1 = B2;
        if (fabs(det) <= 1e-16 * fabs(A1 * B0)) break;
        if (fabs(B1) > 0x1p200) {
            A0 *= 0x1p-200; B0 *= 0x1p-200; A1 *= 0x1p-200; B1 *= 0x1p-200;
            det *= 0x1p-400;
        }
    }
    return A1 / B1;
}

// The same continued fraction in fp32 (modified Lentz, MUFU reciprocals) for the estimates of
// proposal masses that only ever meet a uniform: relative error < 1e-5 including the float
// arguments; returns a negative value if 64 terms do not converge.
__device__ __forceinline__ float upper_gamma_cf_f32(float a, float x)
{
    float bb = x + 1.0f - a;
    float c = 1e30f, d = 1.0f / bb, h = d;
    for (int i = 1; i <= 64; ++i) {
        float an = -(float)i * ((float)i - a);
        bb += 2.0f;
        d = an * d + bb;
        d = fabsf(d) < 1e-30f ? 1e-30f : d;
        c = bb + an / c;
        c = fabsf(c) < 1e-30f ? 1e-30f : c;
        d = 1.0f / d;
        float del = d * c;
        h *= del;
        if (fabsf(del - 1.0f) < 3e-7f) return h;
    }
    return -1.0f;
}

// Inverse-Gaussian CDF with the two normal tails taken directly from erfc / erfcx instead of
// through log Phi: Phi(b) + exp(2 lambda / mu) Phi(a), a < 0 (same quantity as p_igauss below;
// used where only a proposal weight depends on it).
__device__ __forceinline__ double p_igauss_direct(double x, double mu, double lambda)
{
    double Z = 1.0 / mu;
    double s = sqrt(lambda / x);
    double b = s * (x * Z - 1.0);
    double t = s * (x * Z + 1.0) * kSqrt1_2;
    return 0.5 * erfc(-b * kSqrt1_2) + 0.5 * erfcx(t) * exp(2.0 * lambda * Z - t * t);
}

// RNG::p_gamma_rate(x, shape, rate) = P(shape, x * rate)
__device__ __forceinline__ double p_gamma_rate(double x, double shape, double rate)
{
    return p_gamma_lower(shape, x * rate);
}

// Inverse-Gaussian CDF in log space (Code/R/PG.R:15-23).
__device__ __forceinline__ double p_igauss(double x, double mu, double lambda)
{
    double Z = 1.0 / mu;
    double s = sqrt(lambda / x);
    double b = s * (x * Z - 1.0);
    double a = -1.0 * s * (x * Z + 1.0);
    return exp(log_p_norm(b)) + exp(2.0 * lambda * Z + log_p_norm(a));
}

}  // namespace bl
