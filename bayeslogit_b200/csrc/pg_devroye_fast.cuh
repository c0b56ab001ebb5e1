// Fast path of the Devroye PG(1,z) sampler: fp32 DECISION PRE-FILTERS around the
// fp64 sampler of pg_samplers.cuh.
//
// Every accept/reject decision of PolyaGamma::draw_like_devroye (PolyaGamma.cpp:151-202
// and the helpers :41-55, :65-80, :82-115) is first evaluated in fp32 on the MUFU
// pipe together with a proven error margin.  Only when the fp32 value lands inside
// the margin is the reference's own fp64 expression evaluated and used, so the
// decision taken is ALWAYS the fp64 decision; the draw itself (X) is always
// computed in fp64 from the same variates in the reference's operation order.
// Net effect: identical variate consumption and identical draws as the plain
// fp64 path (tests/test_gpu_parity.py compares both against the oracle and
// against each other), at a fraction of the fp64 instruction count.
//
// Margins (derivations in DESIGN.md "fp32 decision filters"):
//   __logf  : abs err <= 2^-21.4 on [0.5,2], <= 3 ulp elsewhere
//   __expf  : <= 2 + 1.17|x| ulp
//   right-piece mass : |pr32 - pr| < 1e-5 for Z <= 8  -> band 1e-4
//   pair test E1^2 t/2 - E2 : band 2e-6 (0.32 E1^2 + E2 + 1)
//   thinning alpha=exp(-Z^2 X/2), Z < 1.5625 : band 5e-6
//   IG branch (X, mu/(mu+X), X<=t) : relative band 5e-5
//   series r1 = a1/a0 (= 3 exp(-pi^2 X) or 3 exp(-4/X)); a2/a0 < 4e-8 : band 2e-6
#pragma once

#include "pg_samplers.cuh"

namespace bl {

struct DevSetup {
    double Z, fz;     // |z|/2 and pi^2/8 + Z^2/2
    float pr32;       // fp32 estimate of the right-piece proposal mass; < 0: no estimate
    double pr64;      // fp64 value, computed on demand (NaN = not yet)
};

// fp32 estimate of dev_right_mass(Z) (PolyaGamma.cpp:65-80).
__device__ __forceinline__ float dev_right_mass_f32(float Z)
{
    const float t = 0.64f, rt = 1.25f;                 // 1/sqrt(t)
    float fz = 1.2337005501361697f + 0.5f * Z * Z;
    float b = rt * (t * Z - 1.0f);
    float a = -rt * (t * Z + 1.0f);
    float x0 = __logf(fz) + fz * t;
    float ta = -a * 0.70710678f;                       // > 0
    float lpa = __logf(0.5f * erfcxf(ta)) - ta * ta;
    float lpb;
    if (b > 0.0f) {
        lpb = log1pf(-0.5f * erfcf(b * 0.70710678f));
    } else {
        float tb = -b * 0.70710678f;
        lpb = __logf(0.5f * erfcxf(tb)) - tb * tb;
    }
    float q = 1.2732395447351628f * (__expf(x0 - Z + lpb) + __expf(x0 + Z + lpa));
    return 1.0f / (1.0f + q);
}

__device__ __forceinline__ DevSetup dev_setup(double z)
{
    DevSetup s;
    s.Z = fabs(z) * 0.5;
    s.fz = dev_fz(s.Z);
    s.pr64 = nan("");
    s.pr32 = s.Z <= 8.0 ? dev_right_mass_f32((float)s.Z) : -1.0f;
    return s;
}

static __device__ __noinline__ double dev_right_mass_slow(double Z) { return dev_right_mass(Z); }

// U_mix < mass_texpon(Z) ?
__device__ __forceinline__ bool dev_choose_right(DevSetup &s, double u)
{
    if (s.pr32 >= 0.0f) {
        double p = (double)s.pr32;
        if (u < p - 1e-4) return true;
        if (u > p + 1e-4) return false;
    }
    if (isnan(s.pr64)) s.pr64 = dev_right_mass_slow(s.Z);
    return u < s.pr64;
}

// The reference's alternating-series test in fp64 (PolyaGamma.cpp:175-198).
static __device__ __noinline__ bool dev_series_exact(double X, double u)
{
    double S = dev_coef(0, X);
    double Y = u * S;
    int n = 0;
    for (;;) {
        ++n;
        if (n & 1) {
            S = S - dev_coef(n, X);
            if (Y <= S) return true;
        } else {
            S = S + dev_coef(n, X);
            if (Y > S) return false;
        }
    }
}

__device__ __forceinline__ bool dev_series_test(double X, double u)
{
    float x = (float)X;
    // the piece is chosen on the fp64 value, exactly as PolyaGamma::a does (:45): the two
    // pieces are different series and a1/a0 differs by 6% at x = t
    float r1 = X > kTrunc ? 3.0f * __expf(-9.8696044f * x) : 3.0f * __expf(__fdividef(-4.0f, x));
    double thr = 1.0 - (double)r1;
    double band = 2e-6 + 1e-4 * (double)r1;
    if (u < thr - band) return true;
    if (u > thr + band) return false;
    return dev_series_exact(X, u);
}

// The proposal of PolyaGamma::draw_like_devroye in pieces, so that callers can either run them back to back on
// one lane (dev_propose) or regroup draws by piece first (k_devroye_regroup): variates are consumed in exactly
// the reference's order either way.
//   dev_pick          U_mix -> which piece the proposal comes from: 1 right (t + E / fz, PolyaGamma.cpp:171),
//                     2 inverse-chi^2 pairs with thinning (Z < 1/t, :87-101), 3 inverse Gaussian (:103-113)
//   dev_piece_*       the proposal X of that piece
//   dev_series_test   the alternating-series accept test on the next uniform
template <class Src>
__device__ __forceinline__ int dev_pick(Src &s, DevSetup &st)
{
    const double umix = s.unif();
    if (dev_choose_right(st, umix)) return 1;
    return 1.0 / kTrunc > st.Z ? 2 : 3;
}

template <class Src>
__device__ __forceinline__ double dev_piece_right(Src &s, const DevSetup &st)
{
    return dev_x_right(s.expon(), st.fz);
}

template <class Src>
__device__ __forceinline__ double dev_piece_pair(Src &s, const DevSetup &st)
{
    const double t = kTrunc;
    double X = 0.0;
    float hz2 = (float)(0.5 * st.Z * st.Z);
    float alpha32 = 0.0f;        // first pass: alpha = 0, the loop is always entered
    bool first = true;
    for (;;) {
        double ua = s.unif();
        if (!first) {
            bool cont;
            double a = (double)alpha32;
            if (ua > a + 5e-6) cont = true;
            else if (ua < a - 5e-6) cont = false;
            else cont = ua > exp(-0.5 * st.Z * st.Z * X);
            if (!cont) break;
        }
        first = false;
        typename Src::LazyE l1, l2;
        for (;;) {
            l1 = s.expon_lazy();
            l2 = s.expon_lazy();
            float e1 = Src::approx(l1), e2 = Src::approx(l2);
            float d = 0.32f * e1 * e1 - e2;                 // E1^2 > 2 E2 / t  <=>  d > 0
            float band = 2e-6f * (0.32f * e1 * e1 + e2 + 1.0f);
            bool again;
            if (d > band) again = true;
            else if (d < -band) again = false;
            else {
                double E1 = Src::exact(l1), E2 = Src::exact(l2);
                again = E1 * E1 > 2 * E2 / t;
            }
            if (!again) break;
        }
        X = dev_x_pair(Src::exact(l1));
        alpha32 = __expf(-hz2 * (float)X);
    }
    return X;
}

template <class Src>
__device__ __forceinline__ double dev_piece_ig(Src &s, const DevSetup &st)
{
    const double t = kTrunc;
    double X;
    double mu = 1.0 / st.Z;
    float muf = (float)mu;
    for (;;) {
        typename Src::LazyN ln = s.norm_lazy();
        double u = s.unif();
        float nf = Src::approx(ln);
        float w = muf * nf * nf;
        float x32 = __fdividef(muf, 1.0f + 0.5f * w + sqrtf(w + 0.25f * w * w));
        float p32 = __fdividef(muf, muf + x32);
        bool decided = false, accept = false, flip = false;
        if (fabs(u - (double)p32) > 5e-5 * (double)p32 + 1e-7) {
            flip = u > (double)p32;
            float xf = flip ? __fdividef(muf * muf, x32) : x32;
            if (xf > 0.64f * (1.0f + 5e-5f)) { decided = true; accept = false; }
            else if (xf < 0.64f * (1.0f - 5e-5f)) { decided = true; accept = true; }
        }
        if (decided && !accept) continue;
        // accepted or ambiguous: the reference's fp64 expression
        X = dev_x_ig(Src::exact(ln), mu);
        if (decided ? flip : (u > dev_ig_flip_threshold(X, mu))) X = dev_x_ig_flip(X, mu);
        if (decided || !(X > t)) break;
    }
    return X;
}

// One proposal + series test.  Returns true when the proposal X is accepted.
// Variates are consumed in exactly the reference's order.
template <class Src>
__device__ __forceinline__ bool dev_propose(Src &s, DevSetup &st, double &X)
{
    const int piece = dev_pick(s, st);
    if (piece == 1) X = dev_piece_right(s, st);
    else if (piece == 2) X = dev_piece_pair(s, st);
    else X = dev_piece_ig(s, st);
    return dev_series_test(X, s.unif());
}

// Sum of n PG(1,z) draws through the filtered path (PolyaGamma.cpp:126-140).
template <class Src>
__device__ double devroye_sum_fast(Src &s, int n, double z)
{
    if (n < 1) n = 1;
    DevSetup st = dev_setup(z);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        double X;
        while (!dev_propose(s, st, X)) {}
        sum += 0.25 * X;
    }
    return sum;
}

}  // namespace bl
