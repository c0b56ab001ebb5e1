// Alternate Polya-Gamma sampler PG(h, z), h >= 1 (fp64): J*(h', |z|/2)/4 summed over chunks h' <= 4.
//
// Reference statements this file has to agree with:
//   alt_plan      chunking of h                 PolyaGammaAlt.cpp:205-225
//   alt_setup     per-chunk constants           PolyaGammaAlt.cpp:114-137 (+ :51-75)
//   alt_trip      propose / series test         PolyaGammaAlt.cpp:139-202 (+ :6-49, :77-108)
//
// The reference recomputes the chunk constants (truncation point, piece masses, coefficient
// scale) for every chunk although a draw only ever uses two shapes: 4 for the full chunks and
// the remainder (or half of it).  Here they are computed once per draw and shape (alt_setup) and
// the rejection loop is a per-lane state machine (alt_trip: one attempt at a proposal, and the
// alternating-series test once a proposal exists), so that the binned rpg_hybrid path can run
// set-up and loop as two kernels and keep all 32 lanes of a warp busy (pg_hybrid.cu).  The
// per-lane kernels and the tape path run the same two functions back to back (alt_draw).
//
// Only X = 4 omega leaves the loop.  Everything else -- masses, envelope g~(x), coefficients
// a_n(x) -- enters accept/reject comparisons only; there log(x^3) is shared between envelope and
// coefficients (the reference takes log(x), log(x^3) and log(2 pi x^3) separately; the shared
// form differs by at most 2 ulp of the logarithm).
#pragma once

namespace bl {

template <class Src>
__device__ __forceinline__ double alt_rtinvchi2(Src &s, double h, double trunc)
{
    double h2 = h * h;
    double R = trunc / h2;
    double E1 = s.expon();
    double E2 = s.expon();
    while ((E1 * E1) > (2 * E2 / R)) {
        E1 = s.expon();
        E2 = s.expon();
    }
    double X = 1 + E1 * R;
    X = R / (X * X);
    return h2 * X;
}

// PolyaGammaAlt.cpp:51-58, the naive exp(2 lambda z) Phi(a) form as written
__device__ __forceinline__ double alt_pigauss(double x, double z, double lambda)
{
    double sq = sqrt(lambda / x);
    double b = sq * (x * z - 1);
    double a = sq * (x * z + 1) * -1.0;
    return p_norm(b) + ool::exp_(2 * lambda * z) * p_norm(a);
}

enum AltField {
    kAltH = 0,     // chunk shape h' in [1, 4]
    kAltTrunc,     // truncation point t(h'), PolyaGammaAlt.h:14-44
    kAltPr,        // mass of the right (gamma) piece, :137
    kAltCoef,      // 2^h' / sqrt(2 pi), :139
    kAltLgh,       // lgamma(h'), for g~ right of t (:101)
    kAltLd0, kAltLd1,            // log(h'), log(h' + 2): log d_n of the first two coefficients (:45)
    kAltLtB, kAltLtC0, kAltLtLM, // ltgamma(h', rate_z, t): b = rate * trunc, c0, log M (Ch.R:96-101)
    kAltPrBand,    // 0: kAltPr is the fp64 mass; > 0: an fp32 estimate within this distance of it
    kAltSetupDoubles
};
constexpr int kAltStateDoubles = 2 * kAltSetupDoubles;   // [0]: shape 4, [1]: remainder shape

struct AltState {
    double f[kAltStateDoubles];
    __device__ __forceinline__ double get(int k) const { return f[k]; }
};

// strided view: field k at o[k * stride] (shared-memory copy of the lane's draw, see SpStateRef)
struct AltStateRef {
    const double *o;
    size_t stride;
    __device__ __forceinline__ double get(int k) const { return o[(size_t)k * stride]; }
};

// PolyaGammaAlt.cpp:205-225: nfull chunks of shape 4, then nrem (1 or 2) chunks of shape hrem
__device__ __forceinline__ void alt_plan(double h, int &nfull, int &nrem, double &hrem)
{
    double n = floor((h - 1.0) / 4.0);
    double remain = h - 4.0 * n;
    nfull = (int)n;
    if (remain > 4.0) {
        nrem = 2;
        hrem = 0.5 * remain;
    } else {
        nrem = 1;
        hrem = remain;
    }
}

// Mass of the right (gamma) piece of a chunk of shape h at tilt z = |z|/2, fp64, as
// PolyaGammaAlt.cpp:124-137 computes it (w_left :60-68 with the naive pigauss, w_right :70-75).
static __device__ __noinline__ double alt_pr_fp64(double h, double z, double trunc)
{
    const double kLog2 = 0.693147180559945309417232;
    double wl, wr;
    if (z != 0)
        wl = ool::exp_(h * (kLog2 - z)) * alt_pigauss(trunc, z / h, h * h);
    else
        wl = ool::exp_(h * kLog2) * (1.0 - p_gamma_rate(1 / trunc, 0.5, 0.5 * h * h));
    double lambda_z = kPi * kPi * 0.125 + 0.5 * z * z;
    wr = ool::exp_(h * ool::log_((0.5 * kPi) / lambda_z)) * (1.0 - p_gamma_rate(trunc, h, lambda_z));
    return wr / (wr + wl);
}

// fp32 estimate of the same mass for the binned path's set-up kernel: it only ever meets a
// uniform, so the loop compares against estimate -+ kAltPrBand and evaluates alt_pr_fp64 only
// inside the band (same scheme as the saddle-point sampler's pl).
//   wl = 2^h [e^(-hz) Phi(b) + e^(hz) Phi(a)], b = s (t z / h - 1), a = -s (t z / h + 1), s = h / sqrt(t)
//        with the second term as erfcx(-a / sqrt 2) e^(hz - a^2/2) / 2 (no overflow);
//   wr = (pi/2 / lambda)^h Q(h, t lambda), Q by the power series of P or the continued fraction.
// No estimate (negative) when z is large; measured
// |estimate - mass| < 2e-6 elsewhere (test_alternate_pr_estimate), band 2e-5.
constexpr double kAltPrBandWidth = 2e-5;

__device__ __forceinline__ float alt_pr_estimate(double h, double z, double trunc)
{
    const float hf = (float)h, zf = (float)z, tf = (float)trunc;
    if (!(zf < 12.0f)) return -1.0f;
    const float s = hf * rsqrtf(tf);
    const float tz = tf * zf / hf;
    const float b = s * (tz - 1.0f), na = s * (tz + 1.0f);             // na = -a > 0
    const float phib = 0.5f * erfcf(-b * 0.70710678f);
    const float wl = exp2f(hf) * (__expf(-hf * zf) * phib
                                  + 0.5f * erfcxf(na * 0.70710678f) * __expf(hf * zf - 0.5f * na * na));
    const float lam = 1.2337005501361697f + 0.5f * zf * zf;
    const float x = tf * lam;
    // Q(h, x) = 1 - P(h, x): left of the mode region by the power series of P,
    //           P = x^h e^-x / Gamma(h + 1) * sum_k x^k / ((h+1)..(h+k)); right of it by the continued
    //           fraction of Gamma(h, x) directly (1 - P would cancel)
    const float lpre = hf * __logf(x) - x;
    float Q;
    if (x < hf + 1.0f) {
        float term = 1.0f, sum = 1.0f, ap = hf;
        for (int k = 0; k < 60; ++k) {
            ap += 1.0f;
            term *= x / ap;
            sum += term;
            if (term < 1e-8f * sum) break;
        }
        Q = 1.0f - sum * __expf(lpre - lgammaf(hf + 1.0f));
    } else {
        float cf = upper_gamma_cf_f32(hf, x);
        if (!(cf > 0.0f)) return -1.0f;
        Q = cf * __expf(lpre - lgammaf(hf));
    }
    if (!(Q > 0.0f)) return -1.0f;
    const float wr = __expf(hf * __logf(1.5707963267948966f / lam)) * Q;
    const float pr = wr / (wr + wl);
    return pr >= 0.0f && pr <= 1.0f ? pr : -1.0f;
}

// z is |z|/2; writes kAltSetupDoubles values
template <bool kEstimatePr>
static __device__ __noinline__ void alt_setup(double h, double z, double *out)
{
    const double kLog2 = 0.693147180559945309417232;
    int idx = (int)floor((h - 1.0) * 100.0);
    double trunc = PG_TRUNC_SCHEDULE[idx];
    double rate_z = 0.125 * kPi * kPi + 0.5 * z * z;
    float est = kEstimatePr ? alt_pr_estimate(h, z, trunc) : -1.0f;
    out[kAltH] = h;
    out[kAltTrunc] = trunc;
    out[kAltPr] = est >= 0.0f ? (double)est : alt_pr_fp64(h, z, trunc);
    out[kAltPrBand] = est >= 0.0f ? kAltPrBandWidth : 0.0;
    out[kAltCoef] = ool::exp_(h * kLog2 - 0.5 * 1.8378770664093454835606594728112 /* log(2 pi) */);
    out[kAltLgh] = ool::lgamma_(h);
    out[kAltLd0] = ool::log_(h);
    out[kAltLd1] = ool::log_(2.0 + h);
    double b = rate_z * trunc;
    double d1 = b - h;
    double d3 = h - 1.0;
    double c0 = 0.5 * (d1 + sqrt(d1 * d1 + 4.0 * b)) / b;
    out[kAltLtB] = b;
    out[kAltLtC0] = c0;
    out[kAltLtLM] = d3 * ool::log_(d3 / (1.0 - c0)) - d3;   // unused when h' == 1
}

// U < mass of the right piece ?  (PolyaGammaAlt.cpp:143) against the state's mass, exact or estimated
template <class St>
__device__ __forceinline__ bool alt_pick_right(double u, const St &st, int o, double z)
{
    const double pr = st.get(o + kAltPr), band = st.get(o + kAltPrBand);
    if (u < pr - band) return true;
    if (u > pr + band || band == 0.0) return false;
    return u < alt_pr_fp64(st.get(o + kAltH), z, st.get(o + kAltTrunc));
}

struct AltLane {
    double X, alpha, sum;
    int trial;     // proposals made for the current chunk (cap 10000, :141)
    int phase;     // 0 pick a piece, 1 right piece, 2 left piece (inverse chi^2), 3 left piece (IG), 4 proposal ready
    int nfull;     // chunks of shape 4 still to draw
    int nrem;      // remainder chunks still to draw
    __device__ __forceinline__ void start(int nf, int nr)
    {
        X = 0.0; alpha = 0.0; sum = 0.0; trial = 0; phase = 0; nfull = nf; nrem = nr;
    }
};

// a_n(x) of PolyaGammaAlt.cpp:37-49 (a_coef_recursive) with log(x^3) and log(d_n) supplied
__device__ __forceinline__ double alt_coef(double n, double lx3, double ldn, double x, double h,
                                           double coef_h, double &g)
{
    double d_n = 2.0 * n + h;
    if (n != 0)
        g *= (n + h - 1) / n;
    else
        g = 1.0;
    double coef = coef_h * g;
    double log_kernel = -0.5 * (lx3 + d_n * d_n / x) + ldn;
    return coef * ool::exp_(log_kernel);
}

// The piece the next proposal comes from (PolyaGammaAlt.cpp:141-147): phase 0 -> 1, 2 or 3.  When
// the chunk's 10000 proposals are used up it closes with -1 instead (:202).  Returns true when
// that closed the whole draw.  z is |z|/2.
template <class Src, class St>
__device__ __forceinline__ bool alt_pick(Src &src, AltLane &L, double z, const St &st)
{
    const int o = L.nfull > 0 ? 0 : kAltSetupDoubles;
    if (L.trial >= 10000) {
        L.sum += -1.0;
        L.trial = 0;
        if (L.nfull > 0)
            L.nfull--;
        else
            L.nrem--;
        return L.nfull == 0 && L.nrem == 0;
    }
    L.trial++;
    if (alt_pick_right(src.unif(), st, o, z)) {
        L.phase = 1;
    } else {
        L.phase = (st.get(o + kAltH) / z > st.get(o + kAltTrunc)) ? 2 : 3;
        L.alpha = 0.0;
    }
    return false;
}

// The alternating-series test of a proposal X against Y = U g~(X), PolyaGammaAlt.cpp:160-200, in fp64
// as written (a_n by the recursive form :37-49, monotonicity required, 200-term cap).
template <class St>
static __device__ __noinline__ bool alt_series_exact(double X, double u, double h, double trunc, const St &st, int o)
{
    const int max_inner = 200;
    const double coef_h = st.get(o + kAltCoef);
    const double lx3 = ool::log_(X * X * X);
    double g = 1.0;
    double S = alt_coef(0.0, lx3, st.get(o + kAltLd0), X, h, coef_h, g);
    double a_n = S;
    double gt;                                                          // g~(x), :99-108
    if (X > trunc)
        gt = ool::exp_(h * 0.45158270528945486472619522989488 /* log(pi/2) */ + (h - 1) * (lx3 * (1.0 / 3.0))
                       - kPi * kPi * 0.125 * X - st.get(o + kAltLgh));
    else
        gt = h * ool::exp_(h * 0.693147180559945309417232 - 0.5 * (1.8378770664093454835606594728112 + lx3)
                           - 0.5 * h * h / X);
    double Y = u * gt;
    int n = 0;
    bool go = true, accept = false;
    while (go && n < max_inner) {
        ++n;
        double prev = a_n;
        double ldn = n == 1 ? st.get(o + kAltLd1) : ool::log_(2.0 * n + h);
        a_n = alt_coef((double)n, lx3, ldn, X, h, coef_h, g);
        bool decreasing = a_n <= prev;
        if (n & 1) {
            S = S - a_n;
            if (Y <= S && decreasing) {
                accept = true;
                go = false;
            }
        } else {
            S = S + a_n;
            if (Y > S && decreasing) go = false;
        }
    }
    return accept;
}

// fp32 pre-filter of that test.  Dividing by g~ turns it into U against partial sums of
// r_n = a_n / g~ = coef g_n d_n exp(-(3/2) log x - d_n^2 / 2x - log g~): the linear parts are fp64,
// log x and the three exponentials fp32 (MUFU).  Nearly every proposal leaves at the first odd
// term (accept) or the first even term (reject); those two exits are taken when the comparison
// and the monotonicity condition both clear a band of >= 3x the fp32 error bound
// (r_n carries |e_n| 2^-24 + 2 ulp of __expf + 2 |dlog x|, < 3e-6 for any r_n > 1e-8).  Everything
// else -- in-band, later terms, non-monotone coefficients, NaN -- goes to alt_series_exact, so the
// decision taken is always the fp64 decision.
template <class St>
__device__ __forceinline__ bool alt_series_test(double X, double u, double h, double trunc, const St &st, int o)
{
    const double coef_h = st.get(o + kAltCoef);
    const double iX = 1.0 / X;
    const double lx3 = 3.0 * (double)__logf((float)X);
    double lgt;
    if (X > trunc)
        lgt = h * 0.45158270528945486472619522989488 + (h - 1) * (lx3 * (1.0 / 3.0)) - kPi * kPi * 0.125 * X
            - st.get(o + kAltLgh);
    else
        lgt = st.get(o + kAltLd0) + h * 0.693147180559945309417232 - 0.5 * (1.8378770664093454835606594728112 + lx3)
            - 0.5 * h * h * iX;
    const double base = -0.5 * lx3 - lgt;
    const double d0 = h, d1 = h + 2.0, d2 = h + 4.0;
    const double r0 = coef_h * d0 * (double)__expf((float)(base - 0.5 * d0 * d0 * iX));
    const double r1 = coef_h * h * d1 * (double)__expf((float)(base - 0.5 * d1 * d1 * iX));
    const double r2 = coef_h * (0.5 * h * (h + 1.0)) * d2 * (double)__expf((float)(base - 0.5 * d2 * d2 * iX));
    const double band = 1e-5 * (r0 + r1 + r2) + 1e-9;
    const double s1 = r0 - r1;
    if (u <= s1 - band && r1 <= r0 - band) return true;                     // accepted at n = 1
    const bool not_at_1 = u > s1 + band || r1 > r0 + band;
    if (not_at_1 && u > s1 + r2 + band && r2 <= r1 - band) return false;    // rejected at n = 2
    return alt_series_exact(X, u, h, trunc, st, o);
}

// One trip of a lane: returns true when the whole draw is complete (L.sum = omega).  z is |z|/2.
// kPickInside = false: the caller guarantees a piece has been picked (L.phase != 0) -- the
// regrouping kernel picks when the previous trip ends, and should not carry a second copy.
template <bool kPickInside = true, class Src, class St>
__device__ __forceinline__ bool alt_trip(Src &src, AltLane &L, double z, const St &st)
{
    if (kPickInside && L.phase == 0) {
        if (alt_pick(src, L, z, st)) return true;
        if (L.phase == 0) return false;                                // chunk closed at its proposal cap
    }
    const int o = L.nfull > 0 ? 0 : kAltSetupDoubles;
    const double h = st.get(o + kAltH);
    const double trunc = st.get(o + kAltTrunc);
    bool chunk_done = false;
    double chunk_val = 0.0;
    if (L.phase == 1) {
        double rate_z = 0.125 * kPi * kPi + 0.5 * z * z;
        if (h == 1.0) {
            L.X = src.expon() / rate_z + trunc;                         // Ch.R:92
            L.phase = 4;
        } else {
            double b = st.get(o + kAltLtB), c0 = st.get(o + kAltLtC0);
            double x = b + src.expon() / c0;
            double u = src.unif();
            double l_rho = (h - 1.0) * ool::log_(x) - x * (1.0 - c0);
            if (ool::log_(u) <= l_rho - st.get(o + kAltLtLM)) {
                L.X = trunc * (x / b);
                L.phase = 4;
            }
        }
    } else if (L.phase == 2) {
        if (src.unif() > L.alpha) {
            L.X = alt_rtinvchi2(src, h, trunc);
            L.alpha = ool::exp_(-0.5 * z * z * L.X);
        } else {
            L.phase = 4;
        }
    } else if (L.phase == 3) {
        double X = igauss(src, h / z, h * h);
        if (!(X > trunc)) {
            L.X = X;
            L.phase = 4;
        }
    }
    if (L.phase == 4) {
        const double X = L.X;
        const bool accept = alt_series_test(X, src.unif(), h, trunc, st, o);
        L.phase = 0;
        if (accept) {
            chunk_done = true;
            chunk_val = 0.25 * X;
        }
    }
    if (chunk_done) {
        L.sum += chunk_val;
        L.trial = 0;
        if (L.nfull > 0)
            L.nfull--;
        else
            L.nrem--;
        if (L.nfull == 0 && L.nrem == 0) return true;
    }
    return false;
}

// PG(h, z) on one lane: set-up for the (at most two) chunk shapes, then trips until complete.
template <class Src>
__device__ double alt_draw(Src &s, double h, double z)
{
    if (h < 1) return 0;                         // PolyaGammaAlt.cpp:207-210
    if (h != h) return h;                        // NaN shape: no table row to read
    int nfull, nrem;
    double hrem;
    alt_plan(h, nfull, nrem, hrem);
    double zh = fabs(z) * 0.5;
    AltState st;
    if (nfull > 0) alt_setup<false>(4.0, zh, st.f);
    alt_setup<false>(hrem, zh, st.f + kAltSetupDoubles);
    AltLane L;
    L.start(nfull, nrem);
    while (!alt_trip(s, L, zh, st)) {}
    return L.sum;
}

}  // namespace bl
