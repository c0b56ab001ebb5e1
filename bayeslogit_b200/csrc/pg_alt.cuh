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

// z is |z|/2; writes kAltSetupDoubles values
static __device__ __noinline__ void alt_setup(double h, double z, double *out)
{
    const double kLog2 = 0.693147180559945309417232;
    int idx = (int)floor((h - 1.0) * 100.0);
    double trunc = PG_TRUNC_SCHEDULE[idx];
    double rate_z = 0.125 * kPi * kPi + 0.5 * z * z;
    double wl, wr;
    if (z != 0)
        wl = ool::exp_(h * (kLog2 - z)) * alt_pigauss(trunc, z / h, h * h);
    else
        wl = ool::exp_(h * kLog2) * (1.0 - p_gamma_rate(1 / trunc, 0.5, 0.5 * h * h));
    {
        double lambda_z = kPi * kPi * 0.125 + 0.5 * z * z;
        wr = ool::exp_(h * ool::log_((0.5 * kPi) / lambda_z)) * (1.0 - p_gamma_rate(trunc, h, lambda_z));
    }
    out[kAltH] = h;
    out[kAltTrunc] = trunc;
    out[kAltPr] = wr / (wr + wl);
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
    if (src.unif() < st.get(o + kAltPr)) {
        L.phase = 1;
    } else {
        L.phase = (st.get(o + kAltH) / z > st.get(o + kAltTrunc)) ? 2 : 3;
        L.alpha = 0.0;
    }
    return false;
}

// One trip of a lane: returns true when the whole draw is complete (L.sum = omega).  z is |z|/2.
template <class Src, class St>
__device__ __forceinline__ bool alt_trip(Src &src, AltLane &L, double z, const St &st)
{
    const int max_inner = 200;
    if (L.phase == 0) {
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
        const double coef_h = st.get(o + kAltCoef);
        const double lx3 = ool::log_(X * X * X);
        double g = 1.0;
        double S = alt_coef(0.0, lx3, st.get(o + kAltLd0), X, h, coef_h, g);
        double a_n = S;
        double gt;                                                      // g~(x), :99-108
        if (X > trunc)
            gt = ool::exp_(h * 0.45158270528945486472619522989488 /* log(pi/2) */ + (h - 1) * (lx3 * (1.0 / 3.0))
                           - kPi * kPi * 0.125 * X - st.get(o + kAltLgh));
        else
            gt = h * ool::exp_(h * 0.693147180559945309417232 - 0.5 * (1.8378770664093454835606594728112 + lx3)
                               - 0.5 * h * h / X);
        double Y = src.unif() * gt;
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
    if (nfull > 0) alt_setup(4.0, zh, st.f);
    alt_setup(hrem, zh, st.f + kAltSetupDoubles);
    AltLane L;
    L.start(nfull, nrem);
    while (!alt_trip(s, L, zh, st)) {}
    return L.sum;
}

}  // namespace bl
