// Saddle-point Polya-Gamma sampler J*(n, |z|/2)/n for large shape (fp64).
//
// Reference statements this file has to agree with:
//   v_eval                y -> v            InvertY.cpp:57-99 (+ :10-48)
//   sp_setup              envelope set-up   PolyaGammaSP.cpp:169-226 (+ :78-146)
//   sp_loop               propose / accept  PolyaGammaSP.cpp:228-264 (+ :57-76, :148-167)
//
// The sampler is split in two so that the binned rpg_hybrid path (pg_hybrid.cu) can run the
// set-up and the rejection loop as two kernels with a 12-double state per draw in HBM: one
// monolithic kernel was ~110-140 KB of SASS against a 32 KB instruction cache and spent 15 of
// every 20 stall cycles waiting for instructions (profiles/r1_05_*).  The per-lane kernels and
// the tape path call both halves back to back (sp_draw).
//
// y -> v.  The reference runs Newton from an 81-point grid until |dv| <= 1e-9 -- the root to fp64
// rounding -- and clamps every iterate to the grid bracket.  Here V(y) and G(y) = log cos_rt(V(y))
// come from degree-9 polynomials on 128 binary intervals (tools/gen_sp_tables.py; fp64 Horner
// error < 2e-16 max(1,|.|)), which removes the tan/tanh Newton loop, the sqrt and the log from
// every evaluation.  The reference's own iteration (v_eval_ref) still runs wherever its result is
// NOT the plain root: y outside [2^-4, 2^4) (closed forms), y within 5e-5 (2e-4 above y = 2) grid steps of a grid
// point (iterates clamped to the 7-digit table), and |y - 1| < 1e-6 (series branch).
#pragma once

#include "pg_sp_tables.h"

namespace bl {

// y(v): the series branch is the constant 1 because the reference's coefficients
// (1/3), (2/15), (17/315) are integer divisions (InvertY.cpp:19, PolyaGammaSP.cpp:88).
__device__ __forceinline__ double y_of_v(double v, double tol)
{
    double r = sqrt(fabs(v));
    if (v > tol) return ool::tan_(r) / r;
    if (v < -1 * tol) return ool::tanh_(r) / r;
    return 1.0;
}

// InvertY.cpp:57-99 as written
static __device__ __noinline__ double v_eval_ref(double y)
{
    const double tol = 1e-9;
    const int max_iter = 1000;
    if (y < PG_YGRID[0]) return -1. / (y * y);
    if (y > PG_YGRID[PG_YGRID_LEN - 1]) {
        double v = ool::atan_(0.5 * y * kPi);
        return v * v;
    }
    if (y == 1) return 0.0;
    double id = (ool::log_(y) / ool::log_(2.0) + 4.0) / 0.1;
    int idlow = (int)id;
    int idhigh = idlow + 1;
    if (idhigh > PG_VGRID_LEN - 1) idhigh = PG_VGRID_LEN - 1;  // y == 16 exactly, see DESIGN.md
    double vl = PG_VGRID[idlow];
    double vh = PG_VGRID[idhigh];
    int iter = 0;
    double diff = tol + 1.0;
    double vnew = vl, vold = vl;
    while (diff > tol && iter < max_iter) {
        iter++;
        vold = vnew;
        double yv = y_of_v(vold, 1e-8);
        double f0 = yv - y;
        double f1;
        if (fabs(vold) >= 1e-8)
            f1 = 0.5 * (yv * yv + (1 - yv) / vold);
        else
            f1 = 0.5 * (yv * yv);
        vnew = vold - f0 / f1;
        vnew = vnew > vh ? vh : vnew;
        vnew = vnew < vl ? vl : vnew;
        diff = fabs(vnew - vold);
    }
    return vnew;
}

__device__ __forceinline__ double sp_cos_rt(double v)
{
    double r = sqrt(fabs(v));
    return v >= 0 ? ool::cos_(r) : ool::cosh_(r);
}

// Table path of (V, G); false when the reference's iteration has to decide (see header).
__device__ __forceinline__ bool sp_vg_table(double y, double &v, double &g)
{
    if (!(y >= 0.0625 && y < 16.0)) return false;
    if (fabs(y - 1.0) < 1e-6) return false;
    // Position inside the reference's grid in grid steps; fp32 log2 leaves |err| < 1.5e-5 steps.
    // The grid's 7-digit v values sit within 5e-6 steps of the true roots for y < 2 and within
    // 3.8e-5 steps up to y = 16 (tools/gen_sp_tables.py --check-grid), hence the two margins.
    float tt = (__log2f((float)y) + 4.0f) * 10.0f;
    float fr = tt - floorf(tt);
    float margin = tt > 48.0f ? 2e-4f : 5e-5f;
    if (fr < margin || fr > 1.0f - margin) return false;
    int hi = __double2hiint(y), lo = __double2loint(y);
    int e = (hi >> 20) - 1023;                       // -4 .. 3
    int j = (hi >> (20 - SP_TAB_K)) & ((1 << SP_TAB_K) - 1);
    double f = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);   // mantissa in [1,2)
    double u = (f - 1.0) * (double)(2 << SP_TAB_K) - (double)(2 * j + 1);   // exact
    int row = ((e - SP_TAB_ELO) << SP_TAB_K) + j;
    const double2 *cv = reinterpret_cast<const double2 *>(SP_VTAB + row * (SP_TAB_DEG + 1));
    const double2 *cg = reinterpret_cast<const double2 *>(SP_GTAB + row * (SP_TAB_DEG + 1));
    double2 a = __ldg(cv), b = __ldg(cg);
    double pv = fma(a.x, u, a.y), pg = fma(b.x, u, b.y);
#pragma unroll
    for (int k = 1; k < (SP_TAB_DEG + 1) / 2; ++k) {
        a = __ldg(cv + k);
        b = __ldg(cg + k);
        pv = fma(fma(pv, u, a.x), u, a.y);
        pg = fma(fma(pg, u, b.x), u, b.y);
    }
    v = pv;
    g = pg;
    return true;
}

static __device__ __noinline__ void sp_vg_ref(double y, double &v, double &g)
{
    v = v_eval_ref(y);
    g = ool::log_(sp_cos_rt(v));
}

__device__ __forceinline__ void sp_vg(double y, double &v, double &g)
{
    if (!sp_vg_table(y, v, g)) sp_vg_ref(y, v, g);
}

// what the engine uses for InvertY.cpp's v_eval
__device__ __forceinline__ double v_eval(double y)
{
    double v, g;
    if (sp_vg_table(y, v, g)) return v;
    return v_eval_ref(y);
}

// Envelope of one draw: everything PolyaGammaSP.cpp:176-226 computes before its loop, plus the
// constants of the right piece's left-truncated-gamma sampler (Ch.R:83-114 computes them on
// every call from the same three arguments).
enum SpField {
    kSpMd = 0,   // mid point 1.1 xl, PolyaGammaSP.cpp:183
    kSpPl,       // mass of the left (inverse-Gaussian) piece, :226 (fp64, or an fp32 estimate: see kSpPlBand)
    kSpRt2rl,    // sqrt(2 rl), :216
    kSpRl, kSpIl,   // left tangent line: rate (-slope) and intercept, :209-214
    kSpRr, kSpIr,   // right tangent line
    kSpCl, kSpCr,   // 0.5 log(al) + lcn and 0.5 log(ar) + lcn, :245, :252
    kSpLmd,      // log(md)
    kSpLcz,      // log cosh(|z|/2)
    kSpLcn,      // 0.5 log(n / 2 pi), :215
    kSpLtB, kSpLtC0, kSpLtLM,   // ltgamma(n, n rr, md): b = rate * trunc, c0, log M
    kSpLLtB,     // log b
    kSpMu,       // 1 / sqrt(2 rl), mean of the left piece's inverse Gaussian, :232
    kSpC1md,     // 1 - 1/md
    kSpPlBand,   // 0: kSpPl is the fp64 mass; > 0: kSpPl is an estimate within this distance of it
    kSpStateDoubles
};

// by value (per-lane kernels, tape path)
struct SpState {
    double f[kSpStateDoubles];
    __device__ __forceinline__ double get(int k) const { return f[k]; }
};

// strided view (binned path): field k of this lane's draw at o[k * stride].  The loop kernel
// copies a draw's fields from the HBM struct of arrays into shared memory when the lane takes the
// draw, [field][lane] so the reads are conflict-free; they are then fetched where they are used
// instead of being held in 36 registers across the whole rejection loop.
struct SpStateRef {
    const double *o;
    size_t stride;
    __device__ __forceinline__ double get(int k) const { return o[(size_t)k * stride]; }
};

// PolyaGammaSP.cpp:128-146 (tangent_to_eta) with phi_func :115-126 and delta_func :103-113.
// The tangent lines only shape the envelope's masses and the accept test.  Right of the mid
// point the reference takes log(x) - log(mid) at x = 1.2 xl, mid = 1.1 xl: the constant
// log(12/11) to within 3e-16.
__device__ __forceinline__ void sp_tangent(double x, double z, double mid, double lcz, bool right,
                                           double &slope, double &icept)
{
    double v, g;
    sp_vg(x, v, g);
    double u = 0.5 * v;
    double t = u + 0.5 * z * z;
    double phi_val = lcz - g - t * x;
    double phi_der = -1.0 * t;
    double delta_val, delta_der;
    if (right) {
        delta_val = 0.087011376989629699;   // log(1.2 / 1.1)
        delta_der = 1.0 / x;
    } else {
        delta_val = 0.5 * (1 - 1.0 / x) - 0.5 * (1 - 1.0 / mid);
        delta_der = 0.5 / (x * x);
    }
    double eta_val = phi_val - delta_val;
    double eta_der = phi_der - delta_der;
    slope = eta_der;
    icept = eta_val - eta_der * x;
}

// Weight of the right (gamma) piece as PolyaGammaSP.cpp:220-223 writes it
static __device__ __noinline__ double sp_wr_ref(double hra, double lcn, double n, double rr, double ir,
                                                double lmd, double md)
{
    return ool::exp_(hra + lcn - n * ool::log_(n * rr) + n * ir - n * lmd) * ool::tgamma_(n)
         * (1.0 - p_gamma_rate(md, n, n * rr));
}

// Mass of the left piece, pl = wl / (wl + wr) (PolyaGammaSP.cpp:217-226), fp64.  Written as in the
// reference, wr = exp(.. - n log(n rr) - n log(md)) Gamma(n) (1 - P(n, x)) with x = md n rr.
// For x >= n + 1, Gamma(n) Q(n, x) = e^-x x^n CF(n, x) and the three large terms cancel
// exactly, leaving exp(hra + lcn + n ir - x) CF(n, x).  That form is used while Q stays above
// ~4e-7, i.e. while the written form's 1 - P loses less than 3e-10 of Q to rounding
// (Q >= exp(-n (r-1)^2 / (r+1)) / (sqrt(2 pi n) (r-1)), r = x / n).  Further out in the tail
// the reference's 1 - P rounds to a few ulps or to 0 and pl becomes 1: that behaviour, and
// shapes whose Gamma(n) overflows, keep the written form.
__device__ __forceinline__ bool sp_cf_form_applies(double n, double ltb)
{
    double dx = ltb - n;
    return dx >= 1.0 && n <= 171.0 && dx * dx <= 12.0 * (ltb + n);
}

__device__ __forceinline__ double sp_pl_fp64(double n, double md, double rt2rl, double il, double ir, double rr,
                                             double hla, double hra, double lmd, double lcn, double ltb)
{
    double wl = ool::exp_(hla - n * rt2rl + n * il + 0.5 * n * 1. / md) * p_igauss_direct(md, 1. / rt2rl, n);
    double wr;
    if (sp_cf_form_applies(n, ltb))
        wr = ool::exp_(hra + lcn + n * ir - ltb) * upper_gamma_cf(n, ltb);
    else
        wr = sp_wr_ref(hra, lcn, n, rr, ir, lmd, md);
    return wl / (wl + wr);
}

// fp32 estimate of pl for the binned path's set-up kernel, where the fp64 weights were 40 % of
// the instructions and the FP64 pipe the limiter.  pl only ever meets a uniform, so the loop
// kernel compares against pl_est -+ kSpPlBand and evaluates sp_pl_fp64 (from the state) only inside
// the band.  wr / wl = exp(D) CF / PIG with D = log-ratio of the two exponential prefactors
// (fp64, a dozen FMAs: |D| < 60 or no estimate), CF by modified Lentz in fp32 (<= 64 terms,
// relative error < 1e-5 incl. the float arguments), PIG = Phi(b) + e^(-b^2/2) erfcx(t) / 2 from
// erfcf / erfcxf (4 ulp) with b = s (md Z - 1) formed in fp64 before the cast.  Measured against
// sp_pl_fp64 over the shapes the regime sees: |pl_est - pl| < 2e-6 (test_saddle_point_pl_estimate);
// the band is 2e-5.  Returns a negative value when no estimate is offered.
constexpr double kSpPlBandWidth = 2e-5;

__device__ __forceinline__ float sp_pl_estimate(double n, double md, double rt2rl, double il, double ir,
                                                double lmd, double lcn, double ltb)
{
    if (!sp_cf_form_applies(n, ltb)) return -1.0f;
    const double D = -0.5 * lmd + lcn + n * (ir - il) - ltb + n * rt2rl - 0.5 * n / md;
    if (!(fabs(D) < 60.0)) return -1.0f;
    const float s = sqrtf((float)(n / md));
    const float b = s * (float)(md * rt2rl - 1.0);
    const float t = s * (float)(md * rt2rl + 1.0) * 0.70710678f;
    const float pig = 0.5f * erfcf(-b * 0.70710678f) + 0.5f * erfcxf(t) * __expf(-0.5f * b * b);
    const float hcf = upper_gamma_cf_f32((float)n, (float)ltb);     // Gamma(n, x) e^x x^-n
    if (!(hcf > 0.0f) || !(pig > 0.0f)) return -1.0f;
    const float R = __expf((float)D) * hcf / pig;
    if (!(R >= 0.0f) || isinf(R)) return -1.0f;
    return 1.0f / (1.0f + R);
}

// U < pl ?  (PolyaGammaSP.cpp:231) against the state's pl, exact or estimated
template <class St>
static __device__ __noinline__ double sp_pl_from_state(const St &s, double n)
{
    double lcn = s.get(kSpLcn);
    return sp_pl_fp64(n, s.get(kSpMd), s.get(kSpRt2rl), s.get(kSpIl), s.get(kSpIr), s.get(kSpRr),
                      s.get(kSpCl) - lcn, s.get(kSpCr) - lcn, s.get(kSpLmd), lcn, s.get(kSpLtB));
}

template <class St>
__device__ __forceinline__ bool sp_pick_left(double u, const St &s, double n)
{
    const double pl = s.get(kSpPl), band = s.get(kSpPlBand);
    if (u < pl - band) return true;
    if (u > pl + band || band == 0.0) return false;
    return u < sp_pl_from_state(s, n);
}

// n: shape, zraw: tilting parameter as passed to the sampler (the halving is done here)
template <bool kEstimatePl = false>
__device__ __forceinline__ void sp_setup(double n, double zraw, SpState &s)
{
    double z = 0.5 * fabs(zraw);
    double xl = y_of_v(-1 * z * z, 1e-6);
    double md = xl * 1.1;
    double xr = xl * 1.2;
    double vmd = v_eval(md);
    double K2md;
    if (fabs(vmd) >= 1e-6)
        K2md = md * md + (1 - md) / vmd;
    else
        K2md = md * md;
    double m2 = md * md;
    double al = m2 * md / K2md;
    double ar = m2 / K2md;
    // log cosh z enters both tangent intercepts and the density with the same coefficient n, i.e.
    // F and spa -- and wl and wr -- as one common factor: every decision is invariant to its
    // value.  The binned path therefore takes it in fp32 (its error scales F and spa alike by
    // e^(n 1e-7)); the per-lane / tape path keeps the fp64 value the reference computes.
    double lcz;
    if (kEstimatePl) {
        float zf = (float)z;
        lcz = zf < 12.0f ? (double)__logf(coshf(zf)) : (double)(zf - 0.69314718f);
    } else {
        lcz = ool::log_(ool::cosh_(z));
    }
    double sl, il, sr, ir;
    sp_tangent(xl, z, md, lcz, false, sl, il);
    sp_tangent(xr, z, md, lcz, true, sr, ir);
    double rl = -1. * sl;
    double rr = -1. * sr;
    double lcn = 0.5 * ool::log_(0.5 * n / kPi);
    double rt2rl = sqrt(2 * rl);
    // ar = al / md: 0.5 log(ar) is taken as 0.5 log(al) - 0.5 log(md) (weights and accept test only)
    double lmd = ool::log_(md);
    double hla = 0.5 * ool::log_(al), hra = hla - 0.5 * lmd;
    (void)ar;
    // Proposal weights (:217-226): they only enter the decision U < pl (sp_pl_fp64 / sp_pl_estimate)
    double ltb = (n * rr) * md;
    double pl, pl_band = 0.0;
    float est = kEstimatePl ? sp_pl_estimate(n, md, rt2rl, il, ir, lmd, lcn, ltb) : -1.0f;
    if (est >= 0.0f) {
        pl = (double)est;
        pl_band = kSpPlBandWidth;
    } else {
        pl = sp_pl_fp64(n, md, rt2rl, il, ir, rr, hla, hra, lmd, lcn, ltb);
    }
    // left-truncated gamma constants, Ch.R:96-101 with shape n, rate n rr, truncation md
    double d1 = ltb - n;
    double d3 = n - 1.0;
    double c0 = 0.5 * (d1 + sqrt(d1 * d1 + 4.0 * ltb)) / ltb;
    double l_M = d3 * ool::log_(d3 / (1.0 - c0)) - d3;
    s.f[kSpMd] = md;
    s.f[kSpPl] = pl;
    s.f[kSpPlBand] = pl_band;
    s.f[kSpRt2rl] = rt2rl;
    s.f[kSpRl] = rl;
    s.f[kSpIl] = il;
    s.f[kSpRr] = rr;
    s.f[kSpIr] = ir;
    s.f[kSpCl] = hla + lcn;
    s.f[kSpCr] = hra + lcn;
    s.f[kSpLmd] = lmd;
    s.f[kSpLcz] = lcz;
    s.f[kSpLcn] = lcn;
    s.f[kSpLtB] = ltb;
    s.f[kSpLtC0] = c0;
    s.f[kSpLtLM] = l_M;
    s.f[kSpLLtB] = ool::log_(ltb);
    s.f[kSpMu] = 1. / rt2rl;
    s.f[kSpC1md] = 1. - 1. / md;
}

// PolyaGammaSP.cpp:148-167 (sp_approx)
template <class St>
__device__ __forceinline__ double sp_density(double x, double n, double z, const St &s)
{
    double v, g;
    sp_vg(x, v, g);
    double u = 0.5 * v;
    double z2 = z * z;
    double t = u + 0.5 * z2;
    double phi = s.get(kSpLcz) - g - t * x;
    double K2;
    if (fabs(v) >= 1e-6)
        K2 = x * x + (1 - x) / v;
    else
        K2 = x * x;
    double log_spa = s.get(kSpLcn) - 0.5 * ool::log_(K2) + n * phi;
    return ool::exp_(log_spa);
}

// One lane's position inside the rejection loop PolyaGammaSP.cpp:228-264.  A trip makes ONE
// attempt at a proposal (one inverse-Gaussian draw, or one pass of the truncated-gamma
// rejection) and, once a proposal exists, the accept test; the persistent-lane kernel runs
// trips of many draws side by side, the per-lane kernels just loop over trips.  Variates are
// consumed in the reference's order.
struct SpLane {
    double X;
    int iter;
    int phase;   // 0 pick a piece, 1 left piece, 2 right piece, 3 / 4 proposal ready (left / right)
    __device__ __forceinline__ void start() { X = 2.0; iter = 0; phase = 0; }
};

// returns true when the draw is complete: L.X holds X (omega = n X / 4), L.iter the proposals made.
// z is |z|/2.
template <class Src, class St>
__device__ __forceinline__ bool sp_trip(Src &src, SpLane &L, double n, double z, const St &s)
{
    const int maxiter = 200;
    const double md = s.get(kSpMd);
    if (L.phase == 0) {
        if (L.iter >= maxiter) return true;   // only when maxiter proposals were all rejected
        L.iter++;
        L.phase = sp_pick_left(src.unif(), s, n) ? 1 : 2;
    }
    if (L.phase == 1) {
        double mu = 1. / s.get(kSpRt2rl);
        if (md < mu) {
            double X = md + 1.0, alpha = 0.0;
            while (src.unif() > alpha) {
                X = rtinvchi2(src, n, md);
                alpha = ool::exp_(-0.5 * n / (mu * mu) * X);
            }
            L.X = X;
            L.phase = 3;
        } else {
            double X = igauss(src, mu, n);
            if (!(X > md)) {
                L.X = X;
                L.phase = 3;
            }
        }
    } else if (L.phase == 2) {
        if (n > 1.0 && md > 0.0) {
            // one pass of ltgamma's rejection loop (Ch.R:102-111)
            double b = s.get(kSpLtB), c0 = s.get(kSpLtC0);
            double x = b + src.expon() / c0;
            double u = src.unif();
            double l_rho = (n - 1.0) * ool::log_(x) - x * (1.0 - c0);
            if (ool::log_(u) <= l_rho - s.get(kSpLtLM)) {
                L.X = md * (x / b);
                L.phase = 4;
            }
        } else {
            L.X = ltgamma(src, n, n * s.get(kSpRr), md);
            L.phase = 4;
        }
    }
    if (L.phase >= 3) {
        double X = L.X, F;
        if (L.phase == 3) {
            double phi_ev = n * (s.get(kSpIl) - s.get(kSpRl) * X) + 0.5 * n * ((1. - 1. / X) - (1. - 1. / md));
            F = ool::exp_(s.get(kSpCl) - 1.5 * ool::log_(X) + phi_ev);
        } else {
            double phi_ev = n * (s.get(kSpIr) - s.get(kSpRr) * X) + n * (ool::log_(X) - s.get(kSpLmd));
            F = ool::exp_(s.get(kSpCr) + phi_ev) / X;
        }
        double spa = sp_density(X, n, z, s);
        if (F * src.unif() < spa) return true;
        L.phase = 0;
        if (L.iter >= maxiter) return true;
    }
    return false;
}

// Exact forms of the two decisions the staged trip pre-filters in fp32 (cold code).
static __device__ __noinline__ bool sp_ltgamma_accept_exact(double u, double x, double n, double c0, double l_M)
{
    return ool::log_(u) <= (n - 1.0) * ool::log_(x) - x * (1.0 - c0) - l_M;     // Ch.R:108-110
}

template <class St>
static __device__ __noinline__ bool sp_accept_exact(double X, bool left, double n, double z, double u,
                                                    const St &s)
{
    double F;                                                                     // PolyaGammaSP.cpp:243-258
    if (left) {
        double phi_ev = n * (s.get(kSpIl) - s.get(kSpRl) * X) + 0.5 * n * ((1. - 1. / X) - (1. - 1. / s.get(kSpMd)));
        F = ool::exp_(s.get(kSpCl) - 1.5 * ool::log_(X) + phi_ev);
    } else {
        double phi_ev = n * (s.get(kSpIr) - s.get(kSpRr) * X) + n * (ool::log_(X) - s.get(kSpLmd));
        F = ool::exp_(s.get(kSpCr) + phi_ev) / X;
    }
    return F * u < sp_density(X, n, z, s);
}

// The same trip for Philox streams, laid out so that the lanes of a warp share their expensive
// calls and so that fp64 transcendentals are spent only where a VALUE is produced:
//   * whichever piece a lane proposes from, its first variate needs one fp64 logarithm
//     (Box-Muller radius on the left, exponential on the right): one shared call;
//   * X is computed in fp64 by the formulas of igauss() / ltgamma() from the same words;
//   * the two DECISIONS -- the truncated-gamma accept test (Ch.R:108-110) and F U < spa
//     (PolyaGammaSP.cpp:259) -- are compared in the log domain, linear terms in fp64, the
//     logarithms of X, K2, U and the final exponential in fp32 (MUFU).  Each comparison carries
//     a band >= 2.4x its error bound (__logf: 2^-21.4 abs on [0.5,2], 3 ulp elsewhere; __expf:
//     2 + 1.17|x| ulp; float conversion 2^-24); inside the band, or when an exponent leaves
//     (-700, 700), the reference's fp64 expression decides.  The decision taken is therefore the
//     fp64 decision (up to the reference's own rounding, ~1e-13, far inside the band).
template <class St>
__device__ __forceinline__ bool sp_trip_staged(PhiloxSource &src, SpLane &L, double n, double z, const St &s)
{
    const int maxiter = 200;
    const double md = s.get(kSpMd);
    const double mu = s.get(kSpMu);
    if (!(n > 1.0 && md > 0.0 && md >= mu)) {
        // rare shapes: the generic trip, on COPIES of the stream and lane state -- its out-of-line
        // helpers take the stream by reference, which would otherwise pin `src` in local memory
        // for the whole kernel
        PhiloxSource tsrc = src;
        SpLane tl = L;
        bool r = sp_trip(tsrc, tl, n, z, s);
        src = tsrc;
        L = tl;
        return r;
    }
    if (L.phase == 0) {
        if (L.iter >= maxiter) return true;
        L.iter++;
        L.phase = sp_pick_left(src.unif(), s, n) ? 1 : 2;
    }
    const bool left = L.phase == 1;
    PhiloxSource::LazyN ln = {0u, 0u, 0u};
    PhiloxSource::LazyE le = {0u, 0};
    double a1;
    if (left) {
        ln = src.norm_lazy();
        uint64_t m = ((uint64_t)ln.w0 << 21) | (uint64_t)(ln.w1 >> 11);
        a1 = ((double)m + 0.5) * 0x1p-53;
    } else {
        le = src.expon_lazy();
        a1 = word_to_unif(le.w);
    }
    const double L1 = ool::log_(a1);
    const double u = src.unif();
    double X;
    bool ok;
    float lx;
    if (left) {
        double nu = sqrt(-2.0 * L1) * cospi(2.0 * word_to_unif(ln.w2));
        double y = nu * nu;
        double x = mu + 0.5 * mu * mu * y / n - 0.5 * mu / n * sqrt(4.0 * mu * n * y + (mu * y) * (mu * y));
        if (u > mu / (mu + x)) x = mu * mu / x;
        X = x;
        ok = !(X > md);
        lx = __logf((float)X);
    } else {
        double E = (double)le.k * (32.0 * 0.693147180559945309417232) - L1;
        double c0 = s.get(kSpLtC0), b = s.get(kSpLtB);
        double xg = b + E / c0;
        lx = __logf((float)xg);
        float lu = __logf((float)u);
        double d = (n - 1.0) * (double)lx - xg * (1.0 - c0) - s.get(kSpLtLM) - (double)lu;
        double band = 1e-6 * ((n - 1.0) * (1.0 + fabs((double)lx)) + 2.0 + fabs((double)lu));
        if (d > band) ok = true;
        else if (d < -band) ok = false;
        else ok = sp_ltgamma_accept_exact(u, xg, n, c0, s.get(kSpLtLM));
        X = md * (xg / b);
    }
    if (!ok) return false;
    L.X = X;
    // accept test F U < spa, as log(spa) - log(F) against log U
    const double u2 = src.unif();
    double logX, logF, kx;
    if (left) {
        logX = (double)lx;
        logF = s.get(kSpCl) - 1.5 * logX + n * (s.get(kSpIl) - s.get(kSpRl) * X)
             + 0.5 * n * ((1. - 1. / X) - s.get(kSpC1md));
        kx = 1.5;
    } else {
        double dl = (double)lx - s.get(kSpLLtB);                 // log(x / b) = log(X / md)
        logX = s.get(kSpLmd) + dl;
        logF = s.get(kSpCr) + n * (s.get(kSpIr) - s.get(kSpRr) * X) + n * dl - logX;
        kx = n + 1.0;
    }
    double v, g;
    sp_vg(X, v, g);
    double t = 0.5 * v + 0.5 * (z * z);
    double phi = s.get(kSpLcz) - g - t * X;
    double K2 = fabs(v) >= 1e-6 ? X * X + (1 - X) / v : X * X;
    float lK2 = __logf((float)K2);
    double log_spa = s.get(kSpLcn) - 0.5 * (double)lK2 + n * phi;
    double a = log_spa - logF;
    bool accept;
    if (fabs(logF) < 700.0 && fabs(log_spa) < 700.0) {
        double band = 1e-6 * (kx * (1.0 + fabs((double)lx)) + 0.5 * (1.0 + fabs((double)lK2)));
        if (a >= band) {
            accept = true;                                       // exp(a - band) >= 1 > U
        } else {
            double thr = (double)__expf((float)a);
            double rb = band + 5e-7 * (1.0 + fabs(a));
            if (u2 < thr * (1.0 - rb)) accept = true;
            else if (u2 > thr * (1.0 + rb)) accept = false;
            else accept = sp_accept_exact(X, left, n, z, u2, s);
        }
    } else {
        accept = sp_accept_exact(X, left, n, z, u2, s);
    }
    if (accept) return true;
    L.phase = 0;
    return L.iter >= maxiter;
}

template <class Src>
__device__ int sp_draw(Src &src, double &d, double n, double z)
{
    SpState s;
    sp_setup(n, z, s);
    SpLane L;
    L.start();
    double zh = 0.5 * fabs(z);
    while (!sp_trip(src, L, n, zh, s)) {}
    d = n * 0.25 * L.X;
    return L.iter;
}

}  // namespace bl
