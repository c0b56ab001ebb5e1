/*
 * oracle/l0.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See l0.h.
 *
 * Stand-in for the reference's absent jwindle/RNG library.  Every function cites
 * the in-tree statement or the published algorithm it follows.
 */
#include "l0.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 -- Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as  */
/* easy as 1, 2, 3", SC'11.  Independent restatement of the stream contract   */
/* in DESIGN.md; the CUDA engine has its own copy in csrc/philox.cuh.         */
/* ------------------------------------------------------------------------- */
void pgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void pgo_src_tape(pgo_src *s, const double *tu, int lu, const double *te, int le,
                  const double *tn, int ln, const double *tg, int lg)
{
    memset(s, 0, sizeof(*s));
    s->mode = PGO_MODE_TAPE;
    s->tu = tu; s->lu = lu;
    s->te = te; s->le = le;
    s->tn = tn; s->ln = ln;
    s->tg = tg; s->lg = lg;
    /* a dry segment falls back to a fixed Philox stream so that rejection loops
     * still terminate; the draw is flagged `exhausted` and discarded by callers */
    s->key0 = 0x9E3779B9u; s->key1 = 0x243F6A88u;
    s->pos = 4;
}

void pgo_src_philox(pgo_src *s, uint64_t seed, uint64_t obs, uint32_t call_id)
{
    memset(s, 0, sizeof(*s));
    s->mode = PGO_MODE_PHILOX;
    s->key0 = (uint32_t)seed;
    s->key1 = (uint32_t)(seed >> 32);
    s->c0 = (uint32_t)obs;
    s->c1 = (uint32_t)(obs >> 32);
    s->c3 = call_id;
    s->blk = 0;
    s->pos = 4;
}

static uint32_t next_word(pgo_src *s)
{
    if (s->pos == 4) {
        uint32_t ctr[4] = { s->c0, s->c1, s->blk, s->c3 };
        uint32_t key[2] = { s->key0, s->key1 };
        pgo_philox4x32_10(ctr, key, s->buf);
        s->blk++;
        s->pos = 0;
    }
    return s->buf[s->pos++];
}

static double word_to_unif(uint32_t w) { return ((double)w + 0.5) * 0x1p-32; }

/* Stream contract: U = (w + 1/2) 2^-32, one word. */
double pgo_unif(pgo_src *s)
{
    int k = s->cu++;
    if (s->mode == PGO_MODE_TAPE) {
        if (k >= s->lu) { s->exhausted = 1; return word_to_unif(next_word(s)); }
        return s->tu[k];
    }
    return word_to_unif(next_word(s));
}

/* Stream contract: E = -log U with the tail extended by memorylessness: a zero
 * word adds 32 log 2 and draws again. */
double pgo_expon(pgo_src *s)
{
    int k = s->ce++;
    if (s->mode == PGO_MODE_TAPE) {
        if (k >= s->le) { s->exhausted = 1; return -log(word_to_unif(next_word(s))); }
        return s->te[k];
    }
    double acc = 0.0;
    uint32_t w = next_word(s);
    while (w == 0u) { acc += 32.0 * 0.693147180559945309417232; w = next_word(s); }
    return acc - log(word_to_unif(w));
}

/* Stream contract: Box-Muller (cosine branch), radius from a 53-bit uniform
 * (two words), angle from one word. */
double pgo_norm(pgo_src *s)
{
    int k = s->cn++;
    if (s->mode == PGO_MODE_TAPE) {
        if (k >= s->ln) { s->exhausted = 1; return 2.0 * word_to_unif(next_word(s)) - 1.0; }
        return s->tn[k];
    }
    uint32_t w0 = next_word(s), w1 = next_word(s), w2 = next_word(s);
    uint64_t m = ((uint64_t)w0 << 21) | (uint64_t)(w1 >> 11);
    double u1 = ((double)m + 0.5) * 0x1p-53;
    double u2 = word_to_unif(w2);
    return sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}

/* Stream contract: Gamma(a,1) by Marsaglia & Tsang (2000), "A simple method for
 * generating gamma variables", without the squeeze step; a < 1 is boosted with
 * G(a+1) U^(1/a).  On a tape, G is a primitive. */
double pgo_gamma(pgo_src *s, double a)
{
    int k = s->cg++;
    if (s->mode == PGO_MODE_TAPE) {
        if (k >= s->lg) { s->exhausted = 1; return -log(word_to_unif(next_word(s))); }
        return s->tg[k];
    }
    double boost = 1.0;
    if (a < 1.0) {
        s->cu--; /* internal draws of a composite do not count as user variates */
        boost = exp(log(pgo_unif(s)) / a);
        a += 1.0;
    }
    double d = a - 1.0 / 3.0;
    double c = 1.0 / sqrt(9.0 * d);
    for (;;) {
        s->cn--; s->cu--;
        double x = pgo_norm(s);
        double u = pgo_unif(s);
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
    }
}

/* ------------------------------------------------------------------------- */
/* Composites                                                                 */
/* ------------------------------------------------------------------------- */

/* Inverse-Gaussian(mu, lambda) by Michael, Schucany & Haas (1976); statement
 * followed: /root/reference/Code/R/Ch.R:403-413 (rigauss).  Pops N then U. */
double pgo_igauss(pgo_src *s, double mu, double lambda)
{
    double nu = pgo_norm(s);
    double y = nu * nu;
    double x = mu + 0.5 * mu * mu * y / lambda
             - 0.5 * mu / lambda * sqrt(4.0 * mu * lambda * y + (mu * y) * (mu * y));
    if (pgo_unif(s) > mu / (mu + x)) x = mu * mu / x;
    return x;
}

/* Left-truncated Gamma(shape, rate) on (trunc, inf), shape >= 1, by Dagpunar
 * (1978); statement followed: /root/reference/Code/R/Ch.R:83-114
 * (rltgamma.dagpunar.1).  Pops E when shape == 1, else (E U)+. */
double pgo_ltgamma(pgo_src *s, double shape, double rate, double trunc)
{
    double a = shape;
    double b = rate * trunc;
    if (trunc <= 0.0 || shape < 1.0) return 0.0;
    if (shape == 1.0) return pgo_expon(s) / rate + trunc;
    double d1 = b - a;
    double d3 = a - 1.0;
    double c0 = 0.5 * (d1 + sqrt(d1 * d1 + 4.0 * b)) / b;
    double l_M = d3 * log(d3 / (1.0 - c0)) - d3;
    double x;
    for (;;) {
        x = b + pgo_expon(s) / c0;
        double u = pgo_unif(s);
        double l_rho = d3 * log(x) - x * (1.0 - c0);
        if (log(u) <= l_rho - l_M) break;
    }
    return trunc * (x / b);
}

/* One-sided truncated standard normal on (left, inf).  left >= 0: exponential
 * rejection sampler of Robert (1995), "Simulation of truncated normal
 * variables", Stat. Comput. 5:121-125, with the optimal rate
 * a* = (left + sqrt(left^2+4))/2; pops (E U)+.  left < 0: plain rejection from
 * N(0,1); pops N+. */
double pgo_tnorm_left(pgo_src *s, double left)
{
    if (left < 0.0) {
        for (;;) {
            double z = pgo_norm(s);
            if (z > left) return z;
        }
    }
    double astar = 0.5 * (left + sqrt(left * left + 4.0));
    for (;;) {
        double z = pgo_expon(s) / astar + left;
        double rho = exp(-0.5 * (z - astar) * (z - astar));
        if (pgo_unif(s) < rho) return z;
    }
}

/* Right-truncated scaled inverse chi^2(1): X = scale / Z^2, Z ~ N(0,1) truncated
 * to (1/sqrt(trunc/scale), inf).  Statement followed:
 * /root/reference/Code/R/SPSample.R:534-550 (rrtinvch2.1, truncated-normal form).
 * This is what RNG::rtinvchi2 (call site PolyaGammaSP.cpp:64) is taken to be;
 * the Alt sampler uses its own in-tree exponential-pair form
 * (PolyaGammaAlt.cpp:6-22), restated in pg_oracle.c. */
double pgo_rtinvchi2(pgo_src *s, double scale, double trunc)
{
    double R = trunc / scale;
    double z = pgo_tnorm_left(s, 1.0 / sqrt(R));
    return scale / (z * z);
}

/* Two-sided truncated normal N(mu, sd^2) on (left, right), used only by the
 * constrained beta draw (Logit.hpp:393).  No in-tree statement exists (the
 * reference's RNG::tnorm is in the absent library).  Standardised bounds (a, b):
 *   far tails (a >= 4, or b <= -4 mirrored): rejection samplers of Robert (1995) --
 *     narrow interval ((b-a) a < 1): uniform proposal, accept exp((a^2 - z^2)/2), pops (U U)+;
 *     otherwise translated-exponential proposal with rate a* = (a + sqrt(a^2+4))/2,
 *     rejected beyond b, accept exp(-(z - a*)^2/2), pops (E [U])+;
 *   elsewhere: inverse CDF on one U, evaluated on whichever tail keeps precision. */
static double inv_phi_upper(double q);

static double tnorm_tail(pgo_src *s, double a, double b)
{
    if (b < INFINITY && (b - a) * a < 1.0) {
        for (;;) {
            double z = a + (b - a) * pgo_unif(s);
            if (pgo_unif(s) < exp(0.5 * (a * a - z * z))) return z;
        }
    }
    double astar = 0.5 * (a + sqrt(a * a + 4.0));
    for (;;) {
        double z = a + pgo_expon(s) / astar;
        if (z > b) continue;
        if (pgo_unif(s) < exp(-0.5 * (z - astar) * (z - astar))) return z;
    }
}

double pgo_tnorm(pgo_src *s, double left, double right, double mu, double sd)
{
    double a = (left - mu) / sd, b = (right - mu) / sd;
    double z;
    if (!(a < b)) return mu + sd * a;
    if (b <= -4.0) return mu - sd * tnorm_tail(s, -b, -a);
    if (a >= 4.0) return mu + sd * tnorm_tail(s, a, b);
    double u = pgo_unif(s);
    if (a >= 0.0 || (a > -INFINITY && -a < b)) {
        /* work with upper tails Q(x) = 1 - Phi(x) */
        double qa = isinf(a) ? 1.0 : 0.5 * erfc(a / M_SQRT2);
        double qb = isinf(b) ? 0.0 : 0.5 * erfc(b / M_SQRT2);
        z = inv_phi_upper(qa - u * (qa - qb));
    } else {
        double pa = isinf(a) ? 0.0 : 0.5 * erfc(-a / M_SQRT2);
        double pb = isinf(b) ? 1.0 : 0.5 * erfc(-b / M_SQRT2);
        z = -inv_phi_upper(pa + u * (pb - pa));
    }
    if (z < a) z = a;
    if (z > b) z = b;
    return mu + sd * z;
}

/* x with Q(x) = q, by Newton on log Q from the Acklam-style start; q in (0,1). */
static double inv_phi_upper(double q)
{
    if (q <= 0.0) return INFINITY;
    if (q >= 1.0) return -INFINITY;
    if (q > 0.5) return -inv_phi_upper(1.0 - q);
    double t = sqrt(-2.0 * log(q));
    double x = t - (2.515517 + 0.802853 * t + 0.010328 * t * t)
                 / (1.0 + 1.432788 * t + 0.189269 * t * t + 0.001308 * t * t * t);
    for (int it = 0; it < 8; ++it) {
        double Q = 0.5 * erfc(x / M_SQRT2);
        double pdf = exp(-0.5 * x * x) / sqrt(2.0 * M_PI);
        double dx = (Q - q) / pdf;
        x += dx;
        if (fabs(dx) < 1e-15 * (1.0 + fabs(x))) break;
    }
    return x;
}

/* ------------------------------------------------------------------------- */
/* Special functions                                                          */
/* ------------------------------------------------------------------------- */

/* Phi(x) and log Phi(x).  log: erfc form where erfc keeps relative accuracy,
 * asymptotic (Abramowitz & Stegun 26.2.12) expansion below -20 where it
 * underflows/loses digits.  Call sites: PolyaGamma.cpp:74-75 (log), :61,
 * PolyaGammaAlt.cpp:56. */
double pgo_p_norm(double x, int use_log)
{
    if (!use_log) return 0.5 * erfc(-x / M_SQRT2);
    if (x > 0.0) return log1p(-0.5 * erfc(x / M_SQRT2));
    if (x > -20.0) return log(0.5 * erfc(-x / M_SQRT2));
    double x2 = x * x;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k <= 30; ++k) {
        double nt = -term * (2.0 * k - 1.0) / x2;
        if (fabs(nt) >= fabs(term)) break;
        term = nt;
        sum += term;
        if (fabs(term) < 1e-17) break;
    }
    return -0.5 * x2 - log(-x) - 0.5 * log(2.0 * M_PI) + log(sum);
}

/* log of x^a e^-x / Gamma(a).  For a >= 10 the three ~a-sized terms of the direct
 * form cancel to O(1) and lose ~log10(a) digits, so the prefix is rearranged
 * around Stirling's series: -a (mu - log1p mu) - S(a) + log sqrt(a/2pi),
 * mu = (x-a)/a, S(a) = 1/(12a) - 1/(360a^3) + ... (Temme 1979; the same
 * rearrangement Boost.Math calls regularised_gamma_prefix). */
static double gamma_log_prefix(double a, double x)
{
    if (a < 10.0) return a * log(x) - x - lgamma(a);
    double mu = (x - a) / a;
    double phi = mu - log1p(mu);
    double ia = 1.0 / a, ia2 = ia * ia;
    double S = ia * (1.0 / 12 - ia2 * (1.0 / 360 - ia2 * (1.0 / 1260 - ia2 * (1.0 / 1680
             - ia2 * (1.0 / 1188 - ia2 * (691.0 / 360360 - ia2 * (1.0 / 156)))))));
    return -a * phi - S + 0.5 * log(a / (2.0 * 3.14159265358979323846));
}

/* Regularised lower incomplete gamma P(a, x) (Numerical-Recipes-style split:
 * power series for x < a+1, modified-Lentz continued fraction for Q otherwise,
 * returning 1-Q).  Callers always use `1.0 - P` (PolyaGammaAlt.cpp:66,73;
 * PolyaGammaSP.cpp:222), so absolute accuracy ~1e-16 is what matters. */
static double pgamma_lower(double a, double x)
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
        d = an * d + b; if (fabs(d) < tiny) d = tiny;
        c = b + an / c; if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return 1.0 - exp(lpre) * h;
}

double pgo_p_gamma_rate(double x, double shape, double rate)
{
    return pgamma_lower(shape, x * rate);
}

/* Inverse-Gaussian CDF in log space; statement followed:
 * /root/reference/Code/R/PG.R:15-23.  Call site PolyaGammaSP.cpp:218. */
double pgo_p_igauss(double x, double mu, double lambda)
{
    double Z = 1.0 / mu;
    double b = sqrt(lambda / x) * (x * Z - 1.0);
    double a = -1.0 * sqrt(lambda / x) * (x * Z + 1.0);
    return exp(pgo_p_norm(b, 1)) + exp(2.0 * lambda * Z + pgo_p_norm(a, 1));
}

/* Gamma(x) / log Gamma(x).  Call sites PolyaGammaAlt.cpp:103 (log),
 * PolyaGammaSP.cpp:222. */
double pgo_Gamma(double x, int use_log)
{
    return use_log ? lgamma(x) : tgamma(x);
}
