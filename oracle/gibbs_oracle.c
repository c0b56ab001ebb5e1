/*
 * oracle/gibbs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's Gibbs sweeps that consume PG draws:
 *   binary/binomial logit   /root/reference/Code/C/Logit.hpp:174-183, 283-481
 *                           (driver: LogitWrapper.cpp:176-234)
 *   multinomial logit       /root/reference/Code/C/MultLogit.hpp:214-219, 234-372,
 *                           include/Normal.hpp:98-131 (driver: LogitWrapper.cpp:316-374)
 *   negative binomial       /root/reference/Code/R/NBPG-logmean.R:13-113 (beta | omega, d fixed or sampled:
 *                           draw.df / draw.df.real.mean, Code/R/NB-Shape.R:9-96)
 *   posterior mode by EM    /root/reference/Code/C/Logit.hpp:488-554
 * The reference's own model layer cannot be compiled here (it needs the absent
 * jwindle/Matrix library + BLAS/LAPACK), so dense algebra is
 * written as plain loops with the LAPACK semantics the reference calls
 * (dpotrf/dtrsm/dposv) and omega is drawn with the port sampler of pg_oracle.c.
 * PARITY STATUS: "parity unpinned" for this layer -- no reference output exists
 * to pin it; it is validated against the closed-form posterior in
 * tests/test_gibbs_oracle.py.
 *
 * Stream contract for the sweep: iteration t (0-based, burn-in
 * included) draws omega_i from the Philox stream (seed, obs i, call t) -- for
 * mlogit call t*(J-1)+j -- and the beta draw of that iteration from the stream
 * (seed, obs 2^64-1, same call), consuming variates in the reference's statement
 * order (Appendix A.6 of SURVEY.md).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "batch.h"
#include "l0.h"

#ifdef _OPENMP
#include <omp.h>
#endif

double pgo_dev_draw(pgo_src *s, int n, double z);   /* pg_oracle.c */
double pgo_hybrid_draw(pgo_src *s, double b, double z);

#define BETA_OBS 0xFFFFFFFFFFFFFFFFull

/* ---- dense helpers (column-major, leading dimension = rows) -------------------- */

/* Upper Cholesky A = U'U (LAPACK dpotrf 'U'); A is overwritten, strictly lower part zeroed.
 * Returns 0, or k+1 if the leading minor of order k+1 is not positive definite. */
static int chol_upper(double *A, int P)
{
    for (int j = 0; j < P; ++j) {
        double d = A[j + P * j];
        for (int k = 0; k < j; ++k) d -= A[k + P * j] * A[k + P * j];
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d);
        A[j + P * j] = d;
        for (int i = j + 1; i < P; ++i) {
            double s = A[j + P * i];
            for (int k = 0; k < j; ++k) s -= A[k + P * j] * A[k + P * i];
            A[j + P * i] = s / d;
        }
    }
    for (int j = 0; j < P; ++j)
        for (int i = j + 1; i < P; ++i) A[i + P * j] = 0.0;
    return 0;
}

/* Lower Cholesky A = LL' (dpotrf 'L'). */
static int chol_lower(double *A, int P)
{
    for (int j = 0; j < P; ++j) {
        double d = A[j + P * j];
        for (int k = 0; k < j; ++k) d -= A[j + P * k] * A[j + P * k];
        if (!(d > 0.0)) return j + 1;
        d = sqrt(d);
        A[j + P * j] = d;
        for (int i = j + 1; i < P; ++i) {
            double s = A[i + P * j];
            for (int k = 0; k < j; ++k) s -= A[i + P * k] * A[j + P * k];
            A[i + P * j] = s / d;
        }
    }
    for (int j = 0; j < P; ++j)
        for (int i = 0; i < j; ++i) A[i + P * j] = 0.0;
    return 0;
}

/* x <- U^{-T} x  (trsm 'U','L','T') */
static void solve_Ut(const double *U, double *x, int P)
{
    for (int i = 0; i < P; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= U[k + P * i] * x[k];
        x[i] = s / U[i + P * i];
    }
}

/* x <- U^{-1} x  (trsm 'U','L','N') */
static void solve_U(const double *U, double *x, int P)
{
    for (int i = P - 1; i >= 0; --i) {
        double s = x[i];
        for (int k = i + 1; k < P; ++k) s -= U[i + P * k] * x[k];
        x[i] = s / U[i + P * i];
    }
}

/* x <- L^{-1} x  (trsm 'L','L','N') */
static void solve_L(const double *L, double *x, int P)
{
    for (int i = 0; i < P; ++i) {
        double s = x[i];
        for (int k = 0; k < i; ++k) s -= L[i + P * k] * x[k];
        x[i] = s / L[i + P * i];
    }
}

/* PP = P0 + sum_i w_i x_i x_i'  (Logit.hpp:293-301: scale columns by sqrt(w), syrk) */
static void weighted_gram(double *PP, const double *P0, const double *tX, const double *w,
                          int N, int P, int nthreads)
{
    memcpy(PP, P0, sizeof(double) * P * P);
#pragma omp parallel num_threads(nthreads)
    {
        double *acc = (double *)calloc((size_t)P * P, sizeof(double));
        double *col = (double *)malloc(sizeof(double) * P);
#pragma omp for schedule(static)
        for (int i = 0; i < N; ++i) {
            double rt = sqrt(w[i]);
            for (int a = 0; a < P; ++a) col[a] = tX[a + (size_t)P * i] * rt;
            for (int b = 0; b < P; ++b)
                for (int a = 0; a <= b; ++a) acc[a + P * b] += col[a] * col[b];
        }
#pragma omp critical
        for (int b = 0; b < P; ++b)
            for (int a = 0; a <= b; ++a) PP[a + P * b] += acc[a + P * b];
        free(acc);
        free(col);
    }
    for (int b = 0; b < P; ++b)
        for (int a = 0; a < b; ++a) PP[b + P * a] = PP[a + P * b];
}

/* psi = X beta  (gemm(psi, tX, beta, 'T'), Logit.hpp:421,431) */
static void xbeta(double *psi, const double *tX, const double *beta, int N, int P, int nthreads)
{
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int a = 0; a < P; ++a) s += tX[a + (size_t)P * i] * beta[a];
        psi[i] = s;
    }
}

/* Unconstrained draw beta ~ N(PP^{-1} bP, PP^{-1}).  Logit.hpp:291-320 */
static int draw_beta_plain(double *beta, double *PP, const double *bP, int P, pgo_src *s)
{
    if (chol_upper(PP, P)) return 1;
    double *mP = (double *)malloc(sizeof(double) * P);
    for (int i = 0; i < P; ++i) beta[i] = 1.0 * pgo_norm(s);
    memcpy(mP, bP, sizeof(double) * P);
    solve_Ut(PP, mP, P);
    solve_U(PP, mP, P);
    solve_U(PP, beta, P);
    for (int i = 0; i < P; ++i) beta[i] += mP[i];
    free(mP);
    return 0;
}

/* Constrained coordinate-wise draw, beta_j >= 0 for j < P-1 (the variant the
 * reference actually calls, Logit.hpp:322-400 via :429). */
/* The truncated normal of the coordinate sweeps (r.tnorm at Logit.hpp:393 lives in the reference's absent RNG
 * library, so the construction is this project's; any exact sampler restates it).  Wide windows -- cmin < 1,
 * cmax > -1, cmax - cmin >= 1/2 -- try plain rejection from N(0,1) first: up to four normals from the stream
 * (seed, obs 2^64-3, call), consumed in order; the first one inside (cmin, cmax) is the draw.  Otherwise, or after
 * four misses, pgo_tnorm (inverse CDF / Robert's tail samplers) on the beta stream.  The engine precomputes the
 * rejection normals in parallel, which is the point of the rule (gibbs_beta.cuh). */
#define TN_OBS 0xFFFFFFFFFFFFFFFDull

static int draw_beta_constrained(double *beta, double *PP, const double *bP, const double *beta_prev,
                                 int P, pgo_src *s, uint64_t seed, uint32_t call)
{
    pgo_src sn;
    pgo_src_philox(&sn, seed, TN_OBS, call);
    if (chol_upper(PP, P)) return 1;
    double *S = (double *)calloc((size_t)P * P, sizeof(double));
    double *mP = (double *)malloc(sizeof(double) * P);
    double *z = (double *)malloc(sizeof(double) * P);
    unsigned *is = (unsigned *)malloc(sizeof(unsigned) * P);
    for (int j = 0; j < P; ++j) {
        S[j + P * j] = 1.0;
        solve_Ut(PP, S + (size_t)P * j, P);
        solve_U(PP, S + (size_t)P * j, P);
    }
    int bad = chol_lower(S, P);
    if (!bad) {
        const double *L = S;
        memcpy(mP, bP, sizeof(double) * P);
        solve_Ut(PP, mP, P);
        solve_U(PP, mP, P);
        for (int i = 0; i < P; ++i) {
            z[i] = beta_prev[i] - mP[i];
            beta[i] = beta_prev[i];
        }
        solve_L(L, z, P);
        for (int i = 0; i < P; ++i) is[i] = (unsigned)i;
        for (int k = 0; k < P; ++k) {
            for (int i = 0; i < P - 1; ++i) {
                double f = (double)i + ((double)P - (double)i) * pgo_unif(s);   /* r.flat(i, P) */
                unsigned t = (unsigned)f;
                if (t > (unsigned)(P - 1)) t = (unsigned)(P - 1);
                unsigned tmp = is[i]; is[i] = is[t]; is[t] = tmp;
            }
            for (int i = 0; i < P; ++i) {
                unsigned c = is[i];
                double cmin = -INFINITY, cmax = INFINITY;
                double z1 = z[c];
                for (unsigned j = c; j + 1 < (unsigned)P; ++j) {
                    double l1 = L[j + (size_t)P * c];
                    double c1 = z1 - beta[j] / l1;
                    if (l1 > 0.0 && c1 > cmin) cmin = c1;
                    else if (l1 < 0.0 && c1 < cmax) cmax = c1;
                }
                /* one normal of the rejection stream is always tried; up to three more when the window is wide */
                double z2 = pgo_norm(&sn);
                int got = z2 > cmin && z2 < cmax;
                if (!got && cmin < cmax && cmin < 1.0 && cmax > -1.0 && cmax - cmin >= 0.5)
                    for (int tr = 1; tr < 4 && !got; ++tr) {
                        double Z = pgo_norm(&sn);
                        if (Z > cmin && Z < cmax) { z2 = Z; got = 1; }
                    }
                if (!got) z2 = pgo_tnorm(s, cmin, cmax, 0.0, 1.0);
                z[c] = z2;
                for (unsigned j = c; j < (unsigned)P; ++j) beta[j] += L[j + (size_t)P * c] * (z2 - z1);
            }
        }
    }
    free(S); free(mP); free(z); free(is);
    return bad;
}

/* Logit::gibbs with Logit::gibbs_block's slot semantics (Logit.hpp:402-481):
 * burn-in overwrites slot 0; sampling restarts from slot 0 and writes iteration m
 * into slot m-1.  w: N x samp, beta: P x samp, both zero-initialised by the
 * caller like Matrix::resize does.  Returns 0 on success. */
int pgb_logit_gibbs(double *w, double *beta, const double *y, const double *tX, const double *n,
                    const double *m0, const double *P0, int N, int P, int samp, int burn,
                    uint64_t seed, int constrained, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    double *bP = (double *)calloc(P, sizeof(double));
    double *PP = (double *)malloc(sizeof(double) * P * P);
    double *psi = (double *)malloc(sizeof(double) * N);
    double *bnew = (double *)malloc(sizeof(double) * P);
    /* set_prior: b0 = P0 m0; set_bP: bP = b0 + tX (n o (y - 1/2)).  Logit.hpp:174-190 */
    for (int a = 0; a < P; ++a)
        for (int b = 0; b < P; ++b) bP[a] += P0[a + P * b] * m0[b];
    for (int i = 0; i < N; ++i) {
        double alpha = n[i] * (y[i] - 0.5);
        for (int a = 0; a < P; ++a) bP[a] += tX[a + (size_t)P * i] * alpha;
    }
    memset(w, 0, sizeof(double) * (size_t)N * samp);
    memset(beta, 0, sizeof(double) * (size_t)P * samp);
    int status = 0;
    uint32_t t = 0;
    for (int phase = 0; phase < 2 && !status; ++phase) {
        int iters = phase == 0 ? burn : samp;
        double *bcur = beta, *bprev = beta, *wcur = w;
        xbeta(psi, tX, bcur, N, P, nthreads);
        for (int m = 1; m <= iters && !status; ++m, ++t) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
            for (int i = 0; i < N; ++i) {
                pgo_src s;
                pgo_src_philox(&s, seed, (uint64_t)i, t);
                wcur[i] = pgo_dev_draw(&s, (int)n[i], psi[i]);
            }
            weighted_gram(PP, P0, tX, wcur, N, P, nthreads);
            pgo_src sb;
            pgo_src_philox(&sb, seed, BETA_OBS, t);
            status = constrained ? draw_beta_constrained(bnew, PP, bP, bprev, P, &sb, seed, t)
                                 : draw_beta_plain(bnew, PP, bP, P, &sb);
            memcpy(bcur, bnew, sizeof(double) * P);
            xbeta(psi, tX, bcur, N, P, nthreads);
            if (phase == 1) {   /* period 1: advance the slot after every iteration */
                bprev = bcur;
                if (m < iters) { bcur += P; wcur += N; }
            }
        }
    }
    free(bP); free(PP); free(psi); free(bnew);
    return status;
}

/* ---- multinomial logit ----------------------------------------------------------- */

/* beta ~ N(P1^{-1} b1, P1^{-1}) the way Normal::set_from_likelihood + draw do it:
 * V = P1^{-1} (symsolve on I), mean = V b1, lower = chol(V,'L'), draw = mean + lower*N.
 * include/Normal.hpp:98-131 */
static int draw_mvn_from_likelihood(double *beta, double *P1, const double *b1, int P, pgo_src *s)
{
    if (chol_upper(P1, P)) return 1;
    double *V = (double *)calloc((size_t)P * P, sizeof(double));
    double *e = (double *)malloc(sizeof(double) * P);
    for (int j = 0; j < P; ++j) {
        V[j + P * j] = 1.0;
        solve_Ut(P1, V + (size_t)P * j, P);
        solve_U(P1, V + (size_t)P * j, P);
    }
    for (int a = 0; a < P; ++a) {
        double m = 0.0;
        for (int b = 0; b < P; ++b) m += V[a + P * b] * b1[b];
        beta[a] = m;
    }
    int bad = chol_lower(V, P);
    if (!bad) {
        for (int a = 0; a < P; ++a) e[a] = pgo_norm(s);
        for (int a = 0; a < P; ++a) {
            double v = 0.0;
            for (int b = 0; b <= a; ++b) v += V[a + P * b] * e[b];
            beta[a] += v;
        }
    }
    free(V); free(e);
    return bad;
}

/* MultLogit::gibbs (MultLogit.hpp:261-372).  w: N x (J-1) x samp, beta: P x (J-1) x samp,
 * ty: (J-1) x N, m0: P x (J-1), P0: P x P x (J-1).  burn+1 iterations go into slice 0,
 * then samp-1 more; slice m starts from slice m-1's beta only through XB. */
int pgb_mlogit_gibbs(double *w, double *beta, const double *ty, const double *tX, const double *n,
                     const double *m0, const double *P0, int N, int P, int J, int samp, int burn,
                     uint64_t seed, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    int U = J - 1;
    double *Z = (double *)calloc((size_t)P * U, sizeof(double));
    double *b0 = (double *)calloc((size_t)P * U, sizeof(double));
    double *XB = (double *)calloc((size_t)N * J, sizeof(double));     /* last column stays 0 */
    double *XBno = (double *)malloc(sizeof(double) * (size_t)N * U);
    double *cj = (double *)malloc(sizeof(double) * N);
    double *eta = (double *)malloc(sizeof(double) * N);
    double *P1 = (double *)malloc(sizeof(double) * P * P);
    double *b1 = (double *)malloc(sizeof(double) * P);
    /* Z = tX tkappa', kappa = n (y - 1/2).  MultLogit.hpp:214-219 */
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < U; ++j) {
            double k = n[i] * (ty[j + (size_t)U * i] - 0.5);
            for (int a = 0; a < P; ++a) Z[a + (size_t)P * j] += tX[a + (size_t)P * i] * k;
        }
    for (int j = 0; j < U; ++j)
        for (int a = 0; a < P; ++a)
            for (int b = 0; b < P; ++b)
                b0[a + (size_t)P * j] += P0[a + P * b + (size_t)P * P * j] * m0[b + (size_t)P * j];
    memset(w, 0, sizeof(double) * (size_t)N * U * samp);
    memset(beta, 0, sizeof(double) * (size_t)P * U * samp);
    int status = 0;
    int total = burn + samp;      /* burn+1 into slice 0, samp-1 after */
    for (int t = 0; t < total && !status; ++t) {
        int slice = t <= burn ? 0 : t - burn;
        double *wS = w + (size_t)N * U * slice;
        double *bS = beta + (size_t)P * U * slice;
        /* XB_no_j <- columns 1..J-1 of XB */
        for (int j = 0; j < U; ++j) memcpy(XBno + (size_t)N * j, XB + (size_t)N * (j + 1), sizeof(double) * N);
        for (int j = 0; j < U && !status; ++j) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
            for (int i = 0; i < N; ++i) {
                double A = 0.0;
                for (int k = 0; k < U; ++k) A += exp(XBno[i + (size_t)N * k]);
                cj[i] = log(A);
                eta[i] = XB[i + (size_t)N * j] - cj[i];
            }
            uint32_t call = (uint32_t)t * (uint32_t)U + (uint32_t)j;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
            for (int i = 0; i < N; ++i) {
                pgo_src s;
                pgo_src_philox(&s, seed, (uint64_t)i, call);
                wS[i + (size_t)N * j] = pgo_dev_draw(&s, (int)n[i], eta[i]);
            }
            /* P1 = tX Om X + P0_j ; b1 = Z_j + tX Om c_j + b0_j.  MultLogit.hpp:246-253 */
            for (int a = 0; a < P * P; ++a) P1[a] = 0.0;
            for (int a = 0; a < P; ++a) b1[a] = 0.0;
            for (int i = 0; i < N; ++i) {
                double wi = wS[i + (size_t)N * j];
                const double *xi = tX + (size_t)P * i;
                for (int b = 0; b < P; ++b) {
                    double xw = xi[b] * wi;
                    b1[b] += xw * cj[i];
                    for (int a = 0; a <= b; ++a) P1[a + P * b] += xi[a] * xw;
                }
            }
            for (int b = 0; b < P; ++b)
                for (int a = 0; a <= b; ++a) {
                    P1[a + P * b] += P0[a + P * b + (size_t)P * P * j];
                    P1[b + P * a] = P1[a + P * b];
                }
            for (int a = 0; a < P; ++a) b1[a] = Z[a + (size_t)P * j] + b1[a] + b0[a + (size_t)P * j];
            pgo_src sb;
            pgo_src_philox(&sb, seed, BETA_OBS, call);
            status = draw_mvn_from_likelihood(bS + (size_t)P * j, P1, b1, P, &sb);
            xbeta(XB + (size_t)N * j, tX, bS + (size_t)P * j, N, P, nthreads);
            if (j < U - 1) memcpy(XBno + (size_t)N * j, XB + (size_t)N * j, sizeof(double) * N);
        }
    }
    free(Z); free(b0); free(XB); free(XBno); free(cj); free(eta); free(P1); free(b1);
    return status;
}

/* ---- negative binomial, d fixed ---------------------------------------------------- */

/* NB.PG.gibbs with the dispersion d held fixed (NBPG-logmean.R:77-106 without the
 * draw.df step): psi = X beta - log d; w = rpg(N, y+d, psi) (hybrid); kappa = (y-d)/2;
 * beta ~ N(PN^{-1}(X'(kappa + w log d) + P0 b0), PN^{-1}), PN = X'OmX + P0 (:13-34).
 * beta: P x samp (all iterations kept, no burn-in split), w_last: N. */
int pgb_nb_gibbs(double *w_last, double *beta, const double *y, const double *tX, double d,
                 const double *m0, const double *P0, int N, int P, int samp,
                 uint64_t seed, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    double *b0 = (double *)calloc(P, sizeof(double));
    double *PP = (double *)malloc(sizeof(double) * P * P);
    double *bP = (double *)malloc(sizeof(double) * P);
    double *psi = (double *)malloc(sizeof(double) * N);
    double *bcur = (double *)calloc(P, sizeof(double));
    double ld = log(d);
    for (int a = 0; a < P; ++a)
        for (int b = 0; b < P; ++b) b0[a] += P0[a + P * b] * m0[b];
    int status = 0;
    for (int t = 0; t < samp && !status; ++t) {
        xbeta(psi, tX, bcur, N, P, nthreads);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
        for (int i = 0; i < N; ++i) {
            pgo_src s;
            pgo_src_philox(&s, seed, (uint64_t)i, (uint32_t)t);
            w_last[i] = pgo_hybrid_draw(&s, y[i] + d, psi[i] - ld);
        }
        weighted_gram(PP, P0, tX, w_last, N, P, nthreads);
        memcpy(bP, b0, sizeof(double) * P);
        for (int i = 0; i < N; ++i) {
            double k = 0.5 * (y[i] - d) + w_last[i] * ld;
            for (int a = 0; a < P; ++a) bP[a] += tX[a + (size_t)P * i] * k;
        }
        pgo_src sb;
        pgo_src_philox(&sb, seed, BETA_OBS, (uint32_t)t);
        status = draw_beta_plain(bcur, PP, bP, P, &sb);
        memcpy(beta + (size_t)P * t, bcur, sizeof(double) * P);
    }
    free(b0); free(PP); free(bP); free(psi); free(bcur);
    return status;
}

/* ---------------------------------------------------------------------------------------------
 * NB.PG.gibbs with the dispersion sampled (NBPG-logmean.R:36-113 with draw.df, NB-Shape.R:9-53,
 * kernel 1: random-walk Metropolis on the integers).
 *   G[j] = #{y_i > j}, j = 0..ymax-1                                  NBPG-logmean.R:65-67
 *   df.llh(d) = sum_j log(d+j) G[j] + d sum_i (log d - log(mu_i+d)) + sum_i y_i (log mu_i - log(mu_i+d)),
 *               mu_i = exp(phi_i) (log mu_i taken as phi_i)            NB-Shape.R:9-19
 *   proposal uniform on max(d-1,1) .. d+1; accept with exp(llh(d') - llh(d) + lppsl)  :36-50
 * Per iteration j (reference order, :82-95): phi = X beta; d = draw.df; psi = phi - log d;
 * w = rpg(y + d, psi); kappa = (y - d)/2; beta | w, d.
 * Variates of draw.df: two uniforms from the Philox stream (seed, obs 2^64-2, call j):
 * the pick (index floor(U n) into the grid -- R's sample() lives in R's generator) and the
 * accept test.  beta: P x samp, d_out: samp (recorded past burn-in), w_last: N.
 * --------------------------------------------------------------------------------------------- */
#define DF_OBS 0xFFFFFFFFFFFFFFFEull

static double df_llh(const double *y, double d, const double *phi, const double *G, int ymax, int N)
{
    double llh1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int j = 0; j < ymax; ++j) llh1 += log(d + j) * G[j];
    double ld = log(d);
    for (int i = 0; i < N; ++i) {
        double lmd = log(exp(phi[i]) + d);
        s2 += ld - lmd;
        s3 += y[i] * (phi[i] - lmd);
    }
    return llh1 + d * s2 + s3;
}

/* draw.df.real.mean, NB-Shape.R:86-96 (the commented-out alternative at NBPG-logmean.R:87): random walk on
 * the reals, rstar ~ U(r - 1, r + 1) (U(0, 2) when r <= 1), target sum_i dnbinom(y_i, size r, prob mu_i / (mu_i + r),
 * log = TRUE) exactly as written there.  dnbinom's log density is restated as
 *   lgamma(y + r) - lgamma(r) - lgamma(y + 1) + r log p + y log(1 - p),   log p = phi - log(mu + r),
 *   log(1 - p) = log r - log(mu + r)
 * (R evaluates it through dbinom_raw; same function, different rounding).  Variates: the proposal's uniform, then
 * the uniform of lu = log(runif(1)), both from the stream (seed, obs 2^64-2, call j). */
static double dfreal_ll(const double *y, double r, const double *phi, int N)
{
    double s = 0.0, lgr = lgamma(r), lr = log(r);
    for (int i = 0; i < N; ++i) {
        double lmr = log(exp(phi[i]) + r);
        s += lgamma(y[i] + r) - lgr - lgamma(y[i] + 1.0) + r * (phi[i] - lmr) + y[i] * (lr - lmr);
    }
    return s;
}

static int nb_gibbs_df_impl(double *w_last, double *beta, double *d_out, const double *y, const double *tX,
                            double d0, const double *m0, const double *P0, int N, int P, int samp, int burn,
                            uint64_t seed, int nthreads, int real_d)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    int ymax = 0;
    for (int i = 0; i < N; ++i) if ((int)y[i] > ymax) ymax = (int)y[i];
    double *G = (double *)calloc(ymax > 0 ? ymax : 1, sizeof(double));
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < (int)y[i]; ++j) G[j] += 1.0;
    double *b0 = (double *)calloc(P, sizeof(double));
    double *PP = (double *)malloc(sizeof(double) * P * P);
    double *bP = (double *)malloc(sizeof(double) * P);
    double *phi = (double *)malloc(sizeof(double) * N);
    double *bcur = (double *)calloc(P, sizeof(double));
    for (int a = 0; a < P; ++a)
        for (int b = 0; b < P; ++b) b0[a] += P0[a + P * b] * m0[b];
    double d = d0;
    int status = 0;
    for (int t = 0; t < samp + burn && !status; ++t) {
        xbeta(phi, tX, bcur, N, P, nthreads);
        if (real_d) {   /* draw.df.real.mean */
            pgo_src sd;
            pgo_src_philox(&sd, seed, DF_OBS, (uint32_t)t);
            double u = pgo_unif(&sd);
            double rstar = d > 1.0 ? (d - 1.0) + 2.0 * u : 2.0 * u;
            double lalpha = dfreal_ll(y, rstar, phi, N) - dfreal_ll(y, d, phi, N);
            if (log(pgo_unif(&sd)) < lalpha) d = rstar;
        } else {   /* draw.df */
            pgo_src sd;
            pgo_src_philox(&sd, seed, DF_OBS, (uint32_t)t);
            double lower = d - 1.0 > 1.0 ? d - 1.0 : 1.0;
            int nn = (int)(d + 1.0 - lower) + 1;
            int k = (int)floor(pgo_unif(&sd) * nn);
            if (k > nn - 1) k = nn - 1;
            double dp = lower + k;
            double ltarget = df_llh(y, dp, phi, G, ymax, N) - df_llh(y, d, phi, G, ymax, N);
            double lppsl = log(dp == 1.0 ? 0.5 : 1.0 / 3.0) - log(d == 1.0 ? 0.5 : 1.0 / 3.0);
            if (pgo_unif(&sd) < exp(ltarget + lppsl)) d = dp;
        }
        double ld = log(d);
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
        for (int i = 0; i < N; ++i) {
            pgo_src s;
            pgo_src_philox(&s, seed, (uint64_t)i, (uint32_t)t);
            w_last[i] = pgo_hybrid_draw(&s, y[i] + d, phi[i] - ld);
        }
        weighted_gram(PP, P0, tX, w_last, N, P, nthreads);
        memcpy(bP, b0, sizeof(double) * P);
        for (int i = 0; i < N; ++i) {
            double k = 0.5 * (y[i] - d) + w_last[i] * ld;
            for (int a = 0; a < P; ++a) bP[a] += tX[a + (size_t)P * i] * k;
        }
        pgo_src sb;
        pgo_src_philox(&sb, seed, BETA_OBS, (uint32_t)t);
        status = draw_beta_plain(bcur, PP, bP, P, &sb);
        if (t >= burn) {
            memcpy(beta + (size_t)P * (t - burn), bcur, sizeof(double) * P);
            d_out[t - burn] = d;
        }
    }
    free(G); free(b0); free(PP); free(bP); free(phi); free(bcur);
    return status;
}

int pgb_nb_gibbs_df(double *w_last, double *beta, double *d_out, const double *y, const double *tX,
                    double d0, const double *m0, const double *P0, int N, int P, int samp, int burn,
                    uint64_t seed, int nthreads)
{
    return nb_gibbs_df_impl(w_last, beta, d_out, y, tX, d0, m0, P0, N, P, samp, burn, seed, nthreads, 0);
}

int pgb_nb_gibbs_dfreal(double *w_last, double *beta, double *d_out, const double *y, const double *tX,
                        double d0, const double *m0, const double *P0, int N, int P, int samp, int burn,
                        uint64_t seed, int nthreads)
{
    return nb_gibbs_df_impl(w_last, beta, d_out, y, tX, d0, m0, P0, N, P, samp, burn, seed, nthreads, 1);
}

/* ---------------------------------------------------------------------------------------------
 * Posterior mode by EM, Logit::EM (Logit.hpp:488-554) as LogitWrapper.cpp:238-273 calls it: default (flat)
 * prior, beta = 0 to start; per iteration psi = X beta, w_i = E[omega_i] = n_i tanh(psi_i/2) / (psi_i/2) / 4
 * (series below |psi_i/2| < 0.01, :517-523), PP = X' diag(w) X through the sqrt(w)-scaled copy (:530-541),
 * beta = PP^-1 bP by Cholesky (:543-546), dist = max_a |beta_a - beta_old_a| (:551-552);
 * loop while dist > tol && iter < max_iter.  Returns the iteration count (-1: not positive definite).
 * --------------------------------------------------------------------------------------------- */
int pgb_logit_em(double *beta, const double *y, const double *tX, const double *n, int N, int P,
                 double tol, int max_iter, int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    double *bP = (double *)calloc(P, sizeof(double));
    double *P0 = (double *)calloc((size_t)P * P, sizeof(double));
    double *PP = (double *)malloc(sizeof(double) * P * P);
    double *psi = (double *)malloc(sizeof(double) * N);
    double *w = (double *)malloc(sizeof(double) * N);
    double *old = (double *)malloc(sizeof(double) * P);
    for (int i = 0; i < N; ++i) {
        double alpha = n[i] * (y[i] - 0.5);
        for (int a = 0; a < P; ++a) bP[a] += tX[a + (size_t)P * i] * alpha;
    }
    for (int a = 0; a < P; ++a) beta[a] = 0.0;
    double dist = tol + 1.0;
    int iter = 0;
    while (dist > tol && iter < max_iter) {
        xbeta(psi, tX, beta, N, P, nthreads);
        for (int i = 0; i < N; ++i) {
            double h = psi[i] * 0.5;
            if (fabs(h) < 0.01)
                w[i] = n[i] / cosh(h) * (1 + h * h / 6.0 + pow(h, 4.0) / 120.0 + pow(h, 6) / 5040.0) * 0.25;
            else
                w[i] = n[i] * tanh(h) / h * 0.25;
        }
        memcpy(old, beta, sizeof(double) * P);
        weighted_gram(PP, P0, tX, w, N, P, nthreads);
        if (chol_upper(PP, P)) { iter = -1; break; }
        memcpy(beta, bP, sizeof(double) * P);
        solve_Ut(PP, beta, P);
        solve_U(PP, beta, P);
        dist = 0.0;
        for (int a = 0; a < P; ++a) dist = fmax(dist, fabs(beta[a] - old[a]));
        ++iter;
    }
    free(bP); free(P0); free(PP); free(psi); free(w); free(old);
    return iter;
}

