/*
 * oracle/l0.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU stand-in for the reference's missing third-party layer "L0" (jwindle/RNG,
 * unpinned master.zip, fetched by hand per /root/reference/INSTALL:12-27 and
 * git-ignored at /root/reference/.gitignore:69).  The samplers under
 * /root/reference/Code/C call `RNG::{unif, expon_rate, norm, gamma_scale, igauss,
 * ltgamma, rtinvchi2, p_norm, p_gamma_rate, p_igauss, Gamma}`; none of those have
 * a C definition in the reference tree.  This file defines them from
 *   - the in-tree R restatements (cited per function), and
 *   - published algorithms (cited per function) for the special functions.
 *
 * PARITY STATUS OF THIS LAYER: "parity unpinned" -- the reference holds no golden
 * vector for L0; special functions are pinned against scipy.special in
 * tests/test_oracle_l0.py instead.
 *
 * A variate source (`pgo_src`) supplies the four primitive variate kinds
 *   U ~ Uniform(0,1), E ~ Exp(1), N ~ N(0,1), G(a) ~ Gamma(a,1)
 * either from an injected per-observation TAPE (tier-1 parity: the CUDA engine is
 * fed the very same numbers) or from the engine's documented counter-based
 * Philox4x32-10 stream contract (DESIGN.md "Stream contract"), re-implemented
 * here independently in plain C so that seeded GPU draws can be checked
 * draw-for-draw at any size.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or load this.
 */
#ifndef PGO_L0_H
#define PGO_L0_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PGO_MODE_TAPE = 0, PGO_MODE_PHILOX = 1 };

typedef struct pgo_src {
    int mode;
    /* --- tape mode: one segment per observation and per kind --- */
    const double *tu, *te, *tn, *tg;   /* segment base pointers              */
    int lu, le, ln, lg;                /* segment lengths                    */
    int exhausted;                     /* set when a segment ran dry         */
    /* --- philox mode --- */
    uint32_t key0, key1;               /* seed                               */
    uint32_t c0, c1, c3;               /* observation lo/hi, call id         */
    uint32_t blk;                      /* next 4-word block index (ctr word 2)*/
    uint32_t buf[4];
    int pos;                           /* next unread word of buf (4 = empty) */
    /* --- consumption counters (both modes) --- */
    int cu, ce, cn, cg;
} pgo_src;

void pgo_src_tape(pgo_src *s, const double *tu, int lu, const double *te, int le,
                  const double *tn, int ln, const double *tg, int lg);
void pgo_src_philox(pgo_src *s, uint64_t seed, uint64_t obs, uint32_t call_id);

/* Philox4x32-10 block function (Salmon et al., SC'11), exposed for tests. */
void pgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* primitives */
double pgo_unif(pgo_src *s);
double pgo_expon(pgo_src *s);                 /* Exp(1)       */
double pgo_norm(pgo_src *s);                  /* N(0,1)       */
double pgo_gamma(pgo_src *s, double shape);   /* Gamma(a,1)   */

/* composites (call sites: SURVEY.md §8c) */
double pgo_igauss(pgo_src *s, double mu, double lambda);
double pgo_ltgamma(pgo_src *s, double shape, double rate, double trunc);
double pgo_tnorm_left(pgo_src *s, double left);
double pgo_rtinvchi2(pgo_src *s, double scale, double trunc);
double pgo_tnorm(pgo_src *s, double left, double right, double mu, double sd);

/* special functions */
double pgo_p_norm(double x, int use_log);
double pgo_p_gamma_rate(double x, double shape, double rate);  /* P(shape, x*rate) */
double pgo_p_igauss(double x, double mu, double lambda);
double pgo_Gamma(double x, int use_log);

#ifdef __cplusplus
}
#endif
#endif
