/*
 * oracle/batch.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Common batch interface exported, with identical symbols, by BOTH oracle
 * libraries so tests can swap one for the other:
 *
 *   oracle/_ref/libpg_ref.so      kind "reference": the reference's own sampler
 *                                 sources compiled in place (ref_harness.cpp)
 *   oracle/_build/libpg_oracle.so kind "port": the plain-C restatement
 *                                 (pg_oracle.c)
 *
 * The loops restate /root/reference/Code/C/LogitWrapper.cpp:39-167 (which itself
 * cannot be compiled here: it needs the absent jwindle/Matrix library), with one
 * change of shape that does not alter any draw's arithmetic: each observation
 * has its own variate stream (a tape segment, or a Philox stream keyed by the
 * observation index) instead of one process-global generator, so observation i
 * can be reproduced in isolation by a GPU lane.
 */
#ifndef PGB_BATCH_H
#define PGB_BATCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgb_stream {
    int32_t mode;                      /* 0 = tape, 1 = philox (l0.h)          */
    int32_t lu, le, ln, lg;            /* tape segment lengths per observation */
    const double *tu, *te, *tn, *tg;   /* tapes, [num x l*] row-major          */
    uint64_t seed;                     /* philox key                           */
    uint64_t obs0;                     /* global index of observation 0        */
    uint32_t call_id;                  /* philox counter word 3                */
} pgb_stream;

/* per-observation trace row: variates consumed + status */
enum { PGB_TR_U = 0, PGB_TR_E, PGB_TR_N, PGB_TR_G, PGB_TR_EXHAUSTED, PGB_TR_AUX, PGB_TRACE_W };

const char *pgb_kind(void);

void pgb_rpg_devroye(double *x, const int *n, const double *z, int num,
                     const pgb_stream *st, int *trace, int nthreads);
void pgb_rpg_gamma(double *x, const double *n, const double *z, int num, int trunc,
                   const pgb_stream *st, int *trace, int nthreads);
void pgb_rpg_alt(double *x, const double *h, const double *z, int num,
                 const pgb_stream *st, int *trace, int nthreads);
void pgb_rpg_sp(double *x, const double *h, const double *z, int num, int *iter,
                const pgb_stream *st, int *trace, int nthreads);
void pgb_rpg_hybrid(double *x, const double *h, const double *z, int num,
                    const pgb_stream *st, int *trace, int nthreads);

/* exact moments, PolyaGamma.cpp:208-239 */
double pgb_pg_m1(double b, double z);
double pgb_pg_m2(double b, double z);
/* y -> v inversion, InvertY.cpp:57-99 */
double pgb_v_eval(double y);

#ifdef __cplusplus
}
#endif
#endif
