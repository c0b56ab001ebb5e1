// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the reference's OWN sampler classes (compiled unmodified and in place
// from /root/reference/Code/C by oracle/Makefile) through the batch interface of
// oracle/batch.h.  Each loop restates the matching serial loop of
// /root/reference/Code/C/LogitWrapper.cpp (cited per function); threading follows
// the reference's own OpenMP pattern, one RNG + one sampler object per thread
// with schedule(dynamic) (PolyaGammaOMP.h:61-73).
#include <cmath>
#include <cstdio>
#include <cstring>

#include "PolyaGamma.h"
#include "PolyaGammaAlt.h"
#include "PolyaGammaSP.h"
#include "InvertY.hpp"
#include "batch.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

inline void open_stream(RNG &r, const pgb_stream *st, int i)
{
    if (st->mode == PGO_MODE_TAPE) {
        pgo_src_tape(&r.src,
                     st->tu ? st->tu + (size_t)i * st->lu : nullptr, st->lu,
                     st->te ? st->te + (size_t)i * st->le : nullptr, st->le,
                     st->tn ? st->tn + (size_t)i * st->ln : nullptr, st->ln,
                     st->tg ? st->tg + (size_t)i * st->lg : nullptr, st->lg);
    } else {
        pgo_src_philox(&r.src, st->seed, st->obs0 + (uint64_t)i, st->call_id);
    }
}

inline void close_stream(const RNG &r, int *trace, int i, int aux)
{
    if (!trace) return;
    int *t = trace + (size_t)i * PGB_TRACE_W;
    t[PGB_TR_U] = r.src.cu;
    t[PGB_TR_E] = r.src.ce;
    t[PGB_TR_N] = r.src.cn;
    t[PGB_TR_G] = r.src.cg;
    t[PGB_TR_EXHAUSTED] = r.src.exhausted;
    t[PGB_TR_AUX] = aux;
}

inline int pick_threads(int nthreads)
{
#ifdef _OPENMP
    return nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    (void)nthreads;
    return 1;
#endif
}

}  // namespace

extern "C" {

const char *pgb_kind(void) { return "reference"; }

// LogitWrapper.cpp:66-85
void pgb_rpg_devroye(double *x, const int *n, const double *z, int num,
                     const pgb_stream *st, int *trace, int nthreads)
{
    int nt = pick_threads(nthreads);
#pragma omp parallel num_threads(nt)
    {
        RNG r;
        PolyaGamma pg(1);
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < num; ++i) {
            open_stream(r, st, i);
            x[i] = n[i] != 0 ? pg.draw(n[i], z[i], r) : 0.0;
            if (r.src.exhausted) x[i] = NAN;
            close_stream(r, trace, i, 0);
        }
    }
}

// LogitWrapper.cpp:39-62
void pgb_rpg_gamma(double *x, const double *n, const double *z, int num, int trunc,
                   const pgb_stream *st, int *trace, int nthreads)
{
    int nt = pick_threads(nthreads);
#pragma omp parallel num_threads(nt)
    {
        RNG r;
        PolyaGamma pg(trunc);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < num; ++i) {
            open_stream(r, st, i);
            x[i] = n[i] != 0.0 ? pg.draw_sum_of_gammas(n[i], z[i], r) : 0.0;
            if (r.src.exhausted) x[i] = NAN;
            close_stream(r, trace, i, 0);
        }
    }
}

// LogitWrapper.cpp:87-106
void pgb_rpg_alt(double *x, const double *h, const double *z, int num,
                 const pgb_stream *st, int *trace, int nthreads)
{
    int nt = pick_threads(nthreads);
#pragma omp parallel num_threads(nt)
    {
        RNG r;
        PolyaGammaAlt pg;
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < num; ++i) {
            open_stream(r, st, i);
            x[i] = h[i] != 0 ? pg.draw(h[i], z[i], r) : 0.0;
            if (r.src.exhausted) x[i] = NAN;
            close_stream(r, trace, i, 0);
        }
    }
}

// LogitWrapper.cpp:108-127 (iter[i] is left untouched when h[i]==0, as there)
void pgb_rpg_sp(double *x, const double *h, const double *z, int num, int *iter,
                const pgb_stream *st, int *trace, int nthreads)
{
    int nt = pick_threads(nthreads);
#pragma omp parallel num_threads(nt)
    {
        RNG r;
        PolyaGammaSP pg;
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < num; ++i) {
            open_stream(r, st, i);
            int it = 0;
            if (h[i] != 0) {
                it = pg.draw(x[i], h[i], z[i], r);
                if (iter) iter[i] = it;
            } else {
                x[i] = 0.0;
            }
            if (r.src.exhausted) x[i] = NAN;
            close_stream(r, trace, i, it);
        }
    }
}

// LogitWrapper.cpp:129-167
void pgb_rpg_hybrid(double *x, const double *h, const double *z, int num,
                    const pgb_stream *st, int *trace, int nthreads)
{
    int nt = pick_threads(nthreads);
#pragma omp parallel num_threads(nt)
    {
        RNG r;
        PolyaGamma dv;
        PolyaGammaAlt alt;
        PolyaGammaSP sp;
#pragma omp for schedule(dynamic, 256)
        for (int i = 0; i < num; ++i) {
            open_stream(r, st, i);
            double b = h[i];
            int aux = 0;
            if (b > 170) {
                double m = dv.pg_m1(b, z[i]);
                double v = dv.pg_m2(b, z[i]) - m * m;
                x[i] = r.norm(m, sqrt(v));
            } else if (b > 13) {
                aux = sp.draw(x[i], b, z[i], r);
            } else if (b == 1 || b == 2) {
                x[i] = dv.draw((int)b, z[i], r);
            } else if (b > 1) {
                x[i] = alt.draw(b, z[i], r);
            } else if (b > 0) {
                x[i] = dv.draw_sum_of_gammas(b, z[i], r);
            } else {
                x[i] = 0.0;
            }
            if (r.src.exhausted) x[i] = NAN;
            close_stream(r, trace, i, aux);
        }
    }
}

double pgb_pg_m1(double b, double z) { return PolyaGamma::pg_m1(b, z); }
double pgb_pg_m2(double b, double z) { return PolyaGamma::pg_m2(b, z); }
double pgb_v_eval(double y) { return v_eval(y); }

}  // extern "C"
