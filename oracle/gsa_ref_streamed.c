/*
 * gsa_ref_streamed.c -- STREAMED, EXACT-SUM CPU oracle for the Sobol path of GlobalSensitivity.jl.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as gsa_ref.c: only tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs may load it; nothing under globalsensitivity.jl_b200/ does).
 *
 * Why a second oracle: gsa_ref.c keeps the reference's STRUCTURE (it materialises all_points and all_y,
 * src/sobol_sensitivity.jl:128-143), so it cannot run BASELINE configs 4 and 5 at full size (189 GB / 5.5 TB).
 * This file computes the same quantities without ever holding more than one block of samples:
 *
 *   - designs arrive in column chunks (ref_stream_push), any chunking, in any order;
 *   - for every sample the FULL model is evaluated on A_s, B_s and on every swapped point AB_k (and BA_k):
 *     no product-form / additive shortcuts -- the design fusion of :94-108 done one column at a time;
 *   - every per-sample estimator term is formed in fp64 EXACTLY AS THE REFERENCE FORMS IT
 *     (fB*(fAB_k-fA) :192, (fA-fAB_k)^2 :213, fA*(fA-fAB_k) :211, fA*fAB_k :207, fA^2+fAB_k^2 :219, fA+fAB_k :220,
 *      fBA_k*fAB_j - fA*fB :197) and then SUMMED EXACTLY (double-double accumulators, error-free transforms);
 *   - the mean / variance of [fA; fB] (:182-183) come from exact sums of y and y^2, combined in binary128;
 *   - the epilogue (:318-404) is gsa_ref.c's ref_finalize.
 *
 * The only difference from the reference's arithmetic is therefore the SUMMATION ORDER error of Julia's pairwise
 * `sum` (a few 1e-16 relative), which is why tests/test_oracle_golden.py gates this file against gsa_ref.c (and the
 * reference's golden literals) at <= 1e-14 before the GPU tests trust it at BASELINE scale.  Where the two oracles
 * differ by more than that (mean^2 >> var, offsets of 1000 on a unit box), THIS one is the exact value of the
 * reference's formulas -- it is used to adjudicate those cases.
 *
 * Parity status: PINNED (through the gate above).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_API __attribute__((visibility("default")))
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define REF_CLONES __attribute__((target_clones("default", "avx2", "avx512f")))
#else
#define REF_CLONES
#endif

enum { REF_HOMMA1996 = 0, REF_SOBOL2007 = 1, REF_JANSEN1999 = 2, REF_JANON2014 = 3 };
enum { REF_MODEL_ISHIGAMI = 0, REF_MODEL_SOBOL_G = 1, REF_MODEL_LINEAR = 2, REF_MODEL_ISHI_LINEAR = 3 };

/* from gsa_ref.c (same shared object) */
int ref_model_nout(int model, int nparams, int d);
void ref_finalize(const double* Var, const double* Vi, const double* Ei, const double* Vij, int nout, int d, int nboot,
                  int second_order, double conf_level, double* S1, double* S1ci, double* S2, double* S2ci, double* ST,
                  double* STci);
void ref_direction_numbers(const uint32_t* table, int ndim, uint32_t* V);

/* ---------------------------------------------------------------------------------------
 * error-free accumulation
 * ------------------------------------------------------------------------------------- */
typedef struct { double hi, lo; } dd_t;

static inline void dd_acc(dd_t* a, double x) { /* a += x */
    double s = a->hi + x;
    double bb = s - a->hi;
    double e = (a->hi - (s - bb)) + (x - bb);
    e += a->lo;
    double h = s + e;
    a->lo = e - (h - s);
    a->hi = h;
}
static inline void dd_acc_sq(dd_t* a, double x) { /* a += x*x, the product exact */
    double p = x * x;
    double pe = __builtin_fma(x, x, -p);
    dd_acc(a, p);
    dd_acc(a, pe);
}
static inline void dd_merge(dd_t* a, const dd_t* b) {
    dd_acc(a, b->hi);
    dd_acc(a, b->lo);
}

/* ---------------------------------------------------------------------------------------
 * the model on a block of SW samples in structure-of-arrays form: X[i*SW + w], Y[o*SW + w].
 * Per sample the arithmetic (operations AND their order) is that of gsa_ref.c's ref_model_batch, so the values are
 * bit-identical to the materialising oracle's (checked in tests/test_oracle_golden.py); the loops run over w innermost
 * so that the compiler vectorises across samples, which does not change any value (-ffp-contract=off, no fast-math).
 * ------------------------------------------------------------------------------------- */
#define SW 32

REF_CLONES static void model_block(int model, const double* p, int np, int d, int nout, const double* X, double* Y) {
    switch (model) {
        case REF_MODEL_ISHIGAMI: {
            const double a = p[0], b = p[1];
            for (int w = 0; w < SW; ++w) {
                double s1 = sin(X[0 * SW + w]), s2 = sin(X[1 * SW + w]), x3 = X[2 * SW + w];
                Y[w] = s1 + a * (s2 * s2) + b * ((x3 * x3) * (x3 * x3)) * s1;
            }
        } break;
        case REF_MODEL_SOBOL_G: {
            for (int w = 0; w < SW; ++w) Y[w] = 1.0;
            for (int i = 0; i < d; ++i) {
                const double ai = p[i], den = 1.0 + p[i];
                const double* x = X + (size_t)i * SW;
                for (int w = 0; w < SW; ++w) Y[w] *= (fabs(4.0 * x[w] - 2.0) + ai) / den;
            }
        } break;
        case REF_MODEL_LINEAR: {
            for (int o = 0; o < nout; ++o) {
                double* y = Y + (size_t)o * SW;
                for (int w = 0; w < SW; ++w) y[w] = 0.0;
                for (int k = 0; k < d; ++k) {
                    const double c = p[o + (size_t)nout * k];
                    const double* x = X + (size_t)k * SW;
                    for (int w = 0; w < SW; ++w) y[w] += c * x[w];
                }
            }
        } break;
        case REF_MODEL_ISHI_LINEAR: {
            for (int w = 0; w < SW; ++w) {
                double s1 = sin(X[0 * SW + w]), s2 = sin(X[1 * SW + w]), x3 = X[2 * SW + w];
                Y[w] = s1 + 7.0 * (s2 * s2) + 0.1 * ((x3 * x3) * (x3 * x3)) * s1;
                Y[SW + w] = 7.0 * X[0 * SW + w] + 0.1 * X[1 * SW + w];
            }
        } break;
    }
    (void)np;
}

/* ---------------------------------------------------------------------------------------
 * the running state: per (replicate, output) a record of exact sums
 *   0 S  = sum y over [fA; fB]     1 Q  = sum y^2 over [fA; fB]     2 SA = sum fA     3 QA = sum fA^2
 *   4 + k          sum fB (fAB_k - fA)                                  :192
 *   4 + d + k      E0: Jansen (fA-fAB_k)^2 | Sobol2007 fA (fA-fAB_k) | Homma, Janon fA fAB_k
 *   4 + 2d + k     E1: Janon fA^2 + fAB_k^2                             :219
 *   4 + 3d + k     E2: Janon fA + fAB_k                                 :220
 *   4 + 4d + p     sum (fBA_k fAB_j - fA fB), pair p = (k<j) k-major    :197
 * ------------------------------------------------------------------------------------- */
typedef struct ref_stream {
    int model, np, d, nout, nboot, second, est, nrec;
    int64_t N, n, pushed;
    double* params;
    dd_t* rec; /* [nboot][nout][nrec] */
} ref_stream;

static int rec_count(int d, int second) { return 4 + 4 * d + (second ? d * (d - 1) / 2 : 0); }

REF_API ref_stream* ref_stream_new(int model, const double* params, int nparams, int d, int64_t N, int nboot, int second_order,
                                   int ei_estimator) {
    ref_stream* s = (ref_stream*)calloc(1, sizeof(ref_stream));
    if (!s) return NULL;
    s->model = model; s->np = nparams; s->d = d; s->nboot = nboot; s->second = second_order; s->est = ei_estimator;
    s->nout = model >= 0 ? ref_model_nout(model, nparams, d) : nparams /* Y mode: nparams carries nout */;
    s->N = N; s->n = N / nboot;
    s->nrec = rec_count(d, second_order);
    s->params = (double*)malloc(sizeof(double) * (size_t)(nparams > 0 ? nparams : 1));
    if (model >= 0 && nparams > 0) memcpy(s->params, params, sizeof(double) * (size_t)nparams);
    s->rec = (dd_t*)calloc((size_t)nboot * s->nout * s->nrec, sizeof(dd_t));
    if (!s->params || !s->rec) { free(s->params); free(s->rec); free(s); return NULL; }
    return s;
}

REF_API void ref_stream_free(ref_stream* s) {
    if (!s) return;
    free(s->params); free(s->rec); free(s);
}

/* add the terms of one sample (all outputs) to a record set `r` = [nout][nrec]; y arrays are [..][SW] columns at w */
static inline void accumulate_sample(const ref_stream* s, dd_t* r, int w, const double* yA, const double* yB,
                                     const double* yAB /* [d][nout][SW] */, const double* yBA) {
    const int d = s->d, nout = s->nout, est = s->est;
    for (int o = 0; o < nout; ++o) {
        dd_t* q = r + (size_t)o * s->nrec;
        const double fa = yA[o * SW + w], fb = yB[o * SW + w];
        dd_acc(&q[0], fa); dd_acc(&q[0], fb);
        dd_acc_sq(&q[1], fa); dd_acc_sq(&q[1], fb);
        dd_acc(&q[2], fa); dd_acc_sq(&q[3], fa);
        for (int k = 0; k < d; ++k) {
            const double fab = yAB[((size_t)k * nout + o) * SW + w];
            dd_acc(&q[4 + k], fb * (fab - fa));
            switch (est) {
                case REF_JANSEN1999: { double df = fa - fab; dd_acc(&q[4 + d + k], df * df); } break;
                case REF_SOBOL2007: dd_acc(&q[4 + d + k], fa * (fa - fab)); break;
                case REF_HOMMA1996: dd_acc(&q[4 + d + k], fa * fab); break;
                default:
                    dd_acc(&q[4 + d + k], fa * fab);
                    dd_acc(&q[4 + 2 * d + k], fa * fa + fab * fab);
                    dd_acc(&q[4 + 3 * d + k], fa + fab);
                    break;
            }
        }
        if (s->second) {
            const double fafb = fa * fb;
            int p = 0;
            for (int k = 0; k < d; ++k) {
                const double fba = yBA[((size_t)k * nout + o) * SW + w];
                for (int j = k + 1; j < d; ++j, ++p) dd_acc(&q[4 + 4 * d + p], fba * yAB[((size_t)j * nout + o) * SW + w] - fafb);
            }
        }
    }
}

/* Columns [col0, col0 + ncols) of the d x N design (A, B point at column col0; Julia layout: the d values of a sample are
 * contiguous).  Columns at or beyond nboot*n are dropped like the reference drops them (:118). */
REF_API void ref_stream_push(ref_stream* s, const double* A, const double* B, int64_t col0, int64_t ncols) {
    const int d = s->d, nout = s->nout;
    const int64_t n = s->n;
    int64_t hi = col0 + ncols;
    if (hi > n * s->nboot) hi = n * s->nboot;
    if (hi <= col0 || n <= 0) return;
    /* work items: blocks of <= SW samples that do not straddle a replicate */
    int64_t nblk = 0;
    for (int64_t c = col0; c < hi;) {
        int64_t rend = (c / n + 1) * n, seg = (rend < hi ? rend : hi) - c;
        nblk += (seg + SW - 1) / SW;
        c += seg;
    }
    int64_t* blk0 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nblk + 1));
    {
        int64_t i = 0;
        for (int64_t c = col0; c < hi;) {
            int64_t rend = (c / n + 1) * n, e = rend < hi ? rend : hi;
            for (int64_t b = c; b < e; b += SW) blk0[i++] = b;
            c = e;
        }
        blk0[nblk] = hi;
    }
    const size_t recsz = (size_t)nout * s->nrec;
#pragma omp parallel
    {
        double* XA = (double*)malloc(sizeof(double) * (size_t)d * SW);
        double* XB = (double*)malloc(sizeof(double) * (size_t)d * SW);
        double* yA = (double*)malloc(sizeof(double) * (size_t)nout * SW);
        double* yB = (double*)malloc(sizeof(double) * (size_t)nout * SW);
        double* yAB = (double*)malloc(sizeof(double) * (size_t)d * nout * SW);
        double* yBA = s->second ? (double*)malloc(sizeof(double) * (size_t)d * nout * SW) : NULL;
        double keep[SW];
        dd_t* loc = (dd_t*)calloc(recsz, sizeof(dd_t));
        int64_t loc_rep = -1;
#pragma omp for schedule(static)
        for (int64_t bi = 0; bi < nblk; ++bi) {
            const int64_t c = blk0[bi];
            int64_t e = c + SW;
            const int64_t rep = c / n, rend = (rep + 1) * n;
            if (e > rend) e = rend;
            if (e > hi) e = hi;
            const int cnt = (int)(e - c);
            if (rep != loc_rep) {
                if (loc_rep >= 0) {
#pragma omp critical(ref_stream_merge)
                    for (size_t i = 0; i < recsz; ++i) dd_merge(&s->rec[(size_t)loc_rep * recsz + i], &loc[i]);
                    memset(loc, 0, sizeof(dd_t) * recsz);
                }
                loc_rep = rep;
            }
            /* transpose the block; short blocks repeat their last sample (those lanes are not accumulated) */
            for (int w = 0; w < SW; ++w) {
                const int64_t cc = c + (w < cnt ? w : cnt - 1) - col0;
                const double* a = A + (size_t)d * cc;
                const double* b = B + (size_t)d * cc;
                for (int i = 0; i < d; ++i) { XA[(size_t)i * SW + w] = a[i]; XB[(size_t)i * SW + w] = b[i]; }
            }
            model_block(s->model, s->params, s->np, d, nout, XA, yA);
            model_block(s->model, s->params, s->np, d, nout, XB, yB);
            for (int k = 0; k < d; ++k) { /* AB_k: A with parameter k from B (:96-99) */
                double* row = XA + (size_t)k * SW;
                memcpy(keep, row, sizeof(keep));
                memcpy(row, XB + (size_t)k * SW, sizeof(keep));
                model_block(s->model, s->params, s->np, d, nout, XA, yAB + (size_t)k * nout * SW);
                memcpy(row, keep, sizeof(keep));
            }
            if (s->second)
                for (int k = 0; k < d; ++k) { /* BA_k: B with parameter k from A (:100-104) */
                    double* row = XB + (size_t)k * SW;
                    memcpy(keep, row, sizeof(keep));
                    memcpy(row, XA + (size_t)k * SW, sizeof(keep));
                    model_block(s->model, s->params, s->np, d, nout, XB, yBA + (size_t)k * nout * SW);
                    memcpy(row, keep, sizeof(keep));
                }
            for (int w = 0; w < cnt; ++w) accumulate_sample(s, loc, w, yA, yB, yAB, yBA);
        }
        if (loc_rep >= 0) {
#pragma omp critical(ref_stream_merge)
            for (size_t i = 0; i < recsz; ++i) dd_merge(&s->rec[(size_t)loc_rep * recsz + i], &loc[i]);
        }
        free(XA); free(XB); free(yA); free(yB); free(yAB); free(yBA); free(loc);
    }
    free(blk0);
    s->pushed += hi - col0;
}

/* The same sums from a precomputed all_y (nout x nboot*n*step, block order [A | B | AB_1.. | BA_1..] per replicate,
 * :181-190): the exact-sum twin of gsa_ref.c's ref_analysis.  The stream must have been created with model = -1 and
 * nparams = nout. */
REF_API void ref_stream_push_y(ref_stream* s, const double* Y) {
    const int d = s->d, nout = s->nout;
    const int64_t n = s->n;
    const int step = s->second ? 2 * d + 2 : d + 2;
    const size_t recsz = (size_t)nout * s->nrec;
    if (n <= 0) return;
    const int64_t bpr = (n + SW - 1) / SW, nblk = bpr * s->nboot;
#pragma omp parallel
    {
        double* yA = (double*)malloc(sizeof(double) * (size_t)nout * SW);
        double* yB = (double*)malloc(sizeof(double) * (size_t)nout * SW);
        double* yAB = (double*)malloc(sizeof(double) * (size_t)d * nout * SW);
        double* yBA = s->second ? (double*)malloc(sizeof(double) * (size_t)d * nout * SW) : NULL;
        dd_t* loc = (dd_t*)calloc(recsz, sizeof(dd_t));
        int64_t loc_rep = -1;
#pragma omp for schedule(static)
        for (int64_t bi = 0; bi < nblk; ++bi) {
            const int64_t rep = bi / bpr, c = (bi % bpr) * SW;
            const int cnt = (int)((c + SW <= n ? c + SW : n) - c);
            if (rep != loc_rep) {
                if (loc_rep >= 0) {
#pragma omp critical(ref_stream_merge)
                    for (size_t i = 0; i < recsz; ++i) dd_merge(&s->rec[(size_t)loc_rep * recsz + i], &loc[i]);
                    memset(loc, 0, sizeof(dd_t) * recsz);
                }
                loc_rep = rep;
            }
            const double* base = Y + (size_t)nout * n * step * rep;
            for (int w = 0; w < cnt; ++w)
                for (int o = 0; o < nout; ++o) {
                    yA[o * SW + w] = base[o + (size_t)nout * (c + w)];
                    yB[o * SW + w] = base[o + (size_t)nout * (n + c + w)];
                    for (int k = 0; k < d; ++k) {
                        yAB[((size_t)k * nout + o) * SW + w] = base[o + (size_t)nout * ((2 + k) * n + c + w)];
                        if (s->second) yBA[((size_t)k * nout + o) * SW + w] = base[o + (size_t)nout * ((2 + d + k) * n + c + w)];
                    }
                }
            for (int w = 0; w < cnt; ++w) accumulate_sample(s, loc, w, yA, yB, yAB, yBA);
        }
        if (loc_rep >= 0) {
#pragma omp critical(ref_stream_merge)
            for (size_t i = 0; i < recsz; ++i) dd_merge(&s->rec[(size_t)loc_rep * recsz + i], &loc[i]);
        }
        free(yA); free(yB); free(yAB); free(yBA); free(loc);
    }
    s->pushed += n * s->nboot;
}

typedef __float128 q_t; /* arithmetic only (libgcc soft-float): no libquadmath needed */
static inline q_t ddq(const dd_t* a) { return (q_t)a->hi + (q_t)a->lo; }

/* per-replicate Var, Vi, Ei, Vij from the exact sums (:182-236), evaluated in binary128 and rounded once */
REF_API void ref_stream_moments(const ref_stream* s, double* Var, double* Vi, double* Ei, double* Vij) {
    const int d = s->d, nout = s->nout;
    const q_t n = (q_t)s->n, m = 2 * n;
    if (s->second) memset(Vij, 0, sizeof(double) * (size_t)s->nboot * nout * d * d);
    for (size_t ro = 0; ro < (size_t)s->nboot * nout; ++ro) {
        const dd_t* q = s->rec + ro * s->nrec;
        const q_t S = ddq(&q[0]), Q = ddq(&q[1]), SA = ddq(&q[2]), QA = ddq(&q[3]);
        Var[ro] = (double)((Q - S * S / m) / (m - 1));                                        /* :183 */
        for (int k = 0; k < d; ++k) {
            Vi[ro * d + k] = (double)(ddq(&q[4 + k]) / n);                                    /* :192 */
            const q_t e0 = ddq(&q[4 + d + k]);
            q_t e;
            switch (s->est) {
                case REF_JANSEN1999: e = e0 / (2 * n); break;                                  /* :213 */
                case REF_SOBOL2007: e = e0 / n; break;                                         /* :211 */
                case REF_HOMMA1996: {                                                          /* :203-209 */
                    const q_t mA = SA / n;
                    e = (QA - SA * SA / n) / (n - 1) - e0 / n + mA * mA;
                } break;
                default: {                                                                     /* :215-235 */
                    const q_t q2 = ddq(&q[4 + 2 * d + k]), s1 = ddq(&q[4 + 3 * d + k]);
                    const q_t t = s1 / (2 * n), mh = s1 / (2 * n), qh = q2 / 2;
                    e = (q2 / (2 * n) - t * t) * (1 - (e0 / n - mh * mh) / (qh / n - mh * mh));
                } break;
            }
            Ei[ro * d + k] = (double)e;
        }
        if (s->second) {
            int p = 0;
            for (int k = 0; k < d; ++k)
                for (int j = k + 1; j < d; ++j, ++p) Vij[(ro * d + k) * d + j] = (double)(ddq(&q[4 + 4 * d + p]) / n);   /* :197 */
        }
    }
}

REF_API int64_t ref_stream_finish(const ref_stream* s, double conf_level, double* S1, double* S1ci, double* S2, double* S2ci,
                                  double* ST, double* STci) {
    const int d = s->d, nout = s->nout;
    const size_t ro_n = (size_t)s->nboot * nout;
    double* Var = (double*)malloc(sizeof(double) * ro_n);
    double* Vi = (double*)malloc(sizeof(double) * ro_n * d);
    double* Ei = (double*)malloc(sizeof(double) * ro_n * d);
    double* Vij = s->second ? (double*)malloc(sizeof(double) * ro_n * d * d) : NULL;
    ref_stream_moments(s, Var, Vi, Ei, Vij);
    ref_finalize(Var, Vi, Ei, Vij, nout, d, s->nboot, s->second, conf_level, S1, S1ci, S2, S2ci, ST, STci);
    free(Var); free(Vi); free(Ei); free(Vij);
    return s->pushed;
}

/* gsa(f, Sobol(...), A, B; batch=true) in one call on an in-memory design, pushed in chunks of `chunk` columns */
REF_API int64_t ref_gsa_sobol_streamed(int model, const double* params, int nparams, const double* A, const double* B, int d,
                                       int64_t N, int nboot, int second_order, int ei_estimator, double conf_level, int64_t chunk,
                                       double* S1, double* S1ci, double* S2, double* S2ci, double* ST, double* STci) {
    ref_stream* s = ref_stream_new(model, params, nparams, d, N, nboot, second_order, ei_estimator);
    if (!s) return -1;
    if (chunk < 1) chunk = 1 << 16;
    for (int64_t c = 0; c < N; c += chunk) {
        const int64_t nc = c + chunk <= N ? chunk : N - c;
        ref_stream_push(s, A + (size_t)d * c, B + (size_t)d * c, c, nc);
    }
    ref_stream_finish(s, conf_level, S1, S1ci, S2, S2ci, ST, STci);
    const int64_t rows = (N / nboot) * nboot * (second_order ? 2 * d + 2 : d + 2);
    ref_stream_free(s);
    return rows;
}

/* The same with the design GENERATED chunk by chunk (nothing of size N is ever held):
 * design_kind 1: generate_design_matrices(N, lb, ub, SobolSample(), 2)            (README.md:55-56, test/sobol_method.jl:29-30)
 * design_kind 2: the p_range entry (:407-417): one 2*nboot*d-dimensional design of N/nboot points, replicate r =
 *                dimension blocks r (A) and nboot + r (B).
 * Points: Gray-code Sobol integers x, u = x * 2^-32, (ub-lb)*u + lb with separately rounded mul and add (gsa_ref.c). */
static void sobol_chunk(const uint32_t* V /* [dims][32] */, int dim0, int d, uint64_t first, int64_t ncols, const double* lb,
                        const double* ub, double* out /* d x ncols */) {
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncols; ++c) {
        const uint64_t t = first + (uint64_t)c, g = t ^ (t >> 1);
        for (int i = 0; i < d; ++i) {
            const uint32_t* v = V + (size_t)32 * (dim0 + i);
            uint32_t x = 0;
            for (uint64_t gg = g; gg; gg &= gg - 1) x ^= v[__builtin_ctzll(gg)];
            const double u = (double)x * 0x1p-32;
            volatile double prod = (ub[i] - lb[i]) * u;
            out[(size_t)i + (size_t)d * c] = prod + lb[i];
        }
    }
}

REF_API int64_t ref_gsa_sobol_streamed_generated(int model, const double* params, int nparams, const uint32_t* table, const double* lb,
                                                 const double* ub, int design_kind, int d, int64_t N, int nboot, int second_order,
                                                 int ei_estimator, double conf_level, int64_t chunk, double* S1, double* S1ci,
                                                 double* S2, double* S2ci, double* ST, double* STci) {
    ref_stream* s = ref_stream_new(model, params, nparams, d, N, nboot, second_order, ei_estimator);
    if (!s) return -1;
    if (chunk < 1) chunk = 1 << 16;
    const int dims = design_kind == 2 ? 2 * nboot * d : 2 * d;
    uint32_t* V = (uint32_t*)malloc(sizeof(uint32_t) * 32 * (size_t)dims);
    double* A = (double*)malloc(sizeof(double) * (size_t)d * chunk);
    double* B = (double*)malloc(sizeof(double) * (size_t)d * chunk);
    if (!V || !A || !B) { free(V); free(A); free(B); ref_stream_free(s); return -1; }
    ref_direction_numbers(table, dims, V);
    const int64_t npts = design_kind == 2 ? N / nboot : N;   /* points of the sequence */
    uint64_t first = 1;
    while ((first << 1) <= (uint64_t)npts) first <<= 1;
    if (design_kind == 2) {
        for (int r = 0; r < nboot; ++r)
            for (int64_t c = 0; c < npts; c += chunk) {
                const int64_t nc = c + chunk <= npts ? chunk : npts - c;
                sobol_chunk(V, r * d, d, first + (uint64_t)c, nc, lb, ub, A);
                sobol_chunk(V, (nboot + r) * d, d, first + (uint64_t)c, nc, lb, ub, B);
                ref_stream_push(s, A, B, (int64_t)r * npts + c, nc);
            }
    } else {
        for (int64_t c = 0; c < N; c += chunk) {
            const int64_t nc = c + chunk <= N ? chunk : N - c;
            sobol_chunk(V, 0, d, first + (uint64_t)c, nc, lb, ub, A);
            sobol_chunk(V, d, d, first + (uint64_t)c, nc, lb, ub, B);
            ref_stream_push(s, A, B, c, nc);
        }
    }
    ref_stream_finish(s, conf_level, S1, S1ci, S2, S2ci, ST, STci);
    const int64_t rows = (N / nboot) * nboot * (second_order ? 2 * d + 2 : d + 2);
    free(V); free(A); free(B);
    ref_stream_free(s);
    return rows;
}

/* exact-sum twin of ref_analyze_y */
REF_API void ref_analyze_y_exact(const double* Y, int nout, int d, int64_t n, int nboot, int second_order, int ei_estimator,
                                 double conf_level, double* S1, double* S1ci, double* S2, double* S2ci, double* ST, double* STci) {
    ref_stream* s = ref_stream_new(-1, NULL, nout, d, n * nboot, nboot, second_order, ei_estimator);
    if (!s) return;
    ref_stream_push_y(s, Y);
    ref_stream_finish(s, conf_level, S1, S1ci, S2, S2ci, ST, STci);
    ref_stream_free(s);
}
