/*
 * gsa_ref.c -- CPU oracle (plain C + OpenMP) for the Sobol path of GlobalSensitivity.jl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under globalsensitivity.jl_b200/ links, loads or calls
 * this file; it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs as the checker and as the timed CPU baseline ("port": Julia is not
 * available in the image, so the reference itself cannot run).
 *
 * Parity status: PINNED through oracle/sobol_oracle.py (same algorithm; that twin reproduces
 * the reference's literal goldens test/sobol_method.jl:41-45,55-71) -- tests/test_oracle_golden.py
 * checks this C file against those goldens directly as well.
 *
 * It keeps the REFERENCE'S STRUCTURE on purpose (so the timing is a fair restatement):
 *   ref_all_points   materialises [A,B,AB_1..AB_d(,BA_1..BA_d)] per replicate
 *                                        src/sobol_sensitivity.jl:94-108,116-134
 *   ref_model_batch  one batch model call over all columns                 :143
 *   ref_analysis     slice-and-sum estimators, pairwise sums               :169-350
 *   ref_finalize     mean / z*std/sqrt(nboot) over replicates              :351-404
 * All matrices use Julia's column-major layout: X[i + d*c] is parameter i of column c,
 * Y[o + nout*c] is output o of column c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_API __attribute__((visibility("default")))

enum { REF_HOMMA1996 = 0, REF_SOBOL2007 = 1, REF_JANSEN1999 = 2, REF_JANON2014 = 3 };
enum { REF_MODEL_ISHIGAMI = 0, REF_MODEL_SOBOL_G = 1, REF_MODEL_LINEAR = 2, REF_MODEL_ISHI_LINEAR = 3 };

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline should use every core it may run on */
REF_API void ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

REF_API int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------
 * Sobol points.  Third-party in the reference (QuasiMonteCarlo.jl -> Sobol.jl, not vendored);
 * published algorithm: Bratley-Fox / Antonov-Saleev Gray code with Joe-Kuo 21201 numbers.
 * `table` = 21201 x 19 uint32 records [poly, m_1..m_18] (tools/make_direction_table.py).
 * ------------------------------------------------------------------------------------- */
REF_API void ref_direction_numbers(const uint32_t* table, int ndim, uint32_t* V /* ndim x 32 */) {
    for (int j = 0; j < ndim; ++j) {
        uint64_t m[32];
        if (j == 0) {
            for (int i = 0; i < 32; ++i) m[i] = 1;
        } else {
            uint32_t p = table[19 * j];
            int deg = 31 - __builtin_clz(p);
            for (int i = 0; i < deg; ++i) m[i] = table[19 * j + 1 + i];
            for (int i = deg; i < 32; ++i) {
                uint64_t v = m[i - deg] ^ (m[i - deg] << deg);
                for (int k = 1; k < deg; ++k)
                    if ((p >> (deg - k)) & 1u) v ^= m[i - k] << k;
                m[i] = v;
            }
        }
        for (int i = 0; i < 32; ++i) V[32 * j + i] = (uint32_t)((m[i] << (31 - i)) & 0xFFFFFFFFu);
    }
}

/* out[m] (m < num_mats) is d x ncols; column c is sequence index 2^floor(log2 n) + col0 + c. */
REF_API void ref_sobol_design(const uint32_t* table, int64_t n, int d, int num_mats, const double* lb,
                              const double* ub, int64_t col0, int64_t ncols, double* out) {
    int nd = d * num_mats;
    uint32_t* V = (uint32_t*)malloc(sizeof(uint32_t) * 32 * (size_t)nd);
    ref_direction_numbers(table, nd, V);
    uint64_t first = 1;
    while ((first << 1) <= (uint64_t)n) first <<= 1;
    first += (uint64_t)col0;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncols; ++c) {
        uint64_t t = first + (uint64_t)c, g = t ^ (t >> 1);
        for (int j = 0; j < nd; ++j) {
            uint32_t x = 0;
            uint64_t gg = g;
            while (gg) {
                int b = __builtin_ctzll(gg);
                x ^= V[32 * j + b];
                gg &= gg - 1;
            }
            double u = (double)x * 0x1p-32;
            int m = j / d, i = j % d;
            volatile double prod = (ub[i] - lb[i]) * u; /* separately rounded: no FMA contraction */
            out[(size_t)m * d * ncols + (size_t)i + (size_t)d * c] = prod + lb[i];
        }
    }
    free(V);
}

/* ---------------------------------------------------------------------------------------
 * Design fusion  (:94-108, :116-134)
 * ------------------------------------------------------------------------------------- */
REF_API int64_t ref_all_points_cols(int d, int64_t N, int nboot, int second_order) {
    int64_t n = N / nboot;
    return n * (second_order ? 2 * d + 2 : d + 2) * nboot;
}

REF_API void ref_all_points(const double* A, const double* B, int d, int64_t N, int nboot,
                            int second_order, double* P /* d x cols */) {
    int64_t n = N / nboot;
    int step = second_order ? 2 * d + 2 : d + 2;
    for (int r = 0; r < nboot; ++r) {
        const double* Ar = A + (size_t)d * n * r;
        const double* Br = B + (size_t)d * n * r;
        double* Pr = P + (size_t)d * n * step * r;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < n; ++c) {
            const double* a = Ar + (size_t)d * c;
            const double* b = Br + (size_t)d * c;
            memcpy(Pr + (size_t)d * c, a, sizeof(double) * d);
            memcpy(Pr + (size_t)d * (n + c), b, sizeof(double) * d);
            for (int k = 0; k < d; ++k) {
                double* q = Pr + (size_t)d * ((2 + k) * n + c);
                memcpy(q, a, sizeof(double) * d);
                q[k] = b[k];
            }
            if (second_order)
                for (int k = 0; k < d; ++k) {
                    double* q = Pr + (size_t)d * ((2 + d + k) * n + c);
                    memcpy(q, b, sizeof(double) * d);
                    q[k] = a[k];
                }
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * Batch models (the user's `f` in the reference; shapes of test/sobol_method.jl:4-24,146-158)
 * ------------------------------------------------------------------------------------- */
REF_API int ref_model_nout(int model, int nparams, int d) {
    if (model == REF_MODEL_LINEAR) return nparams / d;
    if (model == REF_MODEL_ISHI_LINEAR) return 2;
    return 1;
}

REF_API void ref_model_batch(int model, const double* params, int nparams, const double* X, int d,
                             int64_t cols, double* Y) {
    int nout = ref_model_nout(model, nparams, d);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < cols; ++c) {
        const double* x = X + (size_t)d * c;
        double* y = Y + (size_t)nout * c;
        switch (model) {
            case REF_MODEL_ISHIGAMI: {
                double a = params[0], b = params[1], s1 = sin(x[0]), s2 = sin(x[1]);
                double x3 = x[2];
                y[0] = s1 + a * (s2 * s2) + b * ((x3 * x3) * (x3 * x3)) * s1;
            } break;
            case REF_MODEL_SOBOL_G: {
                double p = 1.0;
                for (int i = 0; i < d; ++i) p *= (fabs(4.0 * x[i] - 2.0) + params[i]) / (1.0 + params[i]);
                y[0] = p;
            } break;
            case REF_MODEL_LINEAR: { /* params = W, nout x d column-major */
                for (int o = 0; o < nout; ++o) y[o] = 0.0;
                for (int k = 0; k < d; ++k)
                    for (int o = 0; o < nout; ++o) y[o] += params[o + (size_t)nout * k] * x[k];
            } break;
            case REF_MODEL_ISHI_LINEAR: {
                double s1 = sin(x[0]), s2 = sin(x[1]), x3 = x[2];
                y[0] = s1 + 7.0 * (s2 * s2) + 0.1 * ((x3 * x3) * (x3 * x3)) * s1;
                y[1] = 7.0 * x[0] + 0.1 * x[1];
            } break;
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * Julia Base numerics: pairwise sum (block 1024), two-pass corrected variance.
 * Strided so that a row of an nout x cols matrix can be reduced without a copy.
 * ------------------------------------------------------------------------------------- */
typedef double (*term_fn)(const double* const* s, int64_t i, int64_t stride, const double* aux);

static double pw_sum(term_fn f, const double* const* s, int64_t lo, int64_t hi, int64_t stride,
                     const double* aux) {
    if (hi - lo <= 1024) {
        double acc = 0.0;
        for (int64_t i = lo; i < hi; ++i) acc += f(s, i, stride, aux);
        return acc;
    }
    int64_t mid = (lo + hi) >> 1;
    return pw_sum(f, s, lo, mid, stride, aux) + pw_sum(f, s, mid, hi, stride, aux);
}

#define AT(p, i) ((p)[(i) * stride])
static double t_id(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return AT(s[0], i); }
static double t_csq(const double* const* s, int64_t i, int64_t stride, const double* a) { double v = AT(s[0], i) - a[0]; return v * v; }
/* fB .* (fAi .- fA)                        :192  s = {fB, fAi, fA} */
static double t_vi(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return AT(s[0], i) * (AT(s[1], i) - AT(s[2], i)); }
/* (fBi[k] .* fAi[j]) .- (fA .* fB)         :197  s = {fBk, fAj, fA, fB} */
static double t_vij(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return AT(s[0], i) * AT(s[1], i) - AT(s[2], i) * AT(s[3], i); }
/* abs2(fA - fAi)                           :213  s = {fA, fAi} */
static double t_jansen(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; double v = AT(s[0], i) - AT(s[1], i); return v * v; }
/* fA .* (fA .- fAi)                        :211 */
static double t_sobol07(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return AT(s[0], i) * (AT(s[0], i) - AT(s[1], i)); }
/* fA .* fAi                                :207 */
static double t_prod(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return AT(s[0], i) * AT(s[1], i); }
/* fA.^2 + fAi.^2                           :219 */
static double t_sq2(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return AT(s[0], i) * AT(s[0], i) + AT(s[1], i) * AT(s[1], i); }
static double t_sq2h(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return (AT(s[0], i) * AT(s[0], i) + AT(s[1], i) * AT(s[1], i)) / 2; }
/* fA + fAi                                 :220 */
static double t_add(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return AT(s[0], i) + AT(s[1], i); }
static double t_addh(const double* const* s, int64_t i, int64_t stride, const double* a) { (void)a; return (AT(s[0], i) + AT(s[1], i)) / 2; }

/* mean/var over the 2n contiguous columns [fA; fB] (:182-183): the two blocks are adjacent in all_y */
static double var2n(const double* yAB, int64_t n2, int64_t stride) {
    const double* s[1] = {yAB};
    double m = pw_sum(t_id, s, 0, n2, stride, NULL) / (double)n2;
    return pw_sum(t_csq, s, 0, n2, stride, &m) / (double)(n2 - 1);
}

/* ---------------------------------------------------------------------------------------
 * Analysis (:169-350).  Outputs, per replicate r and output o:
 *   Var[r*nout+o], Vi/Ei[(r*nout+o)*d + k], Vij[((r*nout+o)*d + k)*d + j] (k<j, else 0)
 * ------------------------------------------------------------------------------------- */
REF_API void ref_analysis(const double* Y, int nout, int d, int64_t n, int nboot, int second_order,
                          int ei_estimator, double* Var, double* Vi, double* Ei, double* Vij) {
    int step = second_order ? 2 * d + 2 : d + 2;
    int64_t stride = nout;
    double dn = (double)n;
    if (second_order) memset(Vij, 0, sizeof(double) * (size_t)nboot * nout * d * d);
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int ro = 0; ro < nboot * nout; ++ro) {
        for (int k = 0; k < d; ++k) {
            int r = ro / nout, o = ro % nout;
            const double* base = Y + o + (size_t)nout * n * step * r;
#define BLK(b) (base + (size_t)nout * n * (b))
            const double *fA = BLK(0), *fB = BLK(1), *fAk = BLK(2 + k);
            if (k == 0) Var[ro] = var2n(fA, 2 * n, stride);
            {
                const double* s[3] = {fB, fAk, fA};
                Vi[(size_t)ro * d + k] = pw_sum(t_vi, s, 0, n, stride, NULL) / dn;
            }
            const double* s2[2] = {fA, fAk};
            double e;
            switch (ei_estimator) {
                case REF_HOMMA1996: {
                    const double* sa[1] = {fA};
                    double mA = pw_sum(t_id, sa, 0, n, stride, NULL) / dn;
                    e = pw_sum(t_csq, sa, 0, n, stride, &mA) / (dn - 1) - pw_sum(t_prod, s2, 0, n, stride, NULL) / dn + mA * mA;
                } break;
                case REF_SOBOL2007: e = pw_sum(t_sobol07, s2, 0, n, stride, NULL) / dn; break;
                case REF_JANSEN1999: e = pw_sum(t_jansen, s2, 0, n, stride, NULL) / (2 * dn); break;
                default: {
                    double q = pw_sum(t_sq2, s2, 0, n, stride, NULL), s1 = pw_sum(t_add, s2, 0, n, stride, NULL);
                    double sp = pw_sum(t_prod, s2, 0, n, stride, NULL);
                    double mh = (1.0 / dn) * pw_sum(t_addh, s2, 0, n, stride, NULL);
                    double qh = pw_sum(t_sq2h, s2, 0, n, stride, NULL);
                    double t = s1 / (2 * dn);
                    e = (q / (2 * dn) - t * t) * (1.0 - ((1.0 / dn) * sp - mh * mh) / ((1.0 / dn) * qh - mh * mh));
                } break;
            }
            Ei[(size_t)ro * d + k] = e;
            if (second_order)
                for (int j = k + 1; j < d; ++j) {
                    const double* s4[4] = {BLK(2 + d + k), BLK(2 + j), fA, fB};
                    Vij[((size_t)ro * d + k) * d + j] = pw_sum(t_vij, s4, 0, n, stride, NULL) / dn;
                }
#undef BLK
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * Normal quantile: Wichura AS241 (PPND16) + one Halley step on erfc -- Distributions.jl's
 * quantile(Normal(0,1), p) (:359) to ~1 ulp.
 * ------------------------------------------------------------------------------------- */
REF_API double ref_norm_quantile(double p) {
    static const double a[8] = {3.3871328727963666080e0, 1.3314166789178437745e+2, 1.9715909503065514427e+3, 1.3731693765509461125e+4,
                                4.5921953931549871457e+4, 6.7265770927008700853e+4, 3.3430575583588128105e+4, 2.5090809287301226727e+3};
    static const double b[8] = {1.0, 4.2313330701600911252e+1, 6.8718700749205790830e+2, 5.3941960214247511077e+3,
                                2.1213794301586595867e+4, 3.9307895800092710610e+4, 2.8729085735721942674e+4, 5.2264952788528545610e+3};
    static const double c[8] = {1.42343711074968357734e0, 4.63033784615654529590e0, 5.76949722146069140550e0, 3.64784832476320460504e0,
                                1.27045825245236838258e0, 2.41780725177450611770e-1, 2.27238449892691845833e-2, 7.74545014278341407640e-4};
    static const double dd[8] = {1.0, 2.05319162663775882187e0, 1.67638483018380384940e0, 6.89767334985100004550e-1,
                                 1.48103976427480074590e-1, 1.51986665636164571966e-2, 5.47593808499534494600e-4, 1.05075007164441684324e-9};
    static const double e[8] = {6.65790464350110377720e0, 5.46378491116411436990e0, 1.78482653991729133580e0, 2.96560571828504891230e-1,
                                2.65321895265761230930e-2, 1.24266094738807843860e-3, 2.71155556874348757815e-5, 2.01033439929228813265e-7};
    static const double f[8] = {1.0, 5.99832206555887937690e-1, 1.36929880922735805310e-1, 1.48753612908506148525e-2,
                                7.86869131145613259100e-4, 1.84631831751005468180e-5, 1.42151175831644588870e-7, 2.04426310338993978564e-15};
    double q = p - 0.5, r, x;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        double num = a[7], den = b[7];
        for (int i = 6; i >= 0; --i) { num = num * r + a[i]; den = den * r + b[i]; }
        x = q * num / den;
    } else {
        r = q < 0 ? p : 1.0 - p;
        if (r <= 0) return q < 0 ? -INFINITY : INFINITY;
        r = sqrt(-log(r));
        double num, den;
        if (r <= 5.0) {
            r -= 1.6; num = c[7]; den = dd[7];
            for (int i = 6; i >= 0; --i) { num = num * r + c[i]; den = den * r + dd[i]; }
        } else {
            r -= 5.0; num = e[7]; den = f[7];
            for (int i = 6; i >= 0; --i) { num = num * r + e[i]; den = den * r + f[i]; }
        }
        x = num / den;
        if (q < 0) x = -x;
    }
    /* Halley refinement: Phi(x) - p via erfc */
    double err = 0.5 * erfc(-x / M_SQRT2) - p;
    double u = err * sqrt(2 * M_PI) * exp(x * x / 2);
    x = x - u / (1 + x * u / 2);
    return x;
}

/* ---------------------------------------------------------------------------------------
 * Replicate post-processing (:318-404).  Result layout = the reference's arrays in column-major:
 *   S1/ST[o + nout*k], S2[k + d*j + d*d*o]; the *_ci arrays may be NULL (nboot == 1).
 * ------------------------------------------------------------------------------------- */
static void mean_ci(const double* x, int nboot, size_t stride, double z, double* mean, double* ci) {
    double s = 0.0;
    for (int r = 0; r < nboot; ++r) s += x[r * stride];
    double m = s / nboot;
    *mean = m;
    if (ci) {
        double q = 0.0;
        for (int r = 0; r < nboot; ++r) { double v = x[r * stride] - m; q += v * v; }
        *ci = z * sqrt(q / (nboot - 1)) / sqrt((double)nboot);
    }
}

REF_API void ref_finalize(const double* Var, const double* Vi, const double* Ei, const double* Vij, int nout,
                          int d, int nboot, int second_order, double conf_level, double* S1, double* S1ci,
                          double* S2, double* S2ci, double* ST, double* STci) {
    size_t ro_n = (size_t)nboot * nout;
    double* Si = (double*)malloc(sizeof(double) * ro_n * d);
    double* Ti = (double*)malloc(sizeof(double) * ro_n * d);
    double* Sij = second_order ? (double*)calloc(ro_n * d * d, sizeof(double)) : NULL;
    for (size_t ro = 0; ro < ro_n; ++ro) {
        for (int k = 0; k < d; ++k) {
            Si[ro * d + k] = Vi[ro * d + k] / Var[ro];
            Ti[ro * d + k] = Ei[ro * d + k] / Var[ro];
        }
        if (second_order)
            for (int k = 0; k < d; ++k)
                for (int j = k + 1; j < d; ++j)
                    Sij[(ro * d + k) * d + j] = (Vij[(ro * d + k) * d + j] - (Vi[ro * d + k] + Vi[ro * d + j])) / Var[ro];
    }
    double z = nboot > 1 ? ref_norm_quantile((1 + conf_level) / 2) : 0.0;
    for (int o = 0; o < nout; ++o) {
        for (int k = 0; k < d; ++k) {
            size_t off = (size_t)o * d + k, st = (size_t)nout * d;
            if (nboot > 1) {
                mean_ci(Si + off, nboot, st, z, &S1[o + (size_t)nout * k], S1ci ? &S1ci[o + (size_t)nout * k] : NULL);
                mean_ci(Ti + off, nboot, st, z, &ST[o + (size_t)nout * k], STci ? &STci[o + (size_t)nout * k] : NULL);
            } else {
                S1[o + (size_t)nout * k] = Si[off];
                ST[o + (size_t)nout * k] = Ti[off];
            }
        }
        if (second_order)
            for (int k = 0; k < d; ++k)
                for (int j = 0; j < d; ++j) {
                    size_t off = ((size_t)o * d + k) * d + j, st = (size_t)nout * d * d;
                    size_t dst = (size_t)k + (size_t)d * j + (size_t)d * d * o;
                    if (nboot > 1)
                        mean_ci(Sij + off, nboot, st, z, &S2[dst], S2ci ? &S2ci[dst] : NULL);
                    else
                        S2[dst] = Sij[off];
                }
    }
    free(Si); free(Ti); free(Sij);
}

/* ---------------------------------------------------------------------------------------
 * End to end: gsa(f, Sobol(...), A, B; batch=true)  (:110-149).  Returns the number of
 * Saltelli rows evaluated (nboot * n * step) or -1 if the scratch could not be allocated.
 * ------------------------------------------------------------------------------------- */
REF_API int64_t ref_gsa_sobol(int model, const double* params, int nparams, const double* A, const double* B,
                              int d, int64_t N, int nboot, int second_order, int ei_estimator, double conf_level,
                              double* S1, double* S1ci, double* S2, double* S2ci, double* ST, double* STci) {
    int nout = ref_model_nout(model, nparams, d);
    int64_t n = N / nboot;
    int64_t cols = ref_all_points_cols(d, N, nboot, second_order);
    double* P = (double*)malloc(sizeof(double) * (size_t)d * cols);
    double* Y = (double*)malloc(sizeof(double) * (size_t)nout * cols);
    size_t ro_n = (size_t)nboot * nout;
    double* Var = (double*)malloc(sizeof(double) * ro_n);
    double* Vi = (double*)malloc(sizeof(double) * ro_n * d);
    double* Ei = (double*)malloc(sizeof(double) * ro_n * d);
    double* Vij = second_order ? (double*)malloc(sizeof(double) * ro_n * d * d) : NULL;
    if (!P || !Y || !Var || !Vi || !Ei || (second_order && !Vij)) {
        free(P); free(Y); free(Var); free(Vi); free(Ei); free(Vij);
        return -1;
    }
    ref_all_points(A, B, d, N, nboot, second_order, P);
    ref_model_batch(model, params, nparams, P, d, cols, Y);
    ref_analysis(Y, nout, d, n, nboot, second_order, ei_estimator, Var, Vi, Ei, Vij);
    ref_finalize(Var, Vi, Ei, Vij, nout, d, nboot, second_order, conf_level, S1, S1ci, S2, S2ci, ST, STci);
    free(P); free(Y); free(Var); free(Vi); free(Ei); free(Vij);
    return cols;
}

/* analysis-only entry on a caller-provided Y (the estimator-only path, K3/K4's oracle) */
REF_API void ref_analyze_y(const double* Y, int nout, int d, int64_t n, int nboot, int second_order,
                           int ei_estimator, double conf_level, double* S1, double* S1ci, double* S2,
                           double* S2ci, double* ST, double* STci) {
    size_t ro_n = (size_t)nboot * nout;
    double* Var = (double*)malloc(sizeof(double) * ro_n);
    double* Vi = (double*)malloc(sizeof(double) * ro_n * d);
    double* Ei = (double*)malloc(sizeof(double) * ro_n * d);
    double* Vij = second_order ? (double*)malloc(sizeof(double) * ro_n * d * d) : NULL;
    ref_analysis(Y, nout, d, n, nboot, second_order, ei_estimator, Var, Vi, Ei, Vij);
    ref_finalize(Var, Vi, Ei, Vij, nout, d, nboot, second_order, conf_level, S1, S1ci, S2, S2ci, ST, STci);
    free(Var); free(Vi); free(Ei); free(Vij);
}
