// finalize.cpp -- host epilogue of the Sobol path: per-replicate sums -> SobolResult.
// Follows reference src/sobol_sensitivity.jl:318-404 (S2 normalisation, S1/ST, mean over replicates,
// CI half-width z*std/sqrt(nboot)); the variance of [fA; fB] (:183) is recovered from double-double
// sums of y and y^2, which equals the reference's two-pass value to the last bits.
#include <cmath>
#include <cstring>
#include <vector>

#include "gsa_common.cuh"
#include "gsa_internal.h"

namespace gsa {

namespace {

inline dd two_prod(double a, double b) {
    double p = a * b;
    return {p, __builtin_fma(a, b, -p)};
}
inline dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
inline dd dd_div_d(dd a, double b) {  // a / b, b exact
    double q1 = a.hi / b;
    dd p = two_prod(q1, b);
    double r = ((a.hi - p.hi) - p.lo) + a.lo;
    double q2 = r / b;
    return quick_two_sum(q1, q2);
}
inline dd dd_sub(dd a, dd b) {
    dd nb{-b.hi, -b.lo};
    dd_add_dd(a, nb);
    return a;
}
// corrected variance of m values from dd sums: (Q - S^2/m) / (m - 1)
inline double var_from_sums(dd S, dd Q, double m) {
    dd s2 = dd_div_d(dd_mul(S, S), m);
    dd num = dd_sub(Q, s2);
    dd v = dd_div_d(num, m - 1.0);
    return v.hi + v.lo;
}

// Phi^-1: Acklam's rational start + Halley steps on erfc (converges to < 1 ulp)
double norm_quantile(double p) {
    if (!(p > 0.0 && p < 1.0)) return p == 0.0 ? -INFINITY : (p == 1.0 ? INFINITY : NAN);
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double dcf[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
    double x;
    if (p < 0.02425) {
        double q = std::sqrt(-2 * std::log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((dcf[0] * q + dcf[1]) * q + dcf[2]) * q + dcf[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        double q = p - 0.5, r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        double q = std::sqrt(-2 * std::log1p(-p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((dcf[0] * q + dcf[1]) * q + dcf[2]) * q + dcf[3]) * q + 1);
    }
    for (int it = 0; it < 3; ++it) {
        // e = Phi(x) - p, evaluated on the tail that keeps relative accuracy
        double e = x < 0 ? 0.5 * std::erfc(-x / M_SQRT2) - p : (1.0 - p) - 0.5 * std::erfc(x / M_SQRT2);
        double u = e * std::sqrt(2 * M_PI) * std::exp(x * x / 2);
        x -= u / (1 + x * u / 2);
    }
    return x;
}

// mean over replicates and CI half-width (:356-375)
inline void mean_ci(const double* x, int nboot, size_t stride, double z, double* mean, double* ci) {
    double s = 0.0;
    for (int r = 0; r < nboot; ++r) s += x[r * stride];
    double m = s / nboot;
    if (mean) *mean = m;
    if (ci) {
        double q = 0.0;
        for (int r = 0; r < nboot; ++r) {
            double v = x[r * stride] - m;
            q += v * v;
        }
        *ci = z * std::sqrt(q / (nboot - 1)) / std::sqrt((double)nboot);
    }
}

}  // namespace

double host_norm_quantile(double p) { return norm_quantile(p); }

int finalize_host(const gsa_sobol_opts& o, const double* P, int nout, int d, int64_t n, gsa_sobol_result* out) {
    const int nboot = o.nboot, est = o.ei_estimator;
    const bool second = o.second_order != 0;
    const int nstat = stat_count(d, est, second);
    const size_t ro_n = (size_t)nboot * nout;
    const double dn = (double)n;
    std::vector<double> Si(ro_n * d), Ti(ro_n * d), Sij;
    if (second) Sij.assign(ro_n * d * d, 0.0);
    std::vector<double> Vi(d);
    const int pb = stat_pair_base(d, est);
    for (size_t ro = 0; ro < ro_n; ++ro) {
        const double* r = P + ro * nstat;
        const double var = var_from_sums(dd{r[0], r[1]}, dd{r[2], r[3]}, 2.0 * dn);       // :183
        for (int k = 0; k < d; ++k) {
            Vi[k] = r[stat_v(k)] / dn;                                                    // :192
            double e;
            switch (est) {
                case GSA_EI_HOMMA1996: {                                                  // :203-209
                    double varA = var_from_sums(dd{r[4], r[5]}, dd{r[6], r[7]}, dn);
                    double mA = (r[4] + r[5]) / dn;
                    e = varA - r[stat_e(d, 0, k)] / dn + mA * mA;
                } break;
                case GSA_EI_SOBOL2007: e = r[stat_e(d, 0, k)] / dn; break;                // :211
                case GSA_EI_JANSEN1999: e = r[stat_e(d, 0, k)] / (2 * dn); break;         // :213
                default: {                                                                // :215-235
                    double sp = r[stat_e(d, 0, k)], s2 = r[stat_e(d, 1, k)], s1 = r[stat_e(d, 2, k)];
                    double mh = (1.0 / dn) * (s1 / 2), t = s1 / (2 * dn);
                    e = (s2 / (2 * dn) - t * t) * (1.0 - ((1.0 / dn) * sp - mh * mh) / ((1.0 / dn) * (s2 / 2) - mh * mh));
                } break;
            }
            Si[ro * d + k] = Vi[k] / var;                                                 // :349
            Ti[ro * d + k] = e / var;                                                     // :350
        }
        if (second) {
            int p = 0;
            for (int k = 0; k < d; ++k)
                for (int j = k + 1; j < d; ++j, ++p)
                    Sij[(ro * d + k) * d + j] = (r[pb + p] / dn - (Vi[k] + Vi[j])) / var;     // :197, :325-328
        }
    }
    const double z = nboot > 1 ? norm_quantile((1 + o.conf_level) / 2) : 0.0;            // :359
    for (int oi = 0; oi < nout; ++oi) {
        for (int k = 0; k < d; ++k) {
            const size_t off = (size_t)oi * d + k, st = (size_t)nout * d, dst = oi + (size_t)nout * k;
            if (nboot > 1) {
                double m, c;
                mean_ci(&Si[off], nboot, st, z, &m, &c);
                if (out->S1) out->S1[dst] = m;
                if (out->S1_conf_int) out->S1_conf_int[dst] = c;
                mean_ci(&Ti[off], nboot, st, z, &m, &c);
                if (out->ST) out->ST[dst] = m;
                if (out->ST_conf_int) out->ST_conf_int[dst] = c;
            } else {
                if (out->S1) out->S1[dst] = Si[off];
                if (out->ST) out->ST[dst] = Ti[off];
            }
        }
        if (second)
            for (int k = 0; k < d; ++k)
                for (int j = 0; j < d; ++j) {
                    const size_t off = ((size_t)oi * d + k) * d + j, st = (size_t)nout * d * d;
                    const size_t dst = (size_t)k + (size_t)d * j + (size_t)d * d * oi;
                    if (nboot > 1) {
                        double m, c;
                        mean_ci(&Sij[off], nboot, st, z, &m, &c);
                        if (out->S2) out->S2[dst] = m;
                        if (out->S2_conf_int) out->S2_conf_int[dst] = c;
                    } else if (out->S2) {
                        out->S2[dst] = Sij[off];
                    }
                }
    }
    return GSA_OK;
}

}  // namespace gsa
