// efast.cuh -- the eFAST method of the reference (src/eFAST_sensitivity.jl:75-213), SURVEY 8(f) N4: the sibling variance-based
// method with the same batch / multi-output conventions and result fields S1 / ST.
//
//   design    (:98-135)  d search curves of `samples` points each; in curve i parameter i carries the frequency omega[0], the others
//                        the complementary set; x_j = lb_j + (ub_j - lb_j) (1/2 + asin(sinpi(omega_j s + phi_i)) / pi), s = 2k/samples.
//                        Closed form: generated in registers by the evaluation kernel, never stored (gsa_efast_design stores it for
//                        callers that evaluate their own closure).  phi (the reference's 2 rand(rng), :111) is an INPUT.
//   model               one evaluation per point (samples * d points), thread per point.
//   analysis  (:167-213) per curve and output: the power spectrum |F_j|^2 of the samples, j = 1 .. samples/2 - 1;
//                        S1 = sum_{h <= H} |F_{h omega0}|^2 / z,   ST = 1 - sum_{j <= omega0/2} |F_j|^2 / z,   z = sum_j |F_j|^2.
//
// The analysis needs omega0/2 + H of the samples/2 bins (samples/16 for H = 4) and the TOTAL power.  The total comes from
// Parseval's identity on the centred series -- exact, one pass.  The wanted bins come from a hand-written mixed-radix Stockham
// FFT (efast_fft.cuh) when the size factors into 2, 3 and 5 -- 15000 samples x 4 series: 0.03 ms -- and otherwise (primes such
// as 4099, 15003 = 3^2 * 1667) from the PRUNED DFT below, exact for any size: one thread per bin, the series staged through
// shared memory and broadcast, the twiddle advanced by a rotation that is re-seeded from the exact integer phase
// (j*s mod samples -> sincospi) every EF_RESEED samples, so its error stays below 1e-14; O(samples^2 / 16) per series
// (0.3 ms at 15000 samples, 12 ms at 2^17 x 20 series, where the FFT takes 0.07 ms: profiles/r2_efast_probe.jsonl).
#pragma once
#include "gsa_common.cuh"
#include "models.cuh"

namespace gsa {

constexpr int EF_DMAX = 128;       // coordinates held in thread-local memory by the evaluation kernel
constexpr int EF_NOMAX = 4;
constexpr int EF_THREADS = 128;    // bins per CTA of the spectrum kernel
constexpr int EF_TILE = 512;       // samples staged per step
constexpr int EF_RESEED = 32;      // samples between exact re-seeds of the twiddle

struct EfastGeom {
    int d, nout, H;          // parameters, outputs, harmonics
    int64_t S;               // samples per curve
    int64_t omega0;          // frequency of the parameter of interest
    int nlow, nbins;         // omega0 / 2 low bins; nbins = nlow + H (+ 1: bin (S-1)/2 when S is odd)
};

// frequency carried by parameter j in curve i (:100-108); omega[0..d) in device memory
__device__ __forceinline__ int64_t efast_freq(const int64_t* omega, int i, int j) { return j == i ? omega[0] : (j < i ? omega[j + 1] : omega[j]); }

// the point of curve i at sample k: x[j], j < d   (every operation rounded separately, like the reference's broadcast)
template <class X>
__device__ __forceinline__ void efast_point(X&& x, int d, const double* lb, const double* ub, const int64_t* omega, const double* phi, int i, int64_t k,
                                            int64_t S) {
    const double s = __dmul_rn(2.0 / (double)S, (double)k);               // (2 / samples) * k
    for (int j = 0; j < d; ++j) {
        const double lo = lb[j], hi = ub[j];
        if (lo == hi) { x[j] = lo; continue; }
        const double arg = __dadd_rn(__dmul_rn((double)efast_freq(omega, i, j), s), phi[i]);
        const double q = __dadd_rn(0.5, __dmul_rn(0.3183098861837907 /* 1/pi */, asin(sinpi(arg))));
        x[j] = __dadd_rn(lo, __dmul_rn(q, __dadd_rn(hi, -lo)));           // quantile(Uniform(lo, hi), q) = lo + q (hi - lo)
    }
}

struct EfastDesignArgs {
    const double* lb; const double* ub; const double* phi; const int64_t* omega;
    double* P;               // d x (S*d)
    int d; int64_t S;
};
__global__ void __launch_bounds__(256) efast_design_kernel(EfastDesignArgs a) {
    const int64_t total = a.S * a.d;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(c / a.S);
        const int64_t k = c - (int64_t)i * a.S;
        double* out = a.P + (size_t)c * a.d;
        efast_point(out, a.d, a.lb, a.ub, a.omega, a.phi, i, k, a.S);
    }
}

struct EfastEvalArgs {
    const double* lb; const double* ub; const double* phi; const int64_t* omega;
    const double* params;
    double* Y;               // nout x (S*d), column c = curve i, sample k
    int d, np, nout; int64_t S;
};
template <class Model>
__device__ __forceinline__ void efast_eval_body(const EfastEvalArgs& a) {
    double x[EF_DMAX], y[EF_NOMAX];
    const int64_t total = a.S * a.d;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(c / a.S);
        efast_point(x, a.d, a.lb, a.ub, a.omega, a.phi, i, c - (int64_t)i * a.S, a.S);
        Model::eval(x, a.d, a.params, a.np, y, a.nout);
        for (int o = 0; o < a.nout; ++o) a.Y[o + (size_t)a.nout * c] = y[o];
    }
}
template <class Model>
__global__ void __launch_bounds__(128) efast_eval_kernel(EfastEvalArgs a) { efast_eval_body<Model>(a); }

// ---- analysis -----------------------------------------------------------------------------------------------------------------
struct EfastSpecArgs {
    const double* Y;         // nout x (S*d)
    double* C;               // [d*nout][S] centred series (scratch)
    double* mom;             // [d*nout][4]: S * sum yc^2 (hi, lo), alternating sum, unused
    double* power;           // [d*nout][nbins]
    double* S1; double* ST;  // nout x d, column-major (device)
    EfastGeom g;
};

// one CTA per series (curve i, output o): mean (double-double), centred copy, sum of squares, alternating sum
__global__ void __launch_bounds__(256) efast_center_kernel(EfastSpecArgs a) {
    __shared__ double sh[2][256];
    __shared__ double s_mean;
    const int ser = blockIdx.x, i = ser / a.g.nout, o = ser % a.g.nout;
    const int64_t S = a.g.S;
    const double* y = a.Y + o + (size_t)a.g.nout * ((size_t)i * S);
    dd acc{0.0, 0.0};
    for (int64_t k = threadIdx.x; k < S; k += 256) dd_add(acc, y[(size_t)a.g.nout * k]);
    sh[0][threadIdx.x] = acc.hi; sh[1][threadIdx.x] = acc.lo;
    __syncthreads();
    if (threadIdx.x == 0) {
        dd t{0.0, 0.0};
        for (int w = 0; w < 256; ++w) dd_add_dd(t, dd{sh[0][w], sh[1][w]});
        s_mean = (t.hi + t.lo) / (double)S;
    }
    __syncthreads();
    const double m = s_mean;
    dd q{0.0, 0.0};
    double alt = 0.0;
    double* c = a.C + (size_t)ser * S;
    for (int64_t k = threadIdx.x; k < S; k += 256) {
        const double v = y[(size_t)a.g.nout * k] - m;
        c[k] = v;
        dd_add_sq(q, v);
        alt += (k & 1) ? -v : v;
    }
    __syncthreads();
    sh[0][threadIdx.x] = q.hi; sh[1][threadIdx.x] = q.lo;
    __syncthreads();
    dd tq{0.0, 0.0};
    if (threadIdx.x == 0)
        for (int w = 0; w < 256; ++w) dd_add_dd(tq, dd{sh[0][w], sh[1][w]});
    __syncthreads();
    sh[0][threadIdx.x] = alt;
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0;
        for (int w = 0; w < 256; ++w) ta += sh[0][w];
        a.mom[4 * ser + 0] = tq.hi; a.mom[4 * ser + 1] = tq.lo; a.mom[4 * ser + 2] = ta; a.mom[4 * ser + 3] = m;
    }
}

// bin index of slot b: the low bins 1..nlow, then the harmonics h*omega0, then (S odd) bin (S-1)/2
__device__ __forceinline__ int64_t efast_bin(const EfastGeom& g, int b) {
    if (b < g.nlow) return b + 1;
    if (b < g.nlow + g.H) return (int64_t)(b - g.nlow + 1) * g.omega0;
    return (g.S - 1) / 2;
}

// grid (ceil(nbins / EF_THREADS), series): thread = one bin of one series; pruned DFT with re-seeded rotation
__global__ void __launch_bounds__(EF_THREADS) efast_spectrum_kernel(EfastSpecArgs a) {
    __shared__ double tile[EF_TILE];
    const int ser = blockIdx.y, b = blockIdx.x * EF_THREADS + threadIdx.x;
    const int64_t S = a.g.S;
    const bool live = b < a.g.nbins;
    const int64_t j = live ? efast_bin(a.g, b) : 0;
    const double* c = a.C + (size_t)ser * S;
    const double invS2 = 2.0 / (double)S;
    double rc, rs;                                    // rotation by one sample: exp(-2 pi i j / S)
    sincospi((double)j * invS2, &rs, &rc);            // j < S/2 < 2^52: the product is rounded once
    double re = 0.0, im = 0.0;
    for (int64_t s0 = 0; s0 < S; s0 += EF_TILE) {
        const int cnt = (int)min((int64_t)EF_TILE, S - s0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += EF_THREADS) tile[t] = c[s0 + t];
        __syncthreads();
        for (int t0 = 0; t0 < cnt; t0 += EF_RESEED) {
            // exact phase of sample s0 + t0: (j * s) mod S, then sincospi of 2 r / S
            const unsigned long long r = ((unsigned long long)j * (unsigned long long)(s0 + t0)) % (unsigned long long)S;
            double wc, ws;
            sincospi((double)r * invS2, &ws, &wc);
            const int e = min(cnt, t0 + EF_RESEED);
            for (int t = t0; t < e; ++t) {
                const double v = tile[t];
                re = fma(v, wc, re);
                im = fma(-v, ws, im);
                const double nc = fma(wc, rc, -ws * rs), nsn = fma(wc, rs, ws * rc);   // advance by one sample
                wc = nc; ws = nsn;
            }
        }
    }
    if (live) a.power[(size_t)ser * a.g.nbins + b] = fma(re, re, im * im);
}

// one CTA per series: z from Parseval, the two ratios (fixed-order sums)
__global__ void __launch_bounds__(256) efast_indices_kernel(EfastSpecArgs a) {
    __shared__ double sh[256];
    const int ser = blockIdx.x, i = ser / a.g.nout, o = ser % a.g.nout;
    const double* p = a.power + (size_t)ser * a.g.nbins;
    double low = 0.0;
    for (int b = threadIdx.x; b < a.g.nlow; b += 256) low += p[b];
    sh[threadIdx.x] = low;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double S = (double)a.g.S;
        // sum_{j=0}^{S-1} |F_j|^2 = S sum yc^2, F_0 = 0 (centred); the bins j and S - j carry the same power
        dd sq{a.mom[4 * ser], a.mom[4 * ser + 1]};
        const double tot = S * (sq.hi + sq.lo);
        double z;
        if ((a.g.S & 1) == 0) { const double alt = a.mom[4 * ser + 2]; z = 0.5 * (tot - alt * alt); }   // minus the Nyquist bin
        else z = 0.5 * tot - p[a.g.nlow + a.g.H];                                                        // the reference stops at S/2 - 1 (:186)
        double first = 0.0;
        for (int h = 0; h < a.g.H; ++h) first += p[a.g.nlow + h];
        a.S1[o + (size_t)a.g.nout * i] = first / z;            // (:188, :202)
        a.ST[o + (size_t)a.g.nout * i] = 1.0 - sh[0] / z;      // (:189, :203)
    }
}

}  // namespace gsa
