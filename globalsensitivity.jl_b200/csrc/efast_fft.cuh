// efast_fft.cuh -- the spectrum of the eFAST analysis (reference src/eFAST_sensitivity.jl:185 `rfft`, :199) as a hand-written
// mixed-radix Stockham FFT, batched over the d * nout series.
//
// The reference transforms every series of `samples` values with FFTW and reads samples/16 + H of the bins.  Here:
//   * samples even: the centred real series IS a complex series of M = samples/2 points (z_n = y_2n + i y_2n+1, no copy); its
//     M-point transform is untangled into the wanted bins of the real transform by the power kernel
//     (Y_k = E_k + W^k O_k,  E_k = (Z_k + conj Z_{M-k}) / 2,  O_k = (Z_k - conj Z_{M-k}) / 2i);
//   * samples odd: the series is transformed as M = samples complex points with zero imaginary parts (first pass reads reals).
// M must factor into 4, 2, 3 and 5 (15000 -> 7500 = 4 * 3 * 5^4); other sizes keep the pruned DFT of efast.cuh, which is exact
// for any size.  One launch per radix: thread = one radix-R butterfly of one series, autosort (Stockham) addressing, so the
// result of the last pass is in natural order and no bit reversal exists.  Every twiddle is sincospi of an exactly reduced
// integer phase (t*k < M: the quotient is rounded once), not a recurrence: the error of a bin stays at a few ulp * log(M).
// The passes ping-pong between two scratch arrays that stay in L2 for the reference's sizes (15000 samples x 4 series = 0.5 MB).
#pragma once
#include <vector>
#include "efast.cuh"

namespace gsa {

// radices of an M-point transform, or false when M has a prime factor other than 2, 3, 5
inline bool efast_fft_plan(int64_t M, std::vector<int>* radices) {
    radices->clear();
    if (M < 2) return false;
    int twos = 0;
    while (M % 2 == 0) { M /= 2; ++twos; }
    for (; twos >= 2; twos -= 2) radices->push_back(4);
    if (twos) radices->push_back(2);
    while (M % 5 == 0) { M /= 5; radices->push_back(5); }
    while (M % 3 == 0) { M /= 3; radices->push_back(3); }
    return M == 1;
}

struct EfastFftArgs {
    const double* in;        // complex [nser][M] (or, first pass of an odd size, real [nser][M])
    double* out;             // complex [nser][M]
    int64_t M, Ls;           // transform length; product of the radices already applied
};

#ifdef __CUDACC__
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ double2 cmuli_neg(double2 a) { return make_double2(a.y, -a.x); }   // -i a

// forward transforms of R points (sign -1 in the exponent), in place
template <int R>
__device__ __forceinline__ void efast_dft(double2 (&v)[R]);
template <>
__device__ __forceinline__ void efast_dft<2>(double2 (&v)[2]) {
    const double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void efast_dft<4>(double2 (&v)[4]) {
    const double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]), t2 = cadd(v[1], v[3]), t3 = cmuli_neg(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}
template <>
__device__ __forceinline__ void efast_dft<3>(double2 (&v)[3]) {
    const double h = 0.86602540378443865;   // sin(2 pi / 3)
    const double2 s = cadd(v[1], v[2]), df = csub(v[1], v[2]);
    const double2 m = make_double2(fma(-0.5, s.x, v[0].x), fma(-0.5, s.y, v[0].y));
    const double2 n = make_double2(h * df.y, -h * df.x);   // -i h (b - c)
    v[0] = cadd(v[0], s);
    v[1] = cadd(m, n);
    v[2] = csub(m, n);
}
template <>
__device__ __forceinline__ void efast_dft<5>(double2 (&v)[5]) {
    const double c1 = 0.30901699437494742, c2 = -0.80901699437494742;   // cos(2 pi / 5), cos(4 pi / 5)
    const double s1 = 0.95105651629515357, s2 = 0.58778525229247313;    // sin(2 pi / 5), sin(4 pi / 5)
    const double2 a = v[0];
    const double2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    const double2 m1 = make_double2(fma(c2, t2.x, fma(c1, t1.x, a.x)), fma(c2, t2.y, fma(c1, t1.y, a.y)));
    const double2 m2 = make_double2(fma(c1, t2.x, fma(c2, t1.x, a.x)), fma(c1, t2.y, fma(c2, t1.y, a.y)));
    const double2 n1 = make_double2(fma(s2, t4.x, s1 * t3.x), fma(s2, t4.y, s1 * t3.y));
    const double2 n2 = make_double2(fma(-s1, t4.x, s2 * t3.x), fma(-s1, t4.y, s2 * t3.y));
    v[0] = cadd(a, cadd(t1, t2));
    v[1] = cadd(m1, cmuli_neg(n1));
    v[4] = csub(m1, cmuli_neg(n1));
    v[2] = cadd(m2, cmuli_neg(n2));
    v[3] = csub(m2, cmuli_neg(n2));
}

// one Stockham pass of radix R: grid (ceil(M/R / 256), series)
template <int R, bool REAL_IN>
__global__ void __launch_bounds__(256) efast_fft_pass_kernel(EfastFftArgs a) {
    const int64_t q = a.M / R, j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= q) return;
    const int64_t k = j % a.Ls;
    double2 v[R];
    if constexpr (REAL_IN) {
        const double* in = a.in + (size_t)blockIdx.y * a.M;
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = make_double2(in[j + t * q], 0.0);
    } else {
        const double2* in = reinterpret_cast<const double2*>(a.in) + (size_t)blockIdx.y * a.M;
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = in[j + t * q];
    }
    if (k > 0) {   // twiddles exp(-2 pi i t k / (Ls R)); t k < Ls R <= M
        const double inv = 2.0 / (double)(a.Ls * R);
#pragma unroll
        for (int t = 1; t < R; ++t) {
            double s, c;
            sincospi((double)(t * k) * inv, &s, &c);
            v[t] = cmul(v[t], make_double2(c, -s));
        }
    }
    efast_dft<R>(v);
    double2* out = reinterpret_cast<double2*>(a.out) + (size_t)blockIdx.y * a.M + (j - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) out[t * a.Ls] = v[t];
}

// |Y_k|^2 of the wanted bins (efast_bin) from the transform Z: grid (ceil(nbins / 256), series)
struct EfastPowerArgs {
    const double* Z;         // complex [nser][M]
    double* power;           // [nser][nbins]
    EfastGeom g;
    int64_t M;
};
__global__ void __launch_bounds__(256) efast_fft_power_kernel(EfastPowerArgs a) {
    const int b = blockIdx.x * 256 + threadIdx.x;
    if (b >= a.g.nbins) return;
    const int64_t k = efast_bin(a.g, b);
    const double2* Z = reinterpret_cast<const double2*>(a.Z) + (size_t)blockIdx.y * a.M;
    double2 y;
    if (a.g.S & 1) {
        y = Z[k];
    } else {
        const double2 zk = Z[k], zm = Z[a.M - k];                       // 1 <= k <= M - 1
        const double2 e = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));     // (Z_k + conj Z_{M-k}) / 2
        const double2 o = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));    // (Z_k - conj Z_{M-k}) / 2i
        double s, c;
        sincospi((double)k * (2.0 / (double)a.g.S), &s, &c);
        y = cadd(e, cmul(o, make_double2(c, -s)));
    }
    a.power[(size_t)blockIdx.y * a.g.nbins + b] = fma(y.x, y.x, y.y * y.y);
}
#endif

}  // namespace gsa
