// gsa_common.cuh -- shared device/host definitions of libgsa_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gsa_cuda.h"

namespace gsa {

// ------------------------------------------------------------------------------------------------
// Sufficient-statistics record, per (replicate, output).  All entries are plain sums over samples,
// hence additive across tiles, CTAs and GPUs.  Entries 0..7 are four double-double (hi, lo) pairs.
//   0,1  sum y        over the 2n values [yA; yB]      -> Var  (reference :182-183 / :240-241)
//   2,3  sum y^2      over [yA; yB]
//   4,5  sum yA                                         -> Homma1996 (:203-209)
//   6,7  sum yA^2
//   8 + k                    sum yB (yAB_k - yA)        -> V_k  (:192)
//   8 + d + e*d + k          estimator term e of E_k    (:202-236), e < ne(est)
//   8 + d + ne*d + pair(k,j) sum (yBA_k yAB_j - yA yB)  -> V_kj (:197), k<j, k-major order
// ------------------------------------------------------------------------------------------------
constexpr int STAT_DD = 8;

__host__ __device__ inline int est_nterms(int est) { return est == GSA_EI_JANON2014 ? 3 : 1; }
__host__ __device__ inline int stat_count(int d, int est, int second) {
    return STAT_DD + d + est_nterms(est) * d + (second ? d * (d - 1) / 2 : 0);
}
__host__ __device__ inline int stat_v(int k) { return STAT_DD + k; }
__host__ __device__ inline int stat_e(int d, int e, int k) { return STAT_DD + d + e * d + k; }
__host__ __device__ inline int stat_pair_base(int d, int est) { return STAT_DD + d + est_nterms(est) * d; }
// index of pair (k<j) in k-major order
__host__ __device__ inline int pair_index(int d, int k, int j) { return k * d - k * (k + 1) / 2 + (j - k - 1); }

// ------------------------------------------------------------------------------------------------
// Tile geometry.  Replicate r owns global columns [r*n, (r+1)*n) (reference :116-127).  Every
// replicate is cut into `tpr` tiles of `ts` samples; tile t of the shard maps to replicate
// rep_first + t / tpr.  A tile never straddles a replicate, so one record per (tile, output) is
// written and stage 2 sums the tiles of a replicate in a fixed order (deterministic results).
// ------------------------------------------------------------------------------------------------
struct TileGeom {
    int64_t n;          // columns per replicate
    int64_t col_lo;     // first global column held by this shard (clip)
    int64_t col_hi;     // one past the last valid global column of this shard (<= nboot*n)
    int64_t buf_col0;   // global column index of element 0 of the A/B buffers passed to the kernel
    int ts;             // samples per tile
    int tpr;            // tiles per replicate = ceil(n / ts)
    int rep_first;      // first replicate touched by the shard
    int ntiles;         // nrep_local * tpr
};

__host__ __device__ inline void tile_range(const TileGeom& g, int t, int64_t& c0, int64_t& c1) {
    int r = g.rep_first + t / g.tpr;
    int j = t % g.tpr;
    c0 = (int64_t)r * g.n + (int64_t)j * g.ts;
    c1 = c0 + g.ts;
    int64_t rend = (int64_t)(r + 1) * g.n;
    if (c1 > rend) c1 = rend;
    if (c0 < g.col_lo) c0 = g.col_lo;
    if (c1 > g.col_hi) c1 = g.col_hi;
    if (c1 < c0) c1 = c0;
}

// ------------------------------------------------------------------------------------------------
// double-double helpers (error-free transforms; need IEEE fp64 add/mul/fma, no fast-math)
// ------------------------------------------------------------------------------------------------
struct dd {
    double hi, lo;
};

__host__ __device__ inline dd two_sum(double a, double b) {
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return {s, e};
}
__host__ __device__ inline dd quick_two_sum(double a, double b) {
    double s = a + b;
    return {s, b - (s - a)};
}
// acc += x   (x a plain double)
__host__ __device__ inline void dd_add(dd& acc, double x) {
    dd s = two_sum(acc.hi, x);
    s.lo += acc.lo;
    acc = quick_two_sum(s.hi, s.lo);
}
// acc += x*x exactly (product error captured with an FMA)
__host__ __device__ inline void dd_add_sq(dd& acc, double x) {
    double p = x * x;
#ifdef __CUDA_ARCH__
    double e = __fma_rn(x, x, -p);
#else
    double e = __builtin_fma(x, x, -p);
#endif
    dd s = two_sum(acc.hi, p);
    s.lo += acc.lo + e;
    acc = quick_two_sum(s.hi, s.lo);
}
// acc += b  (both double-double)
__host__ __device__ inline void dd_add_dd(dd& acc, dd b) {
    dd s = two_sum(acc.hi, b.hi);
    dd t = two_sum(acc.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    acc = quick_two_sum(s.hi, s.lo);
}

// ------------------------------------------------------------------------------------------------
// Shifted, compensated running sums of y and y^2 for ONE thread: sum (y - c) and sum (y - c)^2 with c = the thread's first
// value, each Kahan-compensated (10 flops per value instead of the 24 of dd_add + dd_add_sq).  Because the sums are taken
// about a value of the data itself they do not suffer from mean^2 >> var; to_dd() then un-shifts them EXACTLY (double-double)
// into sum y and sum y^2, so everything downstream (tile / replicate / rank reductions, finalize) is unchanged.
// ------------------------------------------------------------------------------------------------
struct ShiftSums {
    double c = 0.0, s = 0.0, sc = 0.0, q = 0.0, qc = 0.0;
    int n = 0;
    __host__ __device__ inline void add(double y) {
        c = n == 0 ? y : c;
        ++n;
        const double t = y - c;
        double v = t - sc, w = s + v;
        sc = (w - s) - v;
        s = w;
        v = t * t - qc;
        w = q + v;
        qc = (w - q) - v;
        q = w;
    }
    // S = sum y = s' + n c ;  Q = sum y^2 = q' + 2 c s' + n c^2     (s' = s - sc, q' = q - qc)
    __host__ __device__ inline void to_dd(dd& S, dd& Q) const {
#ifdef __CUDA_ARCH__
#define GSA_FMA(a, b, c_) __fma_rn(a, b, c_)
#else
#define GSA_FMA(a, b, c_) __builtin_fma(a, b, c_)
#endif
        const dd sp = quick_two_sum(s, -sc), qp = quick_two_sum(q, -qc);
        const double dn = (double)n;
        double p = dn * c, e = GSA_FMA(dn, c, -p);          // n*c exactly
        S = sp;
        dd_add_dd(S, dd{p, e});
        double c2 = c * c, c2e = GSA_FMA(c, c, -c2);        // c^2 exactly
        double m1 = dn * c2, m1e = GSA_FMA(dn, c2, -m1);    // n*c^2 (hi part) ...
        m1e = GSA_FMA(dn, c2e, m1e);                        // ... plus n * (low part of c^2)
        double m2 = 2.0 * c * sp.hi, m2e = GSA_FMA(2.0 * c, sp.hi, -m2);
        m2e = GSA_FMA(2.0 * c, sp.lo, m2e);
        Q = qp;
        dd_add_dd(Q, quick_two_sum(m1, m1e));
        dd_add_dd(Q, quick_two_sum(m2, m2e));
#undef GSA_FMA
    }
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// warp / block reductions in a FIXED order (xor butterflies), for run-to-run reproducibility
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ dd warp_sum_dd(dd v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dd w;
        w.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
        w.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
        // order the operands by lane so that both partners compute bit-identical results
        if ((threadIdx.x & o) != 0) {
            dd t = v; v = w; w = t;
        }
        dd_add_dd(v, w);
    }
    return v;
}

// streaming read-only loads that do not pollute L1.  Deliberately NOT `asm volatile`: the data is immutable for the
// lifetime of the kernel, and the scheduler must be free to batch many of these ahead of their uses (memory-level
// parallelism is what a streaming kernel lives on).
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
    double2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double* p) {
    double r;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
#endif  // __CUDACC__

}  // namespace gsa
