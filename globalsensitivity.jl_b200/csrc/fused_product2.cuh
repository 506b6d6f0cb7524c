// fused_product2.cuh -- K2+K3 for product-form models with SECOND order, d <= 32 (BASELINE config 3: Sobol G, d = 8).
//
// One coordinate per lane: the T = 2^ceil(log2 d) lanes of a group share a sample, lane t owns coordinate t.  A butterfly over
// the group gives every lane the product of the OTHER lanes' factors for both A and B, hence
//     yAB_t = othersA_t * gB_t     (A with coordinate t from B, reference src/sobol_sensitivity.jl:96-99)
//     yBA_t = othersB_t * gA_t     (B with coordinate t from A, :100-104)
// at O(1) multiplies per lane.  The pair sums V_kt = sum (yBA_k * yAB_t - yA yB), k < t (:197), are column t of the pair
// matrix: lane t receives yBA_k from lane k by shuffle and keeps at most T-1 accumulators.  Compared with the thread-per-sample
// register kernel (fused_small.cuh: 254 registers, 8 warps/SM, long dependent product chains) this one has ~60 registers,
// full occupancy and no chain longer than log2 T.
#pragma once
#include "gsa_common.cuh"

namespace gsa {

constexpr int PROD2_THREADS = 256;

struct Product2Args {
    const double* A;
    const double* B;
    const double2* coef;  // [d] (c1_k, c0_k): g_k(x) = |x - 0.5| * c1_k + c0_k
    double* tile_rec;     // [ntiles][nstat]
    TileGeom g;
    int d, nstat, est, t_begin, t_end;
};

template <int T>
__global__ void __launch_bounds__(PROD2_THREADS) fused_product2_kernel(Product2Args a) {
    constexpr int NW = PROD2_THREADS / 32, NG = PROD2_THREADS / T;
    static_assert(T >= 2 && T <= 32, "one coordinate per lane: 2 <= T <= 32");
    constexpr int NP = T - 1;                // pair accumulators per lane (k < t)
    constexpr int RED = T * (NP + 2) + 8;    // per-warp record: V[T], E[T], P[T][NP], 4 dd pairs
    extern __shared__ double sred[];          // [NW][RED]
    const int d = a.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & (T - 1), grp = threadIdx.x / T;
    const unsigned gmask = T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << (lane & ~(T - 1)));   // see fused_product.cuh
    const bool live = t < d;
    const double2 cf = live ? __ldg(a.coef + t) : make_double2(0.0, 1.0);   // padded lanes: g = 1

    for (int tile = a.t_begin + blockIdx.x; tile < a.t_end; tile += gridDim.x) {
        int64_t c0, c1;
        tile_range(a.g, tile, c0, c1);
        double V = 0.0, E = 0.0, P[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) P[k] = 0.0;
        ShiftSums acc;   // lane 0: yA, lane 1: yB (as in fused_product.cuh); shifted + compensated, un-shifted at the flush

        for (int64_t s = c0 + grp; s < c1; s += NG) {
            const size_t row = (size_t)(s - a.g.buf_col0) * d;
            const double xa = live ? ldg_stream(a.A + row + t) : 0.5;
            const double xb = live ? ldg_stream(a.B + row + t) : 0.5;
            const double ga = fma(fabs(xa - 0.5), cf.x, cf.y), gb = fma(fabs(xb - 0.5), cf.x, cf.y);
            double yA = ga, yB = gb, oA = 1.0, oB = 1.0;
#pragma unroll
            for (int o = 1; o < T; o <<= 1) {
                const double pa = __shfl_xor_sync(gmask, yA, o), pb = __shfl_xor_sync(gmask, yB, o);
                oA *= pa; yA *= pa;
                oB *= pb; yB *= pb;
            }
            const double yab = oA * gb, yba = oB * ga;
            const double df = yab - yA;
            V = fma(yB, df, V);    // :192
            E = fma(df, df, E);    // :213
            const double fafb = yA * yB;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                const double v = __shfl_sync(gmask, yba, k, T);      // yBA_k of this sample
                if (k < t) P[k] += v * yab - fafb;                   // :197
            }
            acc.add(t == 0 ? yA : yB);   // lanes >= 2 accumulate a copy that is never used
        }
        dd s1, s2;
        acc.to_dd(s1, s2);
        // ---- flush: over the groups of the warp (xor offsets >= T keep t), then over warps in order ----
#pragma unroll
        for (int o = T; o < 32; o <<= 1) {
            V += __shfl_xor_sync(0xffffffffu, V, o);
            E += __shfl_xor_sync(0xffffffffu, E, o);
#pragma unroll
            for (int k = 0; k < NP; ++k) P[k] += __shfl_xor_sync(0xffffffffu, P[k], o);
            dd w;
            w.hi = __shfl_xor_sync(0xffffffffu, s1.hi, o); w.lo = __shfl_xor_sync(0xffffffffu, s1.lo, o); dd_add_dd(s1, w);
            w.hi = __shfl_xor_sync(0xffffffffu, s2.hi, o); w.lo = __shfl_xor_sync(0xffffffffu, s2.lo, o); dd_add_dd(s2, w);
        }
        dd sB1, sB2;
        sB1.hi = __shfl_sync(0xffffffffu, s1.hi, 1); sB1.lo = __shfl_sync(0xffffffffu, s1.lo, 1);
        sB2.hi = __shfl_sync(0xffffffffu, s2.hi, 1); sB2.lo = __shfl_sync(0xffffffffu, s2.lo, 1);
        double* wr = sred + warp * RED;
        if (lane < T) {
            wr[t] = V;
            wr[T + t] = E;
#pragma unroll
            for (int k = 0; k < NP; ++k) wr[2 * T + t * NP + k] = P[k];
        }
        if (lane == 0) {
            double* q = wr + T * (NP + 2);
            q[0] = s1.hi; q[1] = s1.lo; q[2] = s2.hi; q[3] = s2.lo;
            q[4] = sB1.hi; q[5] = sB1.lo; q[6] = sB2.hi; q[7] = sB2.lo;
        }
        __syncthreads();
        double* rec = a.tile_rec + (size_t)tile * a.nstat;
        const int pbase = stat_pair_base(d, a.est);
        for (int i = threadIdx.x; i < T * (NP + 2); i += PROD2_THREADS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += sred[w * RED + i];
            if (i < T) { if (i < d) rec[stat_v(i)] = s; }
            else if (i < 2 * T) { if (i - T < d) rec[stat_e(d, 0, i - T)] = s; }
            else {
                const int tt = (i - 2 * T) / NP, k = (i - 2 * T) % NP;
                if (k < tt && tt < d) rec[pbase + pair_index(d, k, tt)] = s;
            }
        }
        if (threadIdx.x == 0) {
            dd a1{0, 0}, a2{0, 0}, b1{0, 0}, b2{0, 0};
            for (int w = 0; w < NW; ++w) {
                const double* q = sred + w * RED + T * (NP + 2);
                dd_add_dd(a1, dd{q[0], q[1]}); dd_add_dd(a2, dd{q[2], q[3]});
                dd_add_dd(b1, dd{q[4], q[5]}); dd_add_dd(b2, dd{q[6], q[7]});
            }
            rec[4] = a1.hi; rec[5] = a1.lo; rec[6] = a2.hi; rec[7] = a2.lo;   // sum yA, sum yA^2
            dd_add_dd(a1, b1); dd_add_dd(a2, b2);
            rec[0] = a1.hi; rec[1] = a1.lo; rec[2] = a2.hi; rec[3] = a2.lo;   // over [yA; yB]
        }
        __syncthreads();
    }
}

template <int T>
constexpr size_t product2_smem_bytes() { return sizeof(double) * (PROD2_THREADS / 32) * (T * (T - 1 + 2) + 8); }

}  // namespace gsa
