// fused_product.cuh -- K2+K3 for PRODUCT-FORM models  f(x) = prod_k g_k(x_k)  (Sobol G-function), S1/ST.
//
// For such models  f(AB_k) = g_k(B_k) * prod_{j != k} g_j(A_j), so all d+2 model values of a sample cost
// O(d) instead of O(d^2): one pass of prefix products, one of suffix products (no division -- g may be 0).
// This is what makes BASELINE config 5 (d = 100, N = 2^26) HBM-bound (16*d bytes per sample, A and B read once)
// instead of FP64-bound (SURVEY 7 "hard parts").
//
// Mapping: T lanes (a power of two) cooperate on one sample; lane t owns coordinates k = t + T*i, i < CPL, so
// the T lanes of a group read T consecutive doubles of the sample's row (full 32-byte sectors, no staging
// needed) and everything else lives in registers:
//     gA[i], gB[i]           factor values of the lane's coordinates
//     others = prod of the OTHER lanes' A-products   (xor butterfly, log2 T shuffles; also yields yA, yB)
//     yAB_k  = (others * prod_{i'<i} gA[i']) * gB[i] * prod_{i'>i} gA[i']
//     V[i] += yB (yAB_k - yA)        (reference src/sobol_sensitivity.jl:192)
//     E[i] += (yA - yAB_k)^2         (:213, Jansen1999;  yA (yA - yAB_k) for Sobol2007 :211,  yA yAB_k for Homma1996 :207)
// Accumulators stay in registers for the whole tile; one butterfly + shared-memory combine per tile, in a fixed
// order (bitwise reproducible).  The AB_k matrices of fuse_designs (:94-108) never exist.
#pragma once
#include "gsa_common.cuh"
#include "tail.cuh"

namespace gsa {

// ---- the per-coordinate function g_k of a separable model ------------------------------------------------------------------
// built in: Sobol G, coefficients staged in shared memory
struct SobolGFactor {
    static constexpr bool kUser = false;
    __device__ __forceinline__ static double eval(int k, double x, const double2* scoef, const double*, int) {
        const double2 c = scoef[k];
        return fma(fabs(x - 0.5), c.x, c.y);
    }
};
// supplied by the user as relocatable device code (gsa_model_register_separable), resolved by nvJitLink:
//   product form  f(x) = prod_k gsa_user_factor(k, x_k, p, np)        additive  f(x) = sum_k gsa_user_term(k, x_k, p, np)
extern "C" __device__ double gsa_user_factor(int k, double x, const double* p, int np);
extern "C" __device__ double gsa_user_term(int k, double x, const double* p, int np);
struct UserFactor {
    static constexpr bool kUser = true;
    __device__ __forceinline__ static double eval(int k, double x, const double2*, const double* p, int np) { return gsa_user_factor(k, x, p, np); }
};
struct UserTerm {
    static constexpr bool kUser = true;
    __device__ __forceinline__ static double eval(int k, double x, const double2*, const double* p, int np) { return gsa_user_term(k, x, p, np); }
};

struct ProductArgs {
    const double* A;
    const double* B;
    const double* params; // user factors / terms: the model's parameter vector (device memory)
    int np;
    const double2* coef;  // [d] (c1_k, c0_k): g_k(x) = |x - 0.5| * c1_k + c0_k,  c1 = 4/(1+a_k), c0 = a_k/(1+a_k)
    double* tile_rec;     // [ntiles][nstat]
    TileGeom g;
    int d, nstat, t_begin, t_end;
    int est;              // GSA_EI_* (kernels with JANSEN = false: Sobol2007 or Homma1996)
    // GEN only (K1 fused in: the design is generated, never read): gsa(f, ::Sobol, p_range; samples) (:407-417)
    const uint32_t* V;    // [2*nboot*d][32] direction numbers of the 2*nboot*d-dimensional sequence
    const double2* box;   // [d] (ub-lb, lb)
    uint64_t first;       // 0-based sequence index of sample 0 of every replicate: 2^floor(log2 samples)
    int nboot, unit_box;  // unit_box: lb = 0, ub = 1 for every coordinate (the affine map is the identity)
    TailArgs tail;        // tail.counters != nullptr: stage 2, the all-reduce and the host copy run inside this launch (tail.cuh)
};

#ifndef PROD_GEN_MINBLOCKS
#define PROD_GEN_MINBLOCKS 3   // 168 registers (a few spilled words) and 12 warps/SM vs 205 registers and 8 warps/SM
#endif
constexpr int PROD_THREADS = 128;  // ~160 registers/thread at CPL = 13: three CTAs (12 warps) per SM

// shared memory of fused_product_kernel in doubles (GEN adds the box table and the transposed direction numbers)
template <int T, int CPL, bool GEN>
constexpr size_t product_smem_doubles() {
    size_t n = 2 * T * CPL + (PROD_THREADS / 32) * (2 * T * CPL + 8);
    if (GEN) n += 2 * T * CPL + (32 * (2 * T * CPL + 1) + 1) / 2;
    return n;
}

// GEN = true: the kernel GENERATES the Sobol points of the samples it evaluates (K1 fused into K2+K3): no HBM reads at all.
// A group then owns a CONTIGUOUS run of samples, so consecutive points differ by one direction number per coordinate
// (Gray code: x ^= V[j][ctz(q+1)]); the table of the replicate's 2d dimensions sits transposed in shared memory.
// JANSEN = false: the total-order term is sum yA (yA - yAB_k) (Sobol2007, :211) or sum yA yAB_k (Homma1996, :207) instead of
// sum (yA - yAB_k)^2 (:213) -- one accumulator per coordinate either way, so the same kernel serves all three (the headline
// instantiation keeps its code unchanged; Janon2014 needs three terms per coordinate and stays on the generic kernel).
// F: the per-coordinate function (SobolGFactor, UserFactor, UserTerm).  ADD = true: the model is the SUM of its terms instead of
// the product of its factors -- f(AB_k) = yA + (g_k(B_k) - g_k(A_k)) -- same mapping, same accumulators, no prefix/suffix sweeps.
// VW = 2 (even d, 16-byte aligned rows, not GEN): a lane owns PAIRS of adjacent coordinates, k = 2t + 2T*(i/2) + (i & 1), and reads
// each pair with one 16-byte load -- half the load instructions and half the L1 wavefronts of the 8-byte mapping (the T lanes of
// a group then cover one full 128-byte line per request).
template <int T, int CPL, bool GEN, bool JANSEN, class F, bool ADD, int VW = 1>
__device__ __forceinline__ void fused_product_body(const ProductArgs& a) {
    static_assert(VW == 1 || (VW == 2 && CPL % 2 == 0 && !GEN), "pair loads need an even number of slots per lane");
    auto kc = [](int t, int i) { return VW == 1 ? t + T * i : 2 * t + 2 * T * (i >> 1) + (i & 1); };   // coordinate of slot i of lane t
    const bool sobol2007 = a.est == GSA_EI_SOBOL2007;   // (only read when !JANSEN)
    constexpr int NW = PROD_THREADS / 32;
    constexpr int NG = PROD_THREADS / T;  // sample groups per CTA
    constexpr int SLOTS = T * CPL;        // padded coordinate count
    constexpr int VROW = 2 * SLOTS + 1;   // row stride of the transposed direction-number table (+1: bank skew)
    constexpr int RED = 2 * SLOTS + 8;    // per-warp reduction record: V[SLOTS], E[SLOTS], 4 dd pairs
    extern __shared__ double smem_p[];
    double2* scoef = reinterpret_cast<double2*>(smem_p);  // [SLOTS]
    double* sred = smem_p + 2 * SLOTS;                     // [NW][RED]
    double2* sbox = reinterpret_cast<double2*>(sred + NW * RED);                  // GEN: [SLOTS] (ub-lb, lb)
    uint32_t* Vs = reinterpret_cast<uint32_t*>(sred + NW * RED + 2 * SLOTS);      // GEN: [32][VROW], slot = coordinate (A) | SLOTS + coordinate (B)
    const int d = a.d;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & (T - 1);
    const int grp = threadIdx.x / T;  // sample group within the CTA
    // The sample loop below has a per-GROUP trip count (a tile's length need not be a multiple of the groups per CTA), so the
    // butterfly inside it must name only the T lanes of the group: a full-mask shuffle would wait forever for the lanes of a
    // group that has already left the loop (found the hard way: a hang on N = 100003, nboot = 3).
    const unsigned gmask = T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << (lane & ~(T - 1)));
    for (int i = threadIdx.x; i < SLOTS; i += PROD_THREADS) {
        if (!F::kUser) scoef[i] = i < d ? a.coef[i] : make_double2(0.0, 1.0);
        if (GEN) sbox[i] = i < d ? a.box[i] : make_double2(0.0, 0.0);
    }
    int vs_rep = -1;  // replicate whose direction numbers are in Vs
    __syncthreads();

    for (int tile = a.t_begin + blockIdx.x; tile < a.t_end; tile += gridDim.x) {
        int64_t c0, c1;
        tile_range(a.g, tile, c0, c1);
        double V[CPL], E[CPL];
#pragma unroll
        for (int i = 0; i < CPL; ++i) V[i] = E[i] = 0.0;
        // double-double sums of y, y^2.  T >= 2: lane t==0 carries yA in (sA1, sA2), lane t==1 carries yB there.
        // T == 1: the single lane carries yA in (sA1, sA2) and yB in (sB1, sB2).
        dd sA1{0.0, 0.0}, sA2{0.0, 0.0}, sB1{0.0, 0.0}, sB2{0.0, 0.0};

        // sample range of this group: strided over the tile (loads coalesce across groups) or, when generating, one contiguous run
        int64_t s_begin = c0 + grp, s_end = c1, s_step = NG;
        uint32_t xa[GEN ? CPL : 1], xb[GEN ? CPL : 1];
        uint64_t q = 0;
        if constexpr (GEN) {
            const int rep = a.g.rep_first + tile / a.g.tpr;
            if (rep != vs_rep) {   // (block-uniform) transpose the replicate's 2d table rows into shared memory
                __syncthreads();
                for (int idx = threadIdx.x; idx < 32 * 2 * SLOTS; idx += PROD_THREADS) {
                    const int slot = idx >> 5, b = idx & 31, k = slot < SLOTS ? slot : slot - SLOTS;
                    const int dim = (slot < SLOTS ? rep : a.nboot + rep) * d + k;
                    Vs[b * VROW + slot] = k < d ? __ldg(a.V + (size_t)32 * dim + b) : 0u;   // padded coordinates stay at x = 0
                }
                vs_rep = rep;
                __syncthreads();
            }
            const int64_t per = (c1 - c0 + NG - 1) / NG;
            s_begin = c0 + grp * per;
            s_end = s_begin + per < c1 ? s_begin + per : c1;
            s_step = 1;
            q = a.first + (uint64_t)(s_begin - (int64_t)rep * a.g.n);
#pragma unroll
            for (int i = 0; i < CPL; ++i) xa[i] = xb[i] = 0u;
            if (s_begin < s_end)
                for (uint64_t gq = q ^ (q >> 1); gq; gq &= gq - 1) {   // random-access start of the run
                    const uint32_t* row = Vs + (__ffsll((long long)gq) - 1) * VROW + t;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) { xa[i] ^= row[T * i]; xb[i] ^= row[SLOTS + T * i]; }
                }
        }
        for (int64_t s = s_begin; s < s_end; s += s_step) {
            double gA[CPL], gB[CPL], P[CPL];
            if constexpr (GEN) {
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    gA[i] = (double)xa[i] * 0x1p-32;   // exact
                    gB[i] = (double)xb[i] * 0x1p-32;
                }
                if (!a.unit_box) {
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        const double2 bx = sbox[t + T * i];
                        gA[i] = __dadd_rn(__dmul_rn(bx.x, gA[i]), bx.y);   // (ub-lb)*u + lb, rounded like the design kernel
                        gB[i] = __dadd_rn(__dmul_rn(bx.x, gB[i]), bx.y);
                    }
                }
                // advance to the next point of the run: one direction number per coordinate
                const uint32_t* row = Vs + (__ffsll((long long)~q) - 1 > 31 ? 31 : __ffsll((long long)~q) - 1) * VROW + t;
                ++q;
#pragma unroll
                for (int i = 0; i < CPL; ++i) { xa[i] ^= row[T * i]; xb[i] ^= row[SLOTS + T * i]; }
            } else {
                if constexpr (VW == 2) {
                    const double* as = a.A + (size_t)(s - a.g.buf_col0) * d;
                    const double* bs = a.B + (size_t)(s - a.g.buf_col0) * d;
#pragma unroll
                    for (int i = 0; i < CPL; i += 2) {
                        // padded pairs (k >= d; d is even) re-read coordinates 0, 1 of the row and get coef (0, 1), i.e. g = 1
                        const int k = kc(t, i), off = k < d ? k : 0;
                        const double2 va = ldg_stream2(as + off), vb = ldg_stream2(bs + off);
                        gA[i] = va.x; gA[i + 1] = va.y;
                        gB[i] = vb.x; gB[i + 1] = vb.y;
                    }
                } else {
                    const double* as = a.A + (size_t)(s - a.g.buf_col0) * d + t;
                    const double* bs = a.B + (size_t)(s - a.g.buf_col0) * d + t;
#pragma unroll
                    for (int i = 0; i < CPL; ++i) {
                        // padded coordinates (k >= d) re-read coordinate 0 of the row and get coef (0, 1), i.e. g = 1
                        const int off = (t + T * i) < d ? T * i : -t;
                        gA[i] = ldg_stream(as + off);
                        gB[i] = ldg_stream(bs + off);
                    }
                }
            }
            double LA = ADD ? 0.0 : 1.0, LB = ADD ? 0.0 : 1.0;
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int k = kc(t, i);
                if constexpr (F::kUser) {   // padded coordinates (k >= d) contribute the neutral element
                    gA[i] = k < d ? F::eval(k, gA[i], scoef, a.params, a.np) : (ADD ? 0.0 : 1.0);
                    gB[i] = k < d ? F::eval(k, gB[i], scoef, a.params, a.np) : (ADD ? 0.0 : 1.0);
                } else {
                    gA[i] = F::eval(k, gA[i], scoef, a.params, a.np);
                    gB[i] = F::eval(k, gB[i], scoef, a.params, a.np);
                }
                if constexpr (ADD) { LA += gA[i]; LB += gB[i]; }
                else { LA *= gA[i]; LB *= gB[i]; }
            }
            // butterfly over the T lanes of the group: product of the others (A) and the totals (A, B)
            double others = 1.0, yA = LA, yB = LB;
#pragma unroll
            for (int o = 1; o < T; o <<= 1) {
                const double pa = __shfl_xor_sync(gmask, yA, o);
                const double pb = __shfl_xor_sync(gmask, yB, o);
                if constexpr (ADD) { yA += pa; yB += pb; }
                else { others *= pa; yA *= pa; yB *= pb; }
            }
            if constexpr (ADD) {
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    const double df = gB[i] - gA[i];   // f(AB_k) - f(A): only term k changes
                    const double yAB = yA + df;
                    V[i] = fma(yB, df, V[i]);  // :192
                    if constexpr (JANSEN) E[i] = fma(df, df, E[i]);  // :213
                    else E[i] = fma(yA, sobol2007 ? -df : yAB, E[i]);  // :211 / :207
                }
            } else {
                // prefix products seeded with the other lanes' product, then the suffix sweep
                double run = others;
#pragma unroll
                for (int i = 0; i < CPL; ++i) {
                    P[i] = run;
                    run *= gA[i];
                }
                double suf = 1.0;
#pragma unroll
                for (int i = CPL - 1; i >= 0; --i) {
                    const double yAB = P[i] * (gB[i] * suf);
                    suf *= gA[i];
                    const double df = yAB - yA;
                    V[i] = fma(yB, df, V[i]);  // :192
                    if constexpr (JANSEN) E[i] = fma(df, df, E[i]);  // :213
                    else E[i] = fma(yA, sobol2007 ? -df : yAB, E[i]);  // :211 / :207
                }
            }
            if (T == 1) {
                dd_add(sA1, yA); dd_add_sq(sA2, yA);
                dd_add(sB1, yB); dd_add_sq(sB2, yB);
            } else if (T == 2) {
                const double y = t == 0 ? yA : yB;
                dd_add(sA1, y); dd_add_sq(sA2, y);
            } else if (t < 4) {
                // T >= 4: lanes 0..3 of the group own sum yA, sum yA^2, sum yB, sum yB^2 -- ONE double-double update per sample
                // (x or x*x with its exact product error), and the other lanes of the group do none
                const double y = (t & 2) ? yB : yA;
                const double p = (t & 1) ? y * y : y;
                const double pe = (t & 1) ? __fma_rn(y, y, -p) : 0.0;
                dd s2 = two_sum(sA1.hi, p);
                s2.lo += sA1.lo + pe;
                sA1 = quick_two_sum(s2.hi, s2.lo);
            }
        }

        // ---- tile flush: butterfly over the groups of the warp (lanes with equal t), then across warps ----
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
#pragma unroll
            for (int o = T; o < 32; o <<= 1) {
                V[i] += __shfl_xor_sync(0xffffffffu, V[i], o);
                E[i] += __shfl_xor_sync(0xffffffffu, E[i], o);
            }
        }
        // dd pairs: reduce over groups (xor offsets >= T keep t fixed)
#pragma unroll
        for (int o = (T > 1 ? T : 1); o < 32; o <<= 1) {
            dd w;
            w.hi = __shfl_xor_sync(0xffffffffu, sA1.hi, o); w.lo = __shfl_xor_sync(0xffffffffu, sA1.lo, o); dd_add_dd(sA1, w);
            if (T < 4) { w.hi = __shfl_xor_sync(0xffffffffu, sA2.hi, o); w.lo = __shfl_xor_sync(0xffffffffu, sA2.lo, o); dd_add_dd(sA2, w); }
            if (T == 1) {
                w.hi = __shfl_xor_sync(0xffffffffu, sB1.hi, o); w.lo = __shfl_xor_sync(0xffffffffu, sB1.lo, o); dd_add_dd(sB1, w);
                w.hi = __shfl_xor_sync(0xffffffffu, sB2.hi, o); w.lo = __shfl_xor_sync(0xffffffffu, sB2.lo, o); dd_add_dd(sB2, w);
            }
        }
        if (T >= 4) {  // lanes 0..3 hold sum yA, sum yA^2, sum yB, sum yB^2: bring them to lane 0
            sA2.hi = __shfl_sync(0xffffffffu, sA1.hi, 1); sA2.lo = __shfl_sync(0xffffffffu, sA1.lo, 1);
            sB1.hi = __shfl_sync(0xffffffffu, sA1.hi, 2); sB1.lo = __shfl_sync(0xffffffffu, sA1.lo, 2);
            sB2.hi = __shfl_sync(0xffffffffu, sA1.hi, 3); sB2.lo = __shfl_sync(0xffffffffu, sA1.lo, 3);
        } else if (T > 1) {  // bring lane 1's (yB) sums next to lane 0's
            sB1.hi = __shfl_sync(0xffffffffu, sA1.hi, 1); sB1.lo = __shfl_sync(0xffffffffu, sA1.lo, 1);
            sB2.hi = __shfl_sync(0xffffffffu, sA2.hi, 1); sB2.lo = __shfl_sync(0xffffffffu, sA2.lo, 1);
        }
        double* wr = sred + warp * RED;
        if (lane < T) {
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                wr[kc(t, i)] = V[i];
                wr[SLOTS + kc(t, i)] = E[i];
            }
        }
        if (lane == 0) {
            wr[2 * SLOTS + 0] = sA1.hi; wr[2 * SLOTS + 1] = sA1.lo; wr[2 * SLOTS + 2] = sA2.hi; wr[2 * SLOTS + 3] = sA2.lo;
            wr[2 * SLOTS + 4] = sB1.hi; wr[2 * SLOTS + 5] = sB1.lo; wr[2 * SLOTS + 6] = sB2.hi; wr[2 * SLOTS + 7] = sB2.lo;
        }
        __syncthreads();
        double* rec = a.tile_rec + (size_t)tile * a.nstat;
        for (int k = threadIdx.x; k < d; k += PROD_THREADS) {
            double v = 0.0, e = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                v += sred[w * RED + k];
                e += sred[w * RED + SLOTS + k];
            }
            rec[stat_v(k)] = v;
            rec[stat_e(d, 0, k)] = e;
        }
        if (threadIdx.x == 0) {
            dd a1{0, 0}, a2{0, 0}, b1{0, 0}, b2{0, 0};
            for (int w = 0; w < NW; ++w) {
                const double* q = sred + w * RED + 2 * SLOTS;
                dd_add_dd(a1, dd{q[0], q[1]}); dd_add_dd(a2, dd{q[2], q[3]});
                dd_add_dd(b1, dd{q[4], q[5]}); dd_add_dd(b2, dd{q[6], q[7]});
            }
            rec[4] = a1.hi; rec[5] = a1.lo; rec[6] = a2.hi; rec[7] = a2.lo;   // sum yA, sum yA^2
            dd_add_dd(a1, b1); dd_add_dd(a2, b2);
            rec[0] = a1.hi; rec[1] = a1.lo; rec[2] = a2.hi; rec[3] = a2.lo;   // over [yA; yB]
        }
        __syncthreads();
        if (a.tail.counters) fused_tail_tile_done(a.tail, tile);   // sign off: group -> replicate -> exchange (tail.cuh)
    }
}

template <int T, int CPL, bool GEN, bool JANSEN = true>
__global__ void __launch_bounds__(PROD_THREADS, GEN ? PROD_GEN_MINBLOCKS : 3) fused_product_kernel(const __grid_constant__ ProductArgs a) {
    fused_product_body<T, CPL, GEN, JANSEN, SobolGFactor, false>(a);
}
// pair-load variant (even d, aligned rows)
template <int T, int CPL, bool JANSEN = true>
__global__ void __launch_bounds__(PROD_THREADS, 3) fused_product_kernel_v2(const __grid_constant__ ProductArgs a) {
    fused_product_body<T, CPL, false, JANSEN, SobolGFactor, false, 2>(a);
}

}  // namespace gsa
