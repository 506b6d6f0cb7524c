// fused_generic.cuh -- K2+K3 for ANY model and any d (runtime), thread-per-sample.
//
// For every sample s the kernel evaluates f(A_s), f(B_s), f(AB_k) (and f(BA_k) for second order)
// by swapping ONE coordinate of a thread-private copy of the sample -- the matrices the reference
// materialises in fuse_designs (src/sobol_sensitivity.jl:94-108) never exist -- and accumulates
// the per-sample estimator terms exactly as the reference forms them (difference first, then sum;
// :192, :197, :211-213).  Each term is reduced over the warp with a fixed xor butterfly and added by
// lane 0 to a warp-private shared-memory record; warps are combined in order at the end of a tile.
//
// This is the correctness / generality path (also the body a user `__device__` model is linked into);
// product-form, additive and small-d models have faster kernels (fused_product.cuh, fused_small.cuh,
// fused_linear.cuh).
#pragma once
#include "gsa_common.cuh"
#include "models.cuh"

namespace gsa {

constexpr int GENERIC_DMAX = 128;   // coordinates per sample held in thread-local memory
constexpr int GENERIC_NOMAX = 4;    // outputs per model evaluation
constexpr int GENERIC_YMAX = 256;   // d * nout bound when second order keeps all yAB / yBA

struct GenericArgs {
    const double* A;
    const double* B;
    const double* params;
    double* tile_rec;  // [ntiles][nout][nstat]
    TileGeom g;
    int d, np, nout, nstat, est, t_begin, t_end;
};

template <class Model, bool SECOND>
__device__ __forceinline__ void fused_generic_body(const GenericArgs& a) {
    extern __shared__ double smem[];  // [warps][nout*nstat]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int d = a.d, nout = a.nout, nstat = a.nstat, est = a.est;
    const int recsz = nout * nstat;
    double* acc = smem + (size_t)warp * recsz;

    double x[GENERIC_DMAX];
    double yA[GENERIC_NOMAX], yB[GENERIC_NOMAX], yk[GENERIC_NOMAX];
    double yAB[SECOND ? GENERIC_YMAX : 1], yBA[SECOND ? GENERIC_YMAX : 1];

    for (int t = a.t_begin + blockIdx.x; t < a.t_end; t += gridDim.x) {
        int64_t c0, c1;
        tile_range(a.g, t, c0, c1);
        for (int i = lane; i < recsz; i += 32) acc[i] = 0.0;
        dd sy[GENERIC_NOMAX], syy[GENERIC_NOMAX], sa[GENERIC_NOMAX], saa[GENERIC_NOMAX];
#pragma unroll
        for (int o = 0; o < GENERIC_NOMAX; ++o) sy[o] = syy[o] = sa[o] = saa[o] = dd{0.0, 0.0};
        __syncwarp();

        for (int64_t s0 = c0 + warp * 32; s0 < c1; s0 += 32 * nwarps) {
            const int64_t s = s0 + lane;
            const bool active = s < c1;
            const double* as = a.A + (size_t)(s - a.g.buf_col0) * d;
            const double* bs = a.B + (size_t)(s - a.g.buf_col0) * d;
            if (active) {
                for (int k = 0; k < d; ++k) x[k] = __ldg(bs + k);
                Model::eval(x, d, a.params, a.np, yB, nout);
                for (int k = 0; k < d; ++k) x[k] = __ldg(as + k);
                Model::eval(x, d, a.params, a.np, yA, nout);
#pragma unroll
                for (int o = 0; o < GENERIC_NOMAX; ++o)
                    if (o < nout) {
                        dd_add(sy[o], yA[o]); dd_add(sy[o], yB[o]);
                        dd_add_sq(syy[o], yA[o]); dd_add_sq(syy[o], yB[o]);
                        dd_add(sa[o], yA[o]); dd_add_sq(saa[o], yA[o]);
                    }
            }
            // x holds A_s.  AB_k: coordinate k taken from B (reference :96-99)
            for (int k = 0; k < d; ++k) {
                if (active) {
                    double keep = x[k];
                    x[k] = __ldg(bs + k);
                    Model::eval(x, d, a.params, a.np, yk, nout);
                    x[k] = keep;
                }
                for (int o = 0; o < nout; ++o) {
                    double v = 0.0, e0 = 0.0, e1 = 0.0, e2 = 0.0;
                    if (active) {
                        const double fa = yA[o], fab = yk[o];
                        v = yB[o] * (fab - fa);                                   // :192
                        if (est == GSA_EI_JANSEN1999) { double df = fa - fab; e0 = df * df; }   // :213
                        else if (est == GSA_EI_SOBOL2007) e0 = fa * (fa - fab);   // :211
                        else { e0 = fa * fab; e1 = fa * fa + fab * fab; e2 = fa + fab; }   // :207, :219-224
                        if (SECOND) yAB[k * nout + o] = fab;
                    }
                    v = warp_sum(v);
                    e0 = warp_sum(e0);
                    if (est == GSA_EI_JANON2014) { e1 = warp_sum(e1); e2 = warp_sum(e2); }
                    if (lane == 0) {
                        double* r = acc + o * nstat;
                        r[stat_v(k)] += v;
                        r[stat_e(d, 0, k)] += e0;
                        if (est == GSA_EI_JANON2014) { r[stat_e(d, 1, k)] += e1; r[stat_e(d, 2, k)] += e2; }
                    }
                }
            }
            if (SECOND) {
                // BA_k: B with coordinate k taken from A (:100-104)
                if (active)
                    for (int k = 0; k < d; ++k) x[k] = __ldg(bs + k);
                for (int k = 0; k < d; ++k) {
                    if (active) {
                        double keep = x[k];
                        x[k] = __ldg(as + k);
                        Model::eval(x, d, a.params, a.np, yk, nout);
                        x[k] = keep;
                        for (int o = 0; o < nout; ++o) yBA[k * nout + o] = yk[o];
                    }
                }
                const int pb = stat_pair_base(d, est);
                for (int o = 0; o < nout; ++o) {
                    const double fafb = active ? yA[o] * yB[o] : 0.0;
                    int p = 0;
                    for (int k = 0; k < d; ++k)
                        for (int j = k + 1; j < d; ++j, ++p) {
                            double v = active ? yBA[k * nout + o] * yAB[j * nout + o] - fafb : 0.0;   // :197
                            v = warp_sum(v);
                            if (lane == 0) acc[o * nstat + pb + p] += v;
                        }
                }
            }
        }
        // the four double-double sums: warp butterfly, then into the warp's record
        for (int o = 0; o < nout; ++o) {
            dd r0 = warp_sum_dd(sy[o]), r1 = warp_sum_dd(syy[o]), r2 = warp_sum_dd(sa[o]), r3 = warp_sum_dd(saa[o]);
            if (lane == 0) {
                double* r = acc + o * nstat;
                r[0] = r0.hi; r[1] = r0.lo; r[2] = r1.hi; r[3] = r1.lo;
                r[4] = r2.hi; r[5] = r2.lo; r[6] = r3.hi; r[7] = r3.lo;
            }
        }
        __syncthreads();
        // combine warps in order 0..nwarps-1 and write the tile record
        double* out = a.tile_rec + (size_t)t * recsz;
        for (int i = threadIdx.x; i < recsz; i += blockDim.x) {
            int e = i % nstat;
            if (e < STAT_DD) {
                if ((e & 1) == 0) {
                    dd s{0.0, 0.0};
                    for (int w = 0; w < nwarps; ++w) dd_add_dd(s, dd{smem[(size_t)w * recsz + i], smem[(size_t)w * recsz + i + 1]});
                    out[i] = s.hi;
                    out[i + 1] = s.lo;
                }
            } else {
                double s = 0.0;
                for (int w = 0; w < nwarps; ++w) s += smem[(size_t)w * recsz + i];
                out[i] = s;
            }
        }
        __syncthreads();
    }
}

template <class Model, bool SECOND>
__global__ void __launch_bounds__(256) fused_generic_kernel(GenericArgs a) {
    fused_generic_body<Model, SECOND>(a);
}

}  // namespace gsa
