// fused_generic.cuh -- K2+K3 for ANY model and any d (runtime), thread-per-sample.
//
// For every sample s the kernel evaluates f(A_s), f(B_s), f(AB_k) (and f(BA_k) for second order) by swapping ONE coordinate
// of a thread-private copy of the sample -- the matrices the reference materialises in fuse_designs
// (src/sobol_sensitivity.jl:94-108) never exist -- and accumulates the per-sample estimator terms exactly as the reference
// forms them (difference first, then sum; :192, :197, :211-213).
//
// This is the path of every model without structure the library can exploit -- in particular of user `__device__` models
// linked from a fatbin (user_models.cu) -- so it is built to keep the FP64 pipe busy with MODEL arithmetic:
//   * rows are staged through a per-warp shared-memory tile: a warp reads 32 rows x 32 coordinates as 32 coalesced
//     256-byte requests (all in flight at once) and every thread then picks up its own row conflict-free; the rows live in
//     thread-local arrays (interleaved by the hardware: x[k] of the 32 lanes is one 256-byte line), so the model gets a
//     plain `const double* x`.  (Round 1 read A[s*d + k] with a lane stride of d doubles: every request touched 32 lines.)
//   * the first- and total-order terms accumulate in PER-THREAD arrays for the whole tile; the warp butterflies run once
//     per tile and coordinate, not once per sample and coordinate (round 1: two 5-step butterflies per (sample, k, output)).
//   * only the second-order pair sums (d(d-1)/2 per output -- too many for per-thread storage) are still reduced per
//     sample across the warp; the model evaluations they need dominate there anyway (2d evaluations of O(d) each).
// Product-form, additive, linear and small-d models have faster kernels (fused_product.cuh, fused_small.cuh,
// fused_linear.cuh, fused_pairs_mma.cuh).
#pragma once
#include "gsa_common.cuh"
#include "models.cuh"
#include "tail.cuh"

namespace gsa {

constexpr int GENERIC_DMAX = 128;   // coordinates per sample held in thread-local memory
constexpr int GENERIC_NOMAX = 4;    // outputs per model evaluation
constexpr int GENERIC_YMAX = 256;   // d * nout bound when second order keeps all yAB / yBA
// per-thread accumulator capacity (entries = d * nout) of the two instantiations: S (d <= 32) and L (d <= 128)
constexpr int GENERIC_ACC_S = 64, GENERIC_ACC_L = 256;
constexpr int GENERIC_TILE_DOUBLES = 32 * 33;   // per-warp transposition tile

struct GenericArgs {
    const double* A;
    const double* B;
    const double* params;
    double* tile_rec;  // [ntiles][nout][nstat]
    TileGeom g;
    int d, np, nout, nstat, est, t_begin, t_end;
    TailArgs tail;     // tail.counters != nullptr -> stage 2, exchange and host copy inside this launch (tail.cuh)
};

// shared memory: per warp a transposition tile and a record of nout*nstat doubles
__host__ __device__ inline size_t generic_smem_bytes(int warps, int recsz) { return sizeof(double) * (size_t)warps * (GENERIC_TILE_DOUBLES + recsz); }

// rows [row0, row0 + cnt) of a d-column row-major matrix -> x[0..d) of lane `lane` (lanes >= cnt get zeros)
template <int DM>
__device__ __forceinline__ void load_rows_staged(const double* __restrict__ M, int64_t row0, int cnt, int d, double* tile, double (&x)[DM], int lane) {
    for (int c0 = 0; c0 < d; c0 += 32) {
        const int w = min(32, d - c0);
        const double* src = M + (size_t)row0 * d + c0 + lane;
        // 32 row requests of 256 contiguous bytes each, all issued before the first use
        double v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = (lane < w && r < cnt) ? ldg_stream(src + (size_t)r * d) : 0.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[lane * 33 + r] = v[r];   // tile[coordinate][row], bank = (lane + r) mod 32
        __syncwarp();
        for (int j = 0; j < w; ++j) x[c0 + j] = tile[j * 33 + lane];
        __syncwarp();
    }
}

// DM: capacity of the row copies, NACC: capacity (d * nout) of the per-thread accumulators
template <class Model, bool SECOND, int DM, int NACC>
__device__ __forceinline__ void fused_generic_body(const GenericArgs& a) {
    extern __shared__ double smem[];  // [warps][TILE] then [warps][nout*nstat]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int d = a.d, nout = a.nout, nstat = a.nstat, est = a.est;
    const int recsz = nout * nstat;
    double* tile = smem + (size_t)warp * GENERIC_TILE_DOUBLES;
    double* recs = smem + (size_t)nwarps * GENERIC_TILE_DOUBLES;
    double* acc = recs + (size_t)warp * recsz;
    const bool janon = est == GSA_EI_JANON2014;

    double xa[DM], xb[DM];
    double V[NACC], E0[NACC], E1[NACC], E2[NACC];   // E1, E2: Janon2014 only (never touched otherwise)
    double yA[GENERIC_NOMAX], yB[GENERIC_NOMAX], yk[GENERIC_NOMAX];
    double yAB[SECOND ? GENERIC_YMAX : 1], yBA[SECOND ? GENERIC_YMAX : 1];

    for (int t = a.t_begin + blockIdx.x; t < a.t_end; t += gridDim.x) {
        int64_t c0, c1;
        tile_range(a.g, t, c0, c1);
        for (int i = lane; i < recsz; i += 32) acc[i] = 0.0;
        for (int i = 0; i < d * nout; ++i) {
            V[i] = 0.0; E0[i] = 0.0;
            if (janon) { E1[i] = 0.0; E2[i] = 0.0; }
        }
        dd sy[GENERIC_NOMAX], syy[GENERIC_NOMAX], sa[GENERIC_NOMAX], saa[GENERIC_NOMAX];
#pragma unroll
        for (int o = 0; o < GENERIC_NOMAX; ++o) sy[o] = syy[o] = sa[o] = saa[o] = dd{0.0, 0.0};
        __syncwarp();

        for (int64_t s0 = c0 + warp * 32; s0 < c1; s0 += 32 * nwarps) {
            const int cnt = (int)min((int64_t)32, c1 - s0);
            const bool active = lane < cnt;
            load_rows_staged<DM>(a.A, s0 - a.g.buf_col0, cnt, d, tile, xa, lane);
            load_rows_staged<DM>(a.B, s0 - a.g.buf_col0, cnt, d, tile, xb, lane);
            if (active) {
                Model::eval(xb, d, a.params, a.np, yB, nout);
                Model::eval(xa, d, a.params, a.np, yA, nout);
#pragma unroll
                for (int o = 0; o < GENERIC_NOMAX; ++o)
                    if (o < nout) {
                        dd_add(sy[o], yA[o]); dd_add(sy[o], yB[o]);
                        dd_add_sq(syy[o], yA[o]); dd_add_sq(syy[o], yB[o]);
                        dd_add(sa[o], yA[o]); dd_add_sq(saa[o], yA[o]);
                    }
                // xa holds A_s.  AB_k: coordinate k taken from B (reference :96-99)
                for (int k = 0; k < d; ++k) {
                    const double keep = xa[k];
                    xa[k] = xb[k];
                    Model::eval(xa, d, a.params, a.np, yk, nout);
                    xa[k] = keep;
                    for (int o = 0; o < nout; ++o) {
                        const double fa = yA[o], fab = yk[o];
                        const int i = k * nout + o;
                        V[i] += yB[o] * (fab - fa);                                                   // :192
                        if (est == GSA_EI_JANSEN1999) { const double df = fa - fab; E0[i] += df * df; }   // :213
                        else if (est == GSA_EI_SOBOL2007) E0[i] += fa * (fa - fab);                   // :211
                        else {
                            E0[i] += fa * fab;                                                        // :207
                            if (janon) { E1[i] += fa * fa + fab * fab; E2[i] += fa + fab; }           // :219-224
                        }
                        if (SECOND) yAB[i] = fab;
                    }
                }
            }
            if (SECOND) {
                // BA_k: B with coordinate k taken from A (:100-104)
                if (active)
                    for (int k = 0; k < d; ++k) {
                        const double keep = xb[k];
                        xb[k] = xa[k];
                        Model::eval(xb, d, a.params, a.np, yk, nout);
                        xb[k] = keep;
                        for (int o = 0; o < nout; ++o) yBA[k * nout + o] = yk[o];
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
        // ---- tile flush: one butterfly per (coordinate, output) and per double-double sum, into the warp's record ----
        for (int k = 0; k < d; ++k)
            for (int o = 0; o < nout; ++o) {
                const int i = k * nout + o;
                const double v = warp_sum(V[i]), e0 = warp_sum(E0[i]);
                double e1 = 0.0, e2 = 0.0;
                if (janon) { e1 = warp_sum(E1[i]); e2 = warp_sum(E2[i]); }
                if (lane == 0) {
                    double* r = acc + o * nstat;
                    r[stat_v(k)] = v;
                    r[stat_e(d, 0, k)] = e0;
                    if (janon) { r[stat_e(d, 1, k)] = e1; r[stat_e(d, 2, k)] = e2; }
                }
            }
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
                    for (int w = 0; w < nwarps; ++w) dd_add_dd(s, dd{recs[(size_t)w * recsz + i], recs[(size_t)w * recsz + i + 1]});
                    out[i] = s.hi;
                    out[i + 1] = s.lo;
                }
            } else {
                double s = 0.0;
                for (int w = 0; w < nwarps; ++w) s += recs[(size_t)w * recsz + i];
                out[i] = s;
            }
        }
        __syncthreads();
        if (a.tail.counters) fused_tail_tile_done(a.tail, t);
    }
}

// S: d <= 32 and d*nout <= 64;  L: d <= 128 and d*nout <= 256  (d*nout beyond that: GSA_ERR_UNSUPPORTED)
// (CTAs are at most four warps, one per SM sub-partition: see plan_generic)
constexpr int GENERIC_THREADS_MAX = 128;
template <class Model, bool SECOND, bool LARGE>
__global__ void __launch_bounds__(GENERIC_THREADS_MAX, 5) fused_generic_kernel(const __grid_constant__ GenericArgs a) {
    fused_generic_body<Model, SECOND, LARGE ? GENERIC_DMAX : 32, LARGE ? GENERIC_ACC_L : GENERIC_ACC_S>(a);
}

}  // namespace gsa
