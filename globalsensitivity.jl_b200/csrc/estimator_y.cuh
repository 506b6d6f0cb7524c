// estimator_y.cuh -- K3: the Jansen/Saltelli estimator sums on a precomputed model output Y
// (replaces the slice-and-sum loops of gsa_sobol_all_y_analysis, reference src/sobol_sensitivity.jl:180-317).
//
// Y is nout x (nboot*n*step), Julia column-major, blocks per replicate ordered [A | B | AB_1..AB_d | BA_1..BA_d]
// (:105-107, :181-190).  Purely HBM-bound: every Y value is needed once (8*nout bytes per Saltelli row).
// A CTA owns one tile (replicate r, sample positions [c0,c1)) and one output o.  fA and fB of the tile are
// staged in shared memory (they are needed by every k); then each WARP owns whole blocks k and streams
// fAB_k over the tile with per-lane register accumulators, so no cross-thread reduction happens inside
// the streaming loop -- one xor butterfly per (k, tile) at the end.
#pragma once
#include "gsa_common.cuh"

namespace gsa {

constexpr int KY_TS_DEFAULT = 2048;  // samples per tile (16 bytes of dynamic shared memory each)
constexpr int KY_TS_LIMIT = 12288;   // 192 KB
constexpr int KY_U = 8;               // loads in flight per lane in the streaming loops

struct EstYArgs {
    const double* Y;
    double* tile_rec;  // [ntiles][nout][nstat]
    int64_t n;         // columns per block
    int nout, d, nstat, est, step, ts, tpr, ntiles;
};

template <bool SECOND>
__global__ void __launch_bounds__(1024) estimator_y_kernel(EstYArgs a) {
    extern __shared__ double sm_y[];
    double* sA = sm_y;
    double* sB = sm_y + a.ts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int o = blockIdx.y, d = a.d, est = a.est;
    const int64_t n = a.n, nout = a.nout;
    for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const int r = t / a.tpr, jt = t % a.tpr;
        const int64_t c0 = (int64_t)jt * a.ts;
        const int len = (int)min((int64_t)a.ts, n - c0);
        // block b of replicate r, sample position c0 + s, output o
        const double* base = a.Y + o + nout * ((int64_t)r * a.step * n + c0);
        auto blk = [&](int b) { return base + nout * n * b; };
        double* rec = a.tile_rec + ((size_t)t * a.nout + o) * a.nstat;

        ShiftSums accA, accB;   // shifted + compensated sums of fA, fA^2 and fB, fB^2
        const double *pA = blk(0), *pB = blk(1);
        for (int s = threadIdx.x; s < len; s += blockDim.x) {
            double fa = ldg_stream(pA + nout * s), fb = ldg_stream(pB + nout * s);
            sA[s] = fa; sB[s] = fb;
            accA.add(fa);
            accB.add(fb);
        }
        __syncthreads();
        // dd sums: warp butterfly, then warp 0 combines the warps through shared memory (reuse after sync below)
        __shared__ double sdd[32][8];
        {
            dd sy, syy, sa, saa, sb, sbb;
            accA.to_dd(sa, saa);
            accB.to_dd(sb, sbb);
            sy = sa; syy = saa;
            dd_add_dd(sy, sb); dd_add_dd(syy, sbb);
            dd r0 = warp_sum_dd(sy), r1 = warp_sum_dd(syy), r2 = warp_sum_dd(sa), r3 = warp_sum_dd(saa);
            if (lane == 0) {
                sdd[warp][0] = r0.hi; sdd[warp][1] = r0.lo; sdd[warp][2] = r1.hi; sdd[warp][3] = r1.lo;
                sdd[warp][4] = r2.hi; sdd[warp][5] = r2.lo; sdd[warp][6] = r3.hi; sdd[warp][7] = r3.lo;
            }
        }
        for (int k = warp; k < d; k += nwarps) {
            const double* pk = blk(2 + k);
            double v = 0.0, e0 = 0.0, e1 = 0.0, e2 = 0.0;
            // batches of KY_U loads issued back to back (KY_U * 256 bytes in flight per warp), then consumed
            for (int s0 = lane; s0 < len; s0 += 32 * KY_U) {
                double f[KY_U];
#pragma unroll
                for (int u = 0; u < KY_U; ++u) {
                    const int s = s0 + 32 * u;
                    f[u] = s < len ? ldg_stream(pk + nout * s) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < KY_U; ++u) {
                    const int s = s0 + 32 * u;
                    if (s < len) {
                        const double fab = f[u], fa = sA[s], fb = sB[s];
                        v += fb * (fab - fa);                                             // :192
                        if (est == GSA_EI_JANSEN1999) { double df = fa - fab; e0 += df * df; }   // :213
                        else if (est == GSA_EI_SOBOL2007) e0 += fa * (fa - fab);          // :211
                        else { e0 += fa * fab; e1 += fa * fa + fab * fab; e2 += fa + fab; }      // :207, :219-224
                    }
                }
            }
            v = warp_sum(v); e0 = warp_sum(e0);
            if (est == GSA_EI_JANON2014) { e1 = warp_sum(e1); e2 = warp_sum(e2); }
            if (lane == 0) {
                rec[stat_v(k)] = v;
                rec[stat_e(d, 0, k)] = e0;
                if (est == GSA_EI_JANON2014) { rec[stat_e(d, 1, k)] = e1; rec[stat_e(d, 2, k)] = e2; }
            }
        }
        if (SECOND) {
            const int npairs = d * (d - 1) / 2, pb = stat_pair_base(d, est);
            for (int p = warp; p < npairs; p += nwarps) {
                // invert the k-major pair index
                int k = 0, rem = p;
                while (rem >= d - 1 - k) { rem -= d - 1 - k; ++k; }
                const int j = k + 1 + rem;
                const double *pbk = blk(2 + d + k), *paj = blk(2 + j);
                double v = 0.0;
                for (int s0 = lane; s0 < len; s0 += 32 * KY_U) {
                    double fb_[KY_U], fa_[KY_U];
#pragma unroll
                    for (int u = 0; u < KY_U; ++u) {
                        const int s = s0 + 32 * u;
                        fb_[u] = s < len ? ldg_stream(pbk + nout * s) : 0.0;
                        fa_[u] = s < len ? ldg_stream(paj + nout * s) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < KY_U; ++u) {
                        const int s = s0 + 32 * u;
                        if (s < len) v += fb_[u] * fa_[u] - sA[s] * sB[s];                  // :197
                    }
                }
                v = warp_sum(v);
                if (lane == 0) rec[pb + p] = v;
            }
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            dd s{0.0, 0.0};
            for (int w = 0; w < nwarps; ++w) dd_add_dd(s, dd{sdd[w][2 * threadIdx.x], sdd[w][2 * threadIdx.x + 1]});
            rec[2 * threadIdx.x] = s.hi;
            rec[2 * threadIdx.x + 1] = s.lo;
        }
        __syncthreads();
    }
}

}  // namespace gsa
