// fused_linear.cuh -- K2+K3 for LINEAR multi-output models  y = W x  (BASELINE config 4: d = 20, 256 outputs), S1/ST.
//
// For a linear model every quantity the estimators need is a quadratic form of INPUT moments, so the kernel never
// evaluates the nout outputs per sample.  With a = A_s - c, b = B_s - c (c = first row of the tile, a numerical
// shift only), delta = b - a and w = W[o, :]:
//     yA = y0 + w.a,  yB = y0 + w.b,  y0 = w.c,      yAB_k - yA = W_ok * delta_k            (rank-1 update)
//     sum (yA-y0)^2 + (yB-y0)^2   = w' G w,          G  = sum a a' + b b'                    -> Var   (:183)
//     sum yB (yAB_k - yA)         = W_ok (y0 sum delta_k + sum_l w_l H[l][k]),  H = sum b delta'   -> V_k   (:192)
//     sum (yA - yAB_k)^2          = W_ok^2 D[k],     D  = sum delta_k^2                      -> E_k   (:213)
// Cost per sample: O(d^2) FMAs, independent of nout (the reference's path is O(d * nout * (d+2))): the 189 GB
// `all_y` of config 4 is never formed, and neither is any per-output intermediate.
//
// Kernel 1 (gram): lane l of a warp owns ROW l of G and H (extended with a constant-1 coordinate at index d so
// that first moments come out of the same products); the column values of a sample are broadcast from shared
// memory with warp-uniform 16-byte loads, so one LDS.128 feeds six DFMA warp instructions.  Warps take
// different samples of the tile and are combined in order at the end.
// Kernel 2 (records): per (tile, output) turns the tile's moments into the library's standard sufficient-
// statistics record (double-double sums of y and y^2 recentred exactly), so tiles, replicates and GPUs combine
// through the same reduce / all-reduce / finalize path as every other model.
#pragma once
#include "gsa_common.cuh"

namespace gsa {

constexpr int LIN_THREADS = 256;
constexpr int LIN_S = 64;      // samples staged per pass (tile granule); wide models stage 48 to stay under 48 KB
constexpr int LIN_DPMAX = 32;  // d + 1 <= 32

struct LinearArgs {
    const double* A;
    const double* B;
    double* gram;  // [ntiles][2*DP*DP + DP + DP]: G (DP x DP, row-major), H (DP x DP), D (DP), c (DP; c[d] = #samples)
    TileGeom g;
    int d, t_begin, t_end;
};

__host__ __device__ inline int lin_gram_size(int DP) { return 2 * DP * DP + 2 * DP; }

template <int DP>
__global__ void __launch_bounds__(LIN_THREADS) linear_gram_kernel(LinearArgs a) {
    constexpr int NW = LIN_THREADS / 32;
    constexpr int REC = 3 * DP;  // per staged sample: a_ext[DP] | b_ext[DP] | delta_ext[DP]
    constexpr int SS = DP <= 24 ? LIN_S : 48;
    __shared__ __align__(16) double stage[SS * REC];
    __shared__ double cshift[DP];
    const int d = a.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int tile = a.t_begin + blockIdx.x; tile < a.t_end; tile += gridDim.x) {
        int64_t c0, c1;
        tile_range(a.g, tile, c0, c1);
        double* out = a.gram + (size_t)tile * lin_gram_size(DP);
        if (c1 <= c0) {
            for (int i = threadIdx.x; i < lin_gram_size(DP); i += LIN_THREADS) out[i] = 0.0;
            continue;
        }
        if (threadIdx.x < DP) cshift[threadIdx.x] = threadIdx.x < d ? __ldg(a.A + (size_t)(c0 - a.g.buf_col0) * d + threadIdx.x) : 0.0;
        double G[DP], H[DP], Dl = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) G[k] = H[k] = 0.0;
        __syncthreads();

        for (int64_t p0 = c0; p0 < c1; p0 += SS) {
            const int ns = (int)min((int64_t)SS, c1 - p0);
            // ---- stage: coalesced read of ns rows of A and B, write shifted / differenced records ----
            const double* ga = a.A + (size_t)(p0 - a.g.buf_col0) * d;
            const double* gb = a.B + (size_t)(p0 - a.g.buf_col0) * d;
            for (int e = threadIdx.x; e < ns * DP; e += LIN_THREADS) {
                const int s = e / DP, k = e - s * DP;
                double va = 0.0, vb = 0.0, vd = 0.0;
                if (k < d) {
                    const double xa = ldg_stream(ga + s * d + k), xb = ldg_stream(gb + s * d + k), c = cshift[k];
                    va = xa - c;
                    vb = xb - c;
                    vd = xb - xa;
                } else if (k == d) {
                    va = vb = 1.0;  // constant coordinate: row/column d of G and H carry the first moments
                }
                double* r = stage + s * REC;
                r[k] = va;
                r[DP + k] = vb;
                r[2 * DP + k] = vd;
            }
            __syncthreads();
            // ---- accumulate: warp w takes samples w, w+NW, ...; lane l owns row l ----
            for (int s = warp; s < ns; s += NW) {
                const double* r = stage + s * REC;
                const int l = lane < DP ? lane : 0;
                const double ra = r[l], rb = r[DP + l], rd = r[2 * DP + l];
                Dl = fma(rd, rd, Dl);
#pragma unroll
                for (int k = 0; k < DP; k += 2) {
                    const double2 ca = *reinterpret_cast<const double2*>(r + k);            // warp-uniform address
                    const double2 cb = *reinterpret_cast<const double2*>(r + DP + k);
                    const double2 cd = *reinterpret_cast<const double2*>(r + 2 * DP + k);
                    G[k] = fma(ra, ca.x, G[k]);
                    G[k] = fma(rb, cb.x, G[k]);
                    G[k + 1] = fma(ra, ca.y, G[k + 1]);
                    G[k + 1] = fma(rb, cb.y, G[k + 1]);
                    H[k] = fma(rb, cd.x, H[k]);
                    H[k + 1] = fma(rb, cd.y, H[k + 1]);
                }
            }
            __syncthreads();
        }
        // ---- combine the NW warps in a fixed order: they take turns adding their rows into one buffer (the staging area) ----
        static_assert(2 * DP * DP + DP <= SS * REC, "staging buffer too small for the warp combine");
        for (int w = 0; w < NW; ++w) {
            if (warp == w && lane < DP) {
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    double* pg = stage + lane * DP + k;
                    double* ph = stage + DP * DP + lane * DP + k;
                    *pg = w == 0 ? G[k] : *pg + G[k];
                    *ph = w == 0 ? H[k] : *ph + H[k];
                }
                double* pd = stage + 2 * DP * DP + lane;
                *pd = w == 0 ? Dl : *pd + Dl;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < 2 * DP * DP + DP; i += LIN_THREADS) out[i] = stage[i];
        if (threadIdx.x < DP) out[2 * DP * DP + DP + threadIdx.x] = threadIdx.x < d ? cshift[threadIdx.x] : (threadIdx.x == d ? (double)(c1 - c0) : 0.0);
        __syncthreads();
    }
}

// moments of a tile -> standard record of output o.  One thread per output.
struct LinearRecArgs {
    const double* gram;
    const double* W;   // nout x d column-major
    double* tile_rec;  // [ntiles][nout][nstat]
    int d, DP, nout, nstat, t_begin, t_end;
};

__device__ __forceinline__ dd dd_from(double x) { return dd{x, 0.0}; }
__device__ __forceinline__ dd dd_mul_d(dd a, double b) {  // a * b
    double p = a.hi * b;
    double e = __fma_rn(a.hi, b, -p);
    e = __fma_rn(a.lo, b, e);
    return quick_two_sum(p, e);
}

__global__ void __launch_bounds__(256) linear_records_kernel(LinearRecArgs a) {
    extern __shared__ double sg[];  // the tile's moments
    const int d = a.d, DP = a.DP, gs = lin_gram_size(DP);
    for (int tile = a.t_begin + blockIdx.x; tile < a.t_end; tile += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < gs; i += blockDim.x) sg[i] = a.gram[(size_t)tile * gs + i];
        __syncthreads();
        const double* G = sg;
        const double* H = sg + DP * DP;
        const double* D = sg + 2 * DP * DP;
        const double* c = D + DP;
        const double m = c[d];  // samples in the tile
        for (int o = threadIdx.x; o < a.nout; o += blockDim.x) {
            double* rec = a.tile_rec + ((size_t)tile * a.nout + o) * a.nstat;
            for (int i = 0; i < a.nstat; ++i) rec[i] = 0.0;
            if (m == 0.0) continue;
            // y0 = w.c ; S' = sum_l w_l G[l][d] ; Q' = w' G w (l,k < d)
            double y0 = 0.0, S1 = 0.0, Q = 0.0;
            for (int l = 0; l < d; ++l) {
                const double wl = __ldg(a.W + o + (size_t)a.nout * l);
                y0 = fma(wl, c[l], y0);
                S1 = fma(wl, G[l * DP + d], S1);
                double row = 0.0;
                for (int k = 0; k < d; ++k) row = fma(__ldg(a.W + o + (size_t)a.nout * k), G[l * DP + k], row);
                Q = fma(wl, row, Q);
            }
            // un-shift exactly: sum y = S' + 2m y0 ; sum y^2 = Q' + 2 y0 S' + 2m y0^2   (double-double)
            dd sy = dd_from(S1);
            dd_add_dd(sy, dd_mul_d(dd_from(2.0 * m), y0));
            dd t = dd_mul_d(dd_from(y0), y0);              // y0^2
            dd syy = dd_from(Q);
            dd_add_dd(syy, dd_mul_d(dd_mul_d(dd_from(2.0 * y0), S1), 1.0));
            dd_add_dd(syy, dd_mul_d(t, 2.0 * m));
            rec[0] = sy.hi; rec[1] = sy.lo; rec[2] = syy.hi; rec[3] = syy.lo;
            for (int k = 0; k < d; ++k) {
                const double wk = __ldg(a.W + o + (size_t)a.nout * k);
                double hv = y0 * H[d * DP + k];            // y0 * sum delta_k   (row d of H: b_ext[d] = 1)
                for (int l = 0; l < d; ++l) hv = fma(__ldg(a.W + o + (size_t)a.nout * l), H[l * DP + k], hv);
                rec[stat_v(k)] = wk * hv;                  // sum yB (yAB_k - yA)       (:192)
                rec[stat_e(d, 0, k)] = wk * wk * D[k];     // sum (yA - yAB_k)^2        (:213)
            }
        }
    }
}

}  // namespace gsa
