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
// Kernel 1 (gram): the vectors are extended with a constant-1 coordinate at index d (first moments come out of the same
// products) and cut into 8-double blocks; every LANE owns one 8x8 block pair -- (a,a) and (b,b) upper triangle, (b,delta),
// (delta,delta) diagonal -- i.e. 64 register accumulators fed by 16 operands per sample (4 FMA per operand delivered from
// shared memory; a row-per-lane layout with broadcast loads delivered 1 and ran 3.5x slower, profiles/r1_notes.md).
// Blocks are stored with a 10-double stride so the per-lane 16-byte loads of a quarter-warp fall in distinct banks.
// Kernel 2 (records): per (tile, output) turns the tile's moments into the library's standard sufficient-
// statistics record (double-double sums of y and y^2 recentred exactly), so tiles, replicates and GPUs combine
// through the same reduce / all-reduce / finalize path as every other model.
#pragma once
#include "gsa_common.cuh"

namespace gsa {

constexpr int LIN_THREADS = 128;  // ~190 registers/thread: two or three CTAs per SM overlap staging and accumulation
constexpr int LIN_S = 48;         // samples staged per pass (tile granule)
constexpr int LIN_DPMAX = 32;     // d + 1 <= 32
constexpr int LIN_BS = 10;        // shared-memory stride of an 8-double block (padded: conflict-free 16-byte lane loads)

struct LinearArgs {
    const double* A;
    const double* B;
    double* gram;  // [ntiles][lin_gram_size(DP)]: G (DP x DP, row-major, full), H (DP x DP), D (DP), c (DP; c[d] = #samples)
    TileGeom g;
    int d, t_begin, t_end;
};

#ifdef __CUDACC__
__device__ __forceinline__ void cp_async8_l(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_l(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
#endif
// dynamic shared memory of linear_gram_kernel (bytes)
__host__ __device__ inline size_t lin_smem_bytes(int NB, int d) {
    return sizeof(double) * ((size_t)LIN_S * 3 * NB * 10 + 8 * NB + 4 * (size_t)LIN_S * d);
}
__host__ __device__ inline int lin_gram_size(int DP) { return 2 * DP * DP + 2 * DP; }
// 8x8 block pairs of the extended vectors X = [a_ext | b_ext | delta_ext] (NB blocks each):
//   a x a and b x b upper-triangular block pairs, b x delta (all), delta x delta (diagonal blocks)
__host__ __device__ inline int lin_num_pairs(int NB, int NBD) { return NB * (NB + 1) + NB * NBD + NBD; }

// decode pair p -> (row vector, row block, col vector, col block); vectors: 0 = a, 1 = b, 2 = delta
__host__ __device__ inline void lin_pair(int p, int NB, int NBD, int& rv, int& rb, int& cv, int& cb) {
    const int tri = NB * (NB + 1) / 2;
    if (p < 2 * tri) {
        rv = cv = p / tri;
        int q = p % tri;
        rb = 0;
        while (q >= NB - rb) { q -= NB - rb; ++rb; }
        cb = rb + q;
    } else if (p < 2 * tri + NB * NBD) {
        const int q = p - 2 * tri;
        rv = 1; rb = q / NBD; cv = 2; cb = q % NBD;
    } else {
        rv = cv = 2;
        rb = cb = p - 2 * tri - NB * NBD;
    }
}

// NB = number of 8-blocks of the extended (d+1)-vector; LPS = lanes per sample (power of two >= number of pairs)
template <int NB, int LPS>
__global__ void __launch_bounds__(LIN_THREADS) linear_gram_kernel(LinearArgs a) {
    constexpr int DP = 8 * NB;
    constexpr int REC = 3 * NB * LIN_BS;            // doubles per staged sample
    constexpr int SPW = LPS >= 32 ? 1 : 32 / LPS;   // samples per warp step
    constexpr int WPS = LPS >= 32 ? LPS / 32 : 1;   // warps per sample
    constexpr int NW = LIN_THREADS / 32;
    constexpr int SLOTS = NW * SPW / WPS;           // samples processed concurrently by the CTA
    extern __shared__ __align__(16) double lin_sm[];
    double* stage = lin_sm;                          // [LIN_S][REC]   shifted / differenced, block-padded records
    double* cshift = lin_sm + LIN_S * REC;           // [DP]
    double* raw = cshift + DP;                       // [2][2][LIN_S * d]  raw A and B rows of the next pass (cp.async ring)
    const int d = a.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rawsz = LIN_S * d;
    const int NBD = (d + 7) / 8, npairs = lin_num_pairs(NB, NBD);
    // this lane's pair and sample slot
    const int pair = LPS >= 32 ? (warp % WPS) * 32 + lane : lane % LPS;
    const int slot = LPS >= 32 ? warp / WPS : warp * SPW + lane / LPS;
    const bool active = pair < npairs;
    int rv = 0, rb = 0, cv = 0, cb = 0;
    if (active) lin_pair(pair, NB, NBD, rv, rb, cv, cb);
    const int roff = (rv * NB + rb) * LIN_BS, coff = (cv * NB + cb) * LIN_BS;

    for (int tile = a.t_begin + blockIdx.x; tile < a.t_end; tile += gridDim.x) {
        int64_t c0, c1;
        tile_range(a.g, tile, c0, c1);
        double* out = a.gram + (size_t)tile * lin_gram_size(DP);
        if (c1 <= c0) {
            for (int i = threadIdx.x; i < lin_gram_size(DP); i += LIN_THREADS) out[i] = 0.0;
            continue;
        }
        if (threadIdx.x < DP) cshift[threadIdx.x] = threadIdx.x < d ? __ldg(a.A + (size_t)(c0 - a.g.buf_col0) * d + threadIdx.x) : 0.0;
        double acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.0;

        // asynchronous copy of the raw rows of one pass (contiguous ns*d doubles of A and of B) into ring slot `buf`
        auto prefetch = [&](int64_t p0, int buf) {
            if (p0 < c1) {
                const int cnt = (int)min((int64_t)LIN_S, c1 - p0) * d;
                const double* ga = a.A + (size_t)(p0 - a.g.buf_col0) * d;
                const double* gb = a.B + (size_t)(p0 - a.g.buf_col0) * d;
                double* ra = raw + (size_t)buf * 2 * rawsz;
                double* rbuf = ra + rawsz;
                const bool al16 = ((reinterpret_cast<uintptr_t>(ga) | reinterpret_cast<uintptr_t>(gb)) & 15) == 0 && (rawsz % 2 == 0);
                if (al16) {
                    for (int e = 2 * threadIdx.x; e < cnt; e += 2 * LIN_THREADS) {
                        if (e + 1 < cnt) { cp_async16_l(ra + e, ga + e); cp_async16_l(rbuf + e, gb + e); }
                        else { cp_async8_l(ra + e, ga + e); cp_async8_l(rbuf + e, gb + e); }
                    }
                } else {
                    for (int e = threadIdx.x; e < cnt; e += LIN_THREADS) { cp_async8_l(ra + e, ga + e); cp_async8_l(rbuf + e, gb + e); }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        prefetch(c0, 0);
        int buf = 0;
        for (int64_t p0 = c0; p0 < c1; p0 += LIN_S, buf ^= 1) {
            const int ns = (int)min((int64_t)LIN_S, c1 - p0);
            prefetch(p0 + LIN_S, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncthreads();   // every thread's copies of this pass are visible; the previous accumulate is finished
            // ---- transform: raw rows -> shifted / differenced, block-padded records ----
            const double* ra = raw + (size_t)buf * 2 * rawsz;
            const double* rbuf = ra + rawsz;
            for (int e = threadIdx.x; e < ns * DP; e += LIN_THREADS) {
                const int s = e / DP, k = e - s * DP;
                double va = 0.0, vb = 0.0, vd = 0.0;
                if (k < d) {
                    const double xa = ra[s * d + k], xb = rbuf[s * d + k], c = cshift[k];
                    va = xa - c;
                    vb = xb - c;
                    vd = xb - xa;
                } else if (k == d) {
                    va = vb = 1.0;  // constant coordinate: row/column d of G and H carry the first moments
                }
                double* r = stage + s * REC + (k >> 3) * LIN_BS + (k & 7);
                r[0] = va;
                r[NB * LIN_BS] = vb;
                r[2 * NB * LIN_BS] = vd;
            }
            __syncthreads();
            // ---- accumulate: every lane adds the 8x8 outer product of its (row block, column block) ----
            for (int s = slot; s < ns; s += SLOTS) {
                if (active) {
                    const double* r = stage + s * REC;
                    double rr[8], cc[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double2 x = *reinterpret_cast<const double2*>(r + roff + 2 * j);
                        const double2 y = *reinterpret_cast<const double2*>(r + coff + 2 * j);
                        rr[2 * j] = x.x; rr[2 * j + 1] = x.y; cc[2 * j] = y.x; cc[2 * j + 1] = y.y;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[8 * i + j] = fma(rr[i], cc[j], acc[8 * i + j]);
                }
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- combine: dense G (a x a + b x b, both triangles), H, D in shared memory; sample slots take turns (fixed order) ----
        double* G = stage;               // [DP][DP]
        double* H = stage + DP * DP;     // [DP][DP]
        double* Dv = H + DP * DP;        // [DP]
        static_assert(2 * DP * DP + DP <= LIN_S * REC, "staging buffer too small for the combine");
        for (int i = threadIdx.x; i < 2 * DP * DP + DP; i += LIN_THREADS) stage[i] = 0.0;
        __syncthreads();
        for (int turn = 0; turn < SLOTS * 2; ++turn) {
            // a x a pairs and b x b pairs hit the same G entries: give them separate turns (rv = 0 first, then the rest)
            const int tslot = turn >> 1, phase = turn & 1;
            if (active && slot == tslot && ((rv == 0) == (phase == 0))) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int l = 8 * rb + i, k = 8 * cb + j;
                        const double v = acc[8 * i + j];
                        if (rv == cv && rv < 2) {
                            G[l * DP + k] += v;
                            if (rb != cb) G[k * DP + l] += v;   // mirror the off-diagonal blocks
                        } else if (rv == 1) {
                            H[l * DP + k] += v;
                        } else if (i == j) {
                            Dv[k] += v;
                        }
                    }
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
            double w[LIN_DPMAX];                           // this output's row of W (coalesced loads across outputs)
            for (int l = 0; l < d; ++l) w[l] = __ldg(a.W + o + (size_t)a.nout * l);
            // y0 = w.c ; S' = sum_l w_l G[l][d] ; Q' = w' G w (l,k < d)
            double y0 = 0.0, S1 = 0.0, Q = 0.0;
            for (int l = 0; l < d; ++l) {
                const double wl = w[l];
                y0 = fma(wl, c[l], y0);
                S1 = fma(wl, G[l * DP + d], S1);
                double row = 0.0;
                for (int k = 0; k < d; ++k) row = fma(w[k], G[l * DP + k], row);
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
                const double wk = w[k];
                double hv = y0 * H[d * DP + k];            // y0 * sum delta_k   (row d of H: b_ext[d] = 1)
                for (int l = 0; l < d; ++l) hv = fma(w[l], H[l * DP + k], hv);
                rec[stat_v(k)] = wk * hv;                  // sum yB (yAB_k - yA)       (:192)
                rec[stat_e(d, 0, k)] = wk * wk * D[k];     // sum (yA - yAB_k)^2        (:213)
            }
        }
    }
}

}  // namespace gsa
