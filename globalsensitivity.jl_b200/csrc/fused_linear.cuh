// fused_linear.cuh -- K2+K3 for LINEAR multi-output models  y = W x  (BASELINE config 4: d = 20, 256 outputs), S1/ST.
//
// For a linear model every quantity the estimators need is a quadratic form of INPUT moments, so the kernel never
// evaluates the nout outputs per sample.  With a = A_s - c, b = B_s - c (c = the first row of the replicate that this
// launch sees: a numerical shift only), u = [a; b] (2d coordinates) and w = W[o, :]:
//     yA = y0 + w.a,  yB = y0 + w.b,  y0 = w.c,      yAB_k - yA = W_ok (b_k - a_k)            (rank-1 update)
//     sum (yA-y0)^2 + (yB-y0)^2   = w' (Uaa + Ubb) w                                          -> Var   (:183)
//     sum yB (yAB_k - yA)         = W_ok (y0 (sb_k - sa_k) + sum_l w_l (Ubb - Uba)[l][k])     -> V_k   (:192)
//     sum (yA - yAB_k)^2          = W_ok^2 (Ubb - 2 Uab + Uaa)[k][k]                          -> E_k   (:213)
// where U = sum_s u u' (the 2d x 2d Gram matrix of the stacked rows) and sa, sb the first moments.  Cost per sample:
// O(d^2) FMAs, independent of nout (the reference's path is O(d * nout * (d+2))): the 189 GB `all_y` of config 4 is
// never formed, and neither is any per-output intermediate.
//
// Kernel 1 (gram) -- the one real dense contraction of this library (U = X'X with the SAMPLES as the contraction index), so it
// runs on the FP64 tensor-core path: `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4; tcgen05 has no FP64 kind).  tools/fp64_peak.cu
// measured DMMA = 37.0 and DFMA = 36.6 TFLOP/s on this B200: the same peak, but a DMMA carries 256 FMAs per issue slot and takes
// its operands from registers.  The previous DFMA version (8x8 register blocks fed from shared memory) sat at 42 % of that peak
// on operand delivery.  u is cut into NB = ceil(2d/8) blocks of 8 coordinates; a warp step covers 4 consecutive samples: thread
// (g = lane/4, t = lane%4) loads coordinate 8*blk+g of sample t straight from global memory (the 8 lanes of a sample read
// 64 contiguous bytes), which is at once the A fragment (row g, k = t) and the B fragment (k = t, column g) of every block pair
// -- no transposition.  NB(NB+1)/2 DMMAs per step accumulate the upper block triangle of U.  The operands of the next groups of
// steps are in flight as per-thread cp.async copies into a private shared-memory ring while the DMMAs of the current group run.
// Kernel 2 sums the per-tile moments of a replicate (fixed order).  Kernel 3 turns them into the library's standard sufficient-
// statistics record per (replicate, output) -- double-double sums of y and y^2 recentred exactly -- and adds it to `partials`.
#pragma once
#include "gsa_common.cuh"

namespace gsa {

constexpr int LIN_THREADS = 128;
constexpr int LIN_NW = LIN_THREADS / 32;
constexpr int LIN_DMAX = 32;      // 2d <= 64: at most 8 blocks, 36 accumulator pairs per thread

// per-tile / per-replicate moments: U[DU][DU] (row-major, only block-upper-triangular entries are written), S[DU], count
__host__ __device__ inline int lin_gram_size(int NB) { return 64 * NB * NB + 8 * NB + 8; }

struct LinearArgs {
    const double* A;
    const double* B;
    double* gram;   // [ntiles][lin_gram_size(NB)]
    TileGeom g;
    int d, t_begin, t_end;
};

// first global column of replicate r that this launch holds: the row used as shift for all of the replicate's tiles
__host__ __device__ inline int64_t lin_shift_col(const TileGeom& g, int r) {
    int64_t c = (int64_t)r * g.n;
    if (c < g.col_lo) c = g.col_lo;
    if (c < g.buf_col0) c = g.buf_col0;
    return c;
}

#ifdef __CUDACC__
__device__ __forceinline__ void lin_cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NB = blocks of the stacked vector; UN = warp steps (of 4 samples) per load group; ST = groups in flight per thread.
// Every thread stages its own future operands in a private shared-memory ring with cp.async (no barrier: a thread only ever
// reads what it copied itself), so ST groups = ST*UN*NB 8-byte loads are in flight per thread without holding registers.
__host__ __device__ inline size_t lin_smem_bytes(int NB, int UN, int ST) {
    return sizeof(double) * (64 * (size_t)NB * NB + 8 * NB + (size_t)ST * UN * NB * LIN_THREADS);
}

template <int NB, int UN, int ST>
__global__ void __launch_bounds__(LIN_THREADS) linear_gram_kernel(LinearArgs a) {
    constexpr int DU = 8 * NB, NP = NB * (NB + 1) / 2, GS = 64 * NB * NB + 8 * NB + 8;
    extern __shared__ __align__(16) double lin_sm[];   // U[DU*DU], S[DU], ring[ST][UN][NB][LIN_THREADS]
    double* ring = lin_sm + DU * DU + DU + threadIdx.x;
    const int d = a.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g8 = lane >> 2, t4 = lane & 3;
    // this thread's coordinate of every block: offset inside a row of A (i < d) or of B (d <= i < 2d)
    const double* src[NB];
    bool ok[NB];
    int coff[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        const int i = 8 * k + g8;
        ok[k] = i < 2 * d;
        coff[k] = i < d ? i : i - d;
        src[k] = (i < d ? a.A : a.B) + (ok[k] ? coff[k] : 0);
    }

    for (int tile = a.t_begin + blockIdx.x; tile < a.t_end; tile += gridDim.x) {
        int64_t c0, c1;
        tile_range(a.g, tile, c0, c1);
        double* out = a.gram + (size_t)tile * GS;
        if (c1 <= c0) {
            for (int i = threadIdx.x; i < GS; i += LIN_THREADS) out[i] = 0.0;
            continue;
        }
        const int64_t cs = lin_shift_col(a.g, a.g.rep_first + tile / a.g.tpr);
        double sh[NB], s1[NB], acc[NP][2];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            sh[k] = ok[k] ? __ldg(a.A + (size_t)(cs - a.g.buf_col0) * d + coff[k]) : 0.0;
            s1[k] = 0.0;
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p][0] = acc[p][1] = 0.0;

        const int64_t cnt = c1 - c0, row0 = c0 - a.g.buf_col0;
        const int64_t ngroups = (cnt + 4 * UN - 1) / (4 * UN);
        // group G = steps [G*UN, (G+1)*UN), step q = samples [4q, 4q+4) of the tile; the warps take groups round-robin
        auto issue = [&](int64_t G, int st) {   // asynchronous copies of group G into ring stage st (always commits a group)
            if (G < ngroups) {
                const int64_t s0 = G * (4 * UN) + t4;
                const bool full = (G + 1) * (4 * UN) <= cnt;
#pragma unroll
                for (int j = 0; j < UN; ++j)
#pragma unroll
                    for (int k = 0; k < NB; ++k)
                        if (ok[k] && (full || s0 + 4 * j < cnt))
                            lin_cp_async8(ring + ((st * UN + j) * NB + k) * LIN_THREADS, src[k] + (size_t)(row0 + s0 + 4 * j) * d);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto consume = [&](int64_t G, int st) {
            const int64_t s0 = G * (4 * UN) + t4;
            const bool full = (G + 1) * (4 * UN) <= cnt;
            double v[UN][NB];
#pragma unroll
            for (int j = 0; j < UN; ++j)
#pragma unroll
                for (int k = 0; k < NB; ++k)
                    v[j][k] = (ok[k] && (full || s0 + 4 * j < cnt)) ? ring[((st * UN + j) * NB + k) * LIN_THREADS] - sh[k] : 0.0;
#pragma unroll
            for (int j = 0; j < UN; ++j) {
                int p = 0;
#pragma unroll
                for (int rb = 0; rb < NB; ++rb)
#pragma unroll
                    for (int cb = rb; cb < NB; ++cb, ++p) dmma_8x8x4(acc[p][0], acc[p][1], v[j][rb], v[j][cb]);
#pragma unroll
                for (int k = 0; k < NB; ++k) s1[k] += v[j][k];
            }
        };
#pragma unroll
        for (int st = 0; st < ST - 1; ++st) issue(warp + (int64_t)st * LIN_NW, st);
        int st = 0;
        for (int64_t G = warp; G < ngroups; G += LIN_NW) {
            issue(G + (int64_t)(ST - 1) * LIN_NW, st == 0 ? ST - 1 : st - 1);   // refills the stage consumed one iteration ago
            asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory");
            consume(G, st);
            st = st + 1 == ST ? 0 : st + 1;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");

        // ---- combine the warps in a fixed order (shared memory), write the tile's moments ----
        double* U = lin_sm;
        double* S = lin_sm + DU * DU;
        __syncthreads();   // the previous tile's write-out has finished reading lin_sm
        for (int i = threadIdx.x; i < DU * DU + DU; i += LIN_THREADS) lin_sm[i] = 0.0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NB; ++k) {   // first moments: sum over the 4 sample lanes of a coordinate
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 1);
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 2);
        }
        for (int w = 0; w < LIN_NW; ++w) {
            if (warp == w) {
                int p = 0;
#pragma unroll
                for (int rb = 0; rb < NB; ++rb)
#pragma unroll
                    for (int cb = rb; cb < NB; ++cb, ++p) {
                        double* e = U + (8 * rb + g8) * DU + 8 * cb + 2 * t4;   // C fragment: row g, columns 2t, 2t+1
                        e[0] += acc[p][0];
                        e[1] += acc[p][1];
                    }
                if (t4 == 0) {
#pragma unroll
                    for (int k = 0; k < NB; ++k) S[8 * k + g8] += s1[k];
                }
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < DU * DU + DU; i += LIN_THREADS) out[i] = lin_sm[i];
        if (threadIdx.x < 8) out[DU * DU + DU + threadIdx.x] = threadIdx.x == 0 ? (double)cnt : 0.0;
    }
}

// moments of the tiles [t_begin, t_end) of each replicate -> one moment set per replicate of the launch (fixed order)
constexpr int LGR_X = 32, LGR_Y = 32;
__global__ void __launch_bounds__(LGR_X* LGR_Y) linear_gram_reduce_kernel(const double* __restrict__ gram, double* __restrict__ red, int gs, int tpr,
                                                                         int t_begin, int t_end, int rl_first) {
    __shared__ double sh[LGR_Y][LGR_X];
    const int i = blockIdx.x * LGR_X + threadIdx.x, rl = rl_first + blockIdx.y;
    const int j0 = max(t_begin, rl * tpr), j1 = min(t_end, (rl + 1) * tpr);
    double s = 0.0;
    if (i < gs) {
        int j = j0 + threadIdx.y;
        for (; j + 3 * LGR_Y < j1; j += 4 * LGR_Y) {
            const double v0 = gram[(size_t)j * gs + i], v1 = gram[(size_t)(j + LGR_Y) * gs + i];
            const double v2 = gram[(size_t)(j + 2 * LGR_Y) * gs + i], v3 = gram[(size_t)(j + 3 * LGR_Y) * gs + i];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; j < j1; j += LGR_Y) s += gram[(size_t)j * gs + i];
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && i < gs) {
        double t = 0.0;
        for (int y = 0; y < LGR_Y; ++y) t += sh[y][threadIdx.x];
        red[(size_t)blockIdx.y * gs + i] = t;
    }
}

// moments of a replicate -> standard record of every output, ADDED to partials[replicate][output].  Grid (replicates of the launch,
// groups of LREC_OB outputs); thread (output, j) owns the coordinates j, j + 16: every dot product is a chain of d FMAs, the 16
// lanes of an output are combined by a butterfly (fixed order).  Launches are serialised on the handle's stream and one thread
// owns each entry of a record, so the read-modify-write of `partials` is race-free.
//
// With fa = yA - y0 = w.a', fb = w.b', delta = b - a (a', b' the shifted rows) every term of every estimator is a moment form:
//     sum yB delta_k = y0 sd_k + hB_k,   hB_k = sum_l w_l (Sbb - Sba)[l][k]          sum yA delta_k = y0 sd_k + hA_k,   hA_k = sum_l w_l (Sab - Saa)[l][k]
//     V_k      (:192)  sum yB (yAB_k - yA)          =  w_k (y0 sd_k + hB_k)
//     Jansen   (:213)  sum (yA - yAB_k)^2            =  w_k^2 DD_kk,              DD = sum delta delta'
//     Sobol2007(:211)  sum yA (yA - yAB_k)           = -w_k (y0 sd_k + hA_k)
//     Homma/Janon (:207, :224)  sum yA yAB_k         =  sum yA^2 + w_k (y0 sd_k + hA_k)
//     Janon    (:219)  sum (yA^2 + yAB_k^2)          =  2 sum yA^2 + 2 w_k (y0 sd_k + hA_k) + w_k^2 DD_kk ;   (:220) sum (yA + yAB_k) = 2 sum yA + w_k sd_k
//     pairs    (:197)  sum (yBA_k yAB_j - yA yB)     =  w_j (y0 sd_j + hB_j) - w_k (y0 sd_k + hA_k) - w_k w_j DD_kj          (k < j)
struct LinearRecArgs {
    const double* red;   // [nrep_launch][gs]
    const double* A;     // the launch's A buffer (shift rows)
    const double* W;     // nout x d column-major
    double* partials;    // [nboot][nout][nstat]
    TileGeom g;
    int d, NB, nout, nstat, rl_first, est, second;
};
constexpr int LREC_OB = 16, LREC_J = 16;   // outputs per CTA, lanes per output (d <= 2 * LREC_J)
static_assert(LIN_DMAX <= 2 * LREC_J, "two coordinates per lane");
// G, H always; HA, DD, AA when the options need them (5 matrices of d x (d+1)); 4 vectors; Wt[d][OB]; hB, hA per (output, k)
__host__ __device__ inline size_t lin_rec_smem_bytes(int d) {
    return sizeof(double) * (5 * (size_t)d * (d + 1) + 4 * d + (size_t)d * LREC_OB + 2 * (size_t)LREC_OB * LIN_DMAX);
}

__device__ __forceinline__ dd dd_from(double x) { return dd{x, 0.0}; }
__device__ __forceinline__ dd dd_mul_d(dd a, double b) {  // a * b
    double p = a.hi * b;
    double e = __fma_rn(a.hi, b, -p);
    e = __fma_rn(a.lo, b, e);
    return quick_two_sum(p, e);
}
// shifted sums -> sum y (dd), sum y^2 (dd) of m values:  s' + m y0 ;  q' + 2 y0 s' + m y0^2
__device__ __forceinline__ void lin_unshift(double s1, double q2, double m, double y0, dd& S, dd& Q) {
    S = dd_from(s1);
    dd_add_dd(S, dd_mul_d(dd_from(m), y0));
    Q = dd_from(q2);
    dd_add_dd(Q, dd_mul_d(dd_from(2.0 * y0), s1));
    dd_add_dd(Q, dd_mul_d(dd_mul_d(dd_from(y0), y0), m));
}

__global__ void __launch_bounds__(LREC_OB* LREC_J) linear_records_kernel(LinearRecArgs a) {
    extern __shared__ double sg[];
    const int d = a.d, P = d + 1, DU = 8 * a.NB, gs = lin_gram_size(a.NB), est = a.est;
    const bool second = a.second != 0, need_a = second || est != GSA_EI_JANSEN1999;
    const bool need_ya = est == GSA_EI_HOMMA1996 || est == GSA_EI_JANON2014;
    const int rl = a.rl_first + blockIdx.x, r = a.g.rep_first + rl;
    const double* red = a.red + (size_t)blockIdx.x * gs;
    const double* S = red + DU * DU;
    const double m = red[DU * DU + DU];   // samples of the replicate in this launch
    if (m == 0.0) return;
    double* G = sg;              // Saa + Sbb
    double* H = G + d * P;       // Sbb - Sba
    double* HA = H + d * P;      // Sab - Saa
    double* DD = HA + d * P;     // sum delta delta'
    double* AA = DD + d * P;     // Saa
    double* g1 = AA + d * P;     // sa + sb
    double* sd = g1 + d;         // sb - sa
    double* sa = sd + d;
    double* c = sa + d;
    double* Wt = c + d;          // [d][LREC_OB]
    double* hBs = Wt + d * LREC_OB;             // [LREC_OB][LIN_DMAX]
    double* hAs = hBs + LREC_OB * LIN_DMAX;
    const int ol = threadIdx.x / LREC_J, j = threadIdx.x % LREC_J, o = blockIdx.y * LREC_OB + ol;
    const bool live = o < a.nout;
    // U is stored block-upper-triangular: entry (l, k) lives at [l][k] when block(l) <= block(k), else at [k][l]
    auto Us = [&](int l, int k) { return (l >> 3) <= (k >> 3) ? red[l * DU + k] : red[k * DU + l]; };
    const int64_t cs = lin_shift_col(a.g, r);
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) {
        const int l = e / d, k = e - l * d;
        const double saa = Us(l, k), sbb = Us(d + l, d + k), sba = Us(d + l, k), sab = Us(l, d + k);
        G[l * P + k] = saa + sbb;
        H[l * P + k] = sbb - sba;
        if (need_a) HA[l * P + k] = sab - saa;
        DD[l * P + k] = ((sbb - sba) - sab) + saa;
        if (need_ya) AA[l * P + k] = saa;
    }
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        g1[k] = S[k] + S[d + k];
        sd[k] = S[d + k] - S[k];
        sa[k] = S[k];
        c[k] = __ldg(a.A + (size_t)(cs - a.g.buf_col0) * d + k);
    }
    for (int e = threadIdx.x; e < d * LREC_OB; e += blockDim.x) {
        const int l = e / LREC_OB, oo = blockIdx.y * LREC_OB + (e % LREC_OB);
        Wt[e] = oo < a.nout ? __ldg(a.W + oo + (size_t)a.nout * l) : 0.0;
    }
    __syncthreads();
    // phase 1: y0 = w.c ; S' = w.g1 ; Q' = w' G w (and the same of yA alone) -- lane j adds the rows l = j, j + 16, then a butterfly
    double y0 = 0.0, S1 = 0.0, Q = 0.0, SA1 = 0.0, QA = 0.0;
    for (int l = j; l < d; l += LREC_J) {
        const double wl = Wt[l * LREC_OB + ol];
        double row = 0.0, rowa = 0.0;
        for (int k = 0; k < d; ++k) row = fma(Wt[k * LREC_OB + ol], G[l * P + k], row);
        if (need_ya)
            for (int k = 0; k < d; ++k) rowa = fma(Wt[k * LREC_OB + ol], AA[l * P + k], rowa);
        y0 = fma(wl, c[l], y0);
        S1 = fma(wl, g1[l], S1);
        Q = fma(wl, row, Q);
        SA1 = fma(wl, sa[l], SA1);
        QA = fma(wl, rowa, QA);
    }
#pragma unroll
    for (int x = 1; x < LREC_J; x <<= 1) {
        y0 += __shfl_xor_sync(0xffffffffu, y0, x);
        S1 += __shfl_xor_sync(0xffffffffu, S1, x);
        Q += __shfl_xor_sync(0xffffffffu, Q, x);
        SA1 += __shfl_xor_sync(0xffffffffu, SA1, x);
        QA += __shfl_xor_sync(0xffffffffu, QA, x);
    }
    double* rec = a.partials + ((size_t)r * a.nout + (live ? o : 0)) * a.nstat;
    dd sya{0.0, 0.0}, syya{0.0, 0.0};   // sum yA, sum yA^2 of this launch's samples
    if (need_ya) lin_unshift(SA1, QA, m, y0, sya, syya);
    if (live && j == 0) {
        dd sy, syy;
        lin_unshift(S1, Q, 2.0 * m, y0, sy, syy);
        dd p0{rec[0], rec[1]}, p1{rec[2], rec[3]};
        dd_add_dd(p0, sy);
        dd_add_dd(p1, syy);
        rec[0] = p0.hi; rec[1] = p0.lo; rec[2] = p1.hi; rec[3] = p1.lo;
        if (need_ya) {
            dd p2{rec[4], rec[5]}, p3{rec[6], rec[7]};
            dd_add_dd(p2, sya);
            dd_add_dd(p3, syya);
            rec[4] = p2.hi; rec[5] = p2.lo; rec[6] = p3.hi; rec[7] = p3.lo;
        }
    }
    // phase 2: the coordinates k = j, j + 16 of this output
    for (int k = j; k < d; k += LREC_J) {
        const double wk = Wt[k * LREC_OB + ol];
        double hb = 0.0, ha = 0.0;
        for (int l = 0; l < d; ++l) hb = fma(Wt[l * LREC_OB + ol], H[l * P + k], hb);
        if (need_a)
            for (int l = 0; l < d; ++l) ha = fma(Wt[l * LREC_OB + ol], HA[l * P + k], ha);
        const double yb_d = fma(y0, sd[k], hb), ya_d = fma(y0, sd[k], ha);   // sum yB delta_k, sum yA delta_k
        hBs[ol * LIN_DMAX + k] = yb_d;
        hAs[ol * LIN_DMAX + k] = ya_d;
        if (live) {
            const double dkk = wk * wk * DD[k * P + k];
            rec[stat_v(k)] += wk * yb_d;                                                    // :192
            if (est == GSA_EI_JANSEN1999) rec[stat_e(d, 0, k)] += dkk;                      // :213
            else if (est == GSA_EI_SOBOL2007) rec[stat_e(d, 0, k)] += -(wk * ya_d);         // :211
            else {
                const double syya_d = syya.hi + syya.lo;
                rec[stat_e(d, 0, k)] += syya_d + wk * ya_d;                                 // :207 / :224
                if (est == GSA_EI_JANON2014) {
                    rec[stat_e(d, 1, k)] += (2.0 * syya_d + 2.0 * (wk * ya_d)) + dkk;       // :219
                    rec[stat_e(d, 2, k)] += 2.0 * (sya.hi + sya.lo) + wk * sd[k];           // :220
                }
            }
        }
    }
    if (!second) return;
    __syncthreads();
    // phase 3: the pairs k < jj of this output, 16 lanes take them round-robin
    if (live) {
        const int pbase = stat_pair_base(d, est), npairs = d * (d - 1) / 2;
        int k = 0, first = 0;   // pairs of row k start at index `first`
        for (int p = j; p < npairs; p += LREC_J) {
            while (p >= first + (d - 1 - k)) { first += d - 1 - k; ++k; }
            const int jj = k + 1 + (p - first);
            const double wk = Wt[k * LREC_OB + ol], wj = Wt[jj * LREC_OB + ol];
            rec[pbase + p] += (wj * hBs[ol * LIN_DMAX + jj] - wk * hAs[ol * LIN_DMAX + k]) - (wk * wj) * DD[k * P + jj];   // :197
        }
    }
}
#endif  // __CUDACC__

}  // namespace gsa
