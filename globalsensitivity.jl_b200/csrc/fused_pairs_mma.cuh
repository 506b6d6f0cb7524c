// fused_pairs_mma.cuh -- SECOND-order Sobol sums with the pair terms on the FP64 tensor-core path, d <= 32.
//
// The second-order estimator needs, for every pair k < j,  sum_s (yBA_k(s) * yAB_j(s) - yA(s) yB(s))  (reference
// src/sobol_sensitivity.jl:197): d(d-1)/2 running sums per sample.  Over the samples this is a dense contraction,
//     P[k][j] = sum_s yBA_k(s) * yAB_j(s)        (a d x d matrix product with the SAMPLES as the contraction index),
// so it runs as `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4): a warp step covers 4 samples, thread (g = lane/4, t = lane%4) owns
// coordinate 8*blk+g of sample t for every 8-coordinate block blk < NB; its yBA value is the A fragment (row g, k = t) and its
// yAB value the B fragment (k = t, column g) of the block pairs rb <= cb.  One DMMA replaces 64 (diagonal blocks: 28 useful)
// multiply-adds AND the shuffles that a lane-per-coordinate layout would need to bring yBA_k to the lane of yAB_j.
//
// The products are taken of DIFFERENCES, u_k = yBA_k - yB and w_j = yAB_j - yA (w_j is the difference the first-order term
// needs anyway, :192), so the sums stay of the order of the output's spread, not its mean, and no separate shift is carried:
//     sum (yBA_k yAB_j - yA yB) = sum yB w_j + sum yA u_k + sum u_k w_j = V_j + U_k + P_kj
// with V_j the first-order numerator itself, U_k one more running sum per coordinate and P = sum u w^T the mma product.
// (Round 1 shifted both operands by a constant c and carried four correction sums; this form needs 7 fewer FP64 instructions per
// warp step, and the FP64 pipe -- which DMMA shares with DFMA, profiles/r2_dmma_peak.jsonl -- is what bounds these kernels.)
// The first- and total-order terms are formed per sample exactly as the reference forms them (difference first, :192, :213;
// Sobol2007 :211 and Homma1996 :207 are one-accumulator estimators too and share the kernel; Janon2014 needs three terms).
// sum y and sum y^2 (the variance) are accumulated shifted by c = yA of the tile's first sample and un-shifted in double-double.
//
// Three sources feed the same engine (template parameter SRC):
//   * PRODUCT -- product-form models (Sobol G), 8 < d <= 32: lane (g, t) loads coordinate 8*blk+g of rows A_s, B_s (the 8 lanes
//     of a sample read 64 contiguous bytes), evaluates its factors, and one xor butterfly over the 8 lanes of the sample gives it
//     the product of all OTHER factors (no division: a factor may be 0), hence yAB_i = gB_i * othersA_i, yBA_i = gA_i * othersB_i;
//   * PRODUCT_TILE -- the same models for even d <= 8 (BASELINE config 3): with a single block the butterfly would dominate, so
//     lane L evaluates ALL 2d+2 values of sample L of a 32-sample chunk from prefix/suffix products (no shuffles), writes them
//     to a warp-private tile, and the mma fragments are read back from the tile;
//   * Y -- a precomputed Y = f(all_points) (estimator-only entry): every warp streams chunks of CH = 16 or 32 samples of all 2d+2
//     blocks into such a tile (128 / 256 contiguous bytes per block; 16-byte copies that bypass L1 when the blocks are 16-byte
//     aligned) and reads the fragments from there.
// Operands of future steps are in flight as cp.async copies (per-thread ring / per-warp tiles): no block-wide barrier.
#pragma once
#include <type_traits>
#include "gsa_common.cuh"

namespace gsa {

constexpr int PMMA_THREADS = 128;
constexpr int PMMA_NW = PMMA_THREADS / 32;
constexpr int PMMA_DMAX = 32;

struct PairsArgs {
    // product source
    const double* A;
    const double* B;
    const double2* coef;  // [d] (c1_k, c0_k): g_k(x) = |x - 0.5| * c1_k + c0_k
    // Y source: element (o, col) at Y[o + nout*col], blocks [A | B | AB_1..AB_d | BA_1..BA_d] of n columns per replicate
    const double* Y;
    int64_t n;
    int nout;
    double* tile_rec;     // [ntiles][nout][nstat]
    TileGeom g;
    int d, nstat, est, t_begin, t_end;
};

#ifdef __CUDACC__
__device__ __forceinline__ void pm_cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void pm_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Staging.  Product source: every thread keeps ST groups of UN steps of ITS OWN operands (xa, xb per block) in flight as
// cp.async copies into a private ring (no barrier).  Y source: every WARP keeps ST chunks of CH samples in flight in a
// warp-private tile [block][CH] whose 4-sample groups are xor-swizzled with the block number (pm_yidx), so that the copies (a
// half-warp writes a permutation of one row) and the fragment reads (lanes (g, t) of a half-warp hit 16 distinct 8-byte bank
// pairs) are both conflict-free without padding; only __syncwarp() separates the copies from the reads.  Small chunks and many
// stages: the data a warp waits for was requested (ST-1) chunks earlier, and 8-9 warps per SM fit (the round-1 tile -- 32
// samples, pitch 36, two stages -- spent most of its time waiting for the chunk requested one step earlier: top stall long
// scoreboard at the wait_group, profiles/r2_prof_stalls.txt).
constexpr int PM_CH = 32, PM_PITCH = 36;
template <int NB, int UN, int ST>
__host__ __device__ constexpr size_t pm_ring_doubles_product() { return (size_t)ST * UN * 2 * NB * PMMA_THREADS; }
inline size_t pm_ring_doubles_y(int d, int ST, int CH) { return (size_t)PMMA_NW * ST * (2 + 2 * d) * CH; }
__host__ __device__ __forceinline__ int pm_yidx(int row, int sl, int CH) { return row * CH + (sl ^ ((row & 3) << 2)); }
// thread-per-sample product source (d <= 8, even): per warp ST stages of the chunk's raw A and B rows + one tile of model values
__host__ __device__ inline size_t pm_ring_doubles_tile(int d, int ST) { return (size_t)PMMA_NW * ((size_t)ST * 2 * PM_CH * d + (2 + 2 * d) * PM_PITCH); }
inline size_t pm_smem_bytes(size_t ring_doubles, int NB) { return sizeof(double) * (ring_doubles + 64 * NB * NB + 4 * 8 * NB + 8); }

// NB = 8-coordinate blocks (d <= 8*NB); UN = warp steps per load group (PRODUCT), d itself (PRODUCT_TILE) or the samples per
// chunk (Y: 16 or 32); ST = stages in flight
enum { PM_SRC_PRODUCT = 0, PM_SRC_Y = 1, PM_SRC_PRODUCT_TILE = 2 };
// CTAs per SM the register allocation aims at: what the shared memory of each source leaves room for
constexpr int pm_min_ctas(int SRC, int NB) { return SRC == PM_SRC_PRODUCT ? (NB <= 3 ? 4 : 1) : (SRC == PM_SRC_PRODUCT_TILE || NB == 1) ? 3 : 2; }
template <int SRC, int NB, int UN, int ST, bool JANSEN = true>
__global__ void __launch_bounds__(PMMA_THREADS, pm_min_ctas(SRC, NB)) pairs_mma_kernel(PairsArgs a) {
    constexpr bool YSRC = SRC == PM_SRC_Y, TSRC = SRC == PM_SRC_PRODUCT_TILE;
    static_assert(!TSRC || NB == 1, "thread-per-sample evaluation: one 8-coordinate block (UN carries d)");
    constexpr int DU = 8 * NB, NP = NB * (NB + 1) / 2;
    extern __shared__ __align__(16) double pm_sm[];
    const int d = a.d, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int nblk = 2 + 2 * d;                                   // Y blocks per sample: yA, yB, yAB_1..d, yBA_1..d
    constexpr int YCH = YSRC ? UN : PM_CH;                        // Y source: samples per chunk
    static_assert(!YSRC || YCH == 16 || YCH == 32, "Y source: chunks of 16 or 32 samples");
    const size_t ring_doubles = YSRC ? (size_t)PMMA_NW * ST * nblk * YCH
                              : TSRC ? pm_ring_doubles_tile(d, ST) : pm_ring_doubles_product<NB, UN, ST>();
    double* ring = pm_sm + threadIdx.x;                          // product: [ST][UN][2 NB][THREADS]
    double* wtile = pm_sm + (size_t)warp * ST * nblk * YCH;      // Y: [ST][nblk][YCH] per warp, swizzled (pm_yidx)
    double* Csm = pm_sm + ring_doubles;                          // [DU][DU]
    double* Vsm = Csm + DU * DU;                                 // V[DU], E[DU], U[DU], (unused [DU]), sums[8]
    bool ok[NB];
    double2 cf[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        ok[k] = 8 * k + g8 < d;
        cf[k] = (!YSRC && ok[k]) ? __ldg(a.coef + 8 * k + g8) : make_double2(0.0, 1.0);   // padded coordinates: factor 1
    }
    const bool sobol2007 = a.est == GSA_EI_SOBOL2007;   // JANSEN = false: Sobol2007 or Homma1996 (block-uniform)
    const int units = (a.t_end - a.t_begin) * (YSRC ? a.nout : 1);
    (void)ring; (void)wtile;

    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int tile = a.t_begin + (YSRC ? u / a.nout : u), o = YSRC ? u % a.nout : 0;
        int64_t c0, c1;
        tile_range(a.g, tile, c0, c1);
        double* rec = a.tile_rec + ((size_t)tile * (YSRC ? a.nout : 1) + o) * a.nstat;
        if (c1 <= c0) {
            for (int i = threadIdx.x; i < a.nstat; i += PMMA_THREADS) rec[i] = 0.0;
            continue;
        }
        const int r = a.g.rep_first + tile / a.g.tpr;
        const int64_t cnt = c1 - c0;
        // source geometry: product -> row index inside the A/B buffers; Y -> column of block 0 (yA) of this replicate
        const int64_t row0 = c0 - a.g.buf_col0;
        const int64_t ycol0 = (int64_t)r * (2 * d + 2) * a.n + (c0 - (int64_t)r * a.g.n);
        const double* ybase = YSRC ? a.Y + o + (size_t)a.nout * ycol0 : nullptr;
        const int64_t ybs = (int64_t)a.nout * a.n;   // Y block stride (doubles)

        // ---- the shift: yA of the tile's first sample, bit-identical in every thread ----
        double c;
        if constexpr (YSRC) {
            c = __ldg(ybase);
        } else {
            double la = 1.0;
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const double x = ok[k] ? __ldg(a.A + (size_t)row0 * d + 8 * k + g8) : 0.5;
                la *= fma(fabs(x - 0.5), cf[k].x, cf[k].y);
            }
            la *= __shfl_xor_sync(0xffffffffu, la, 4);
            la *= __shfl_xor_sync(0xffffffffu, la, 8);
            la *= __shfl_xor_sync(0xffffffffu, la, 16);
            c = la;
        }

        double V[NB], E[NB], U[NB], acc[NP][2];
#pragma unroll
        for (int k = 0; k < NB; ++k) V[k] = E[k] = U[k] = 0.0;
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p][0] = acc[p][1] = 0.0;
        double sfa = 0.0, sfb = 0.0, qfa = 0.0, qfb = 0.0;   // sums of fa' = yA - c, fb' = yB - c, fa'^2, fb'^2 (this lane's samples)

        // one warp step: this lane's sample contributes yA, yB and, per block, yAB_i / yBA_i of coordinate i = 8k + g
        auto sample_sums = [&](bool valid, double yA, double yB) {
            const double fa = valid ? yA - c : 0.0, fb = valid ? yB - c : 0.0;
            sfa += fa; sfb += fb;
            qfa = fma(fa, fa, qfa); qfb = fma(fb, fb, qfb);
        };
        // full_c = std::true_type: every sample of the step is inside the tile -- no validity selects at all.  The lanes of padded
        // coordinates then carry finite junk (their operands are valid rows or the constant 0.5 with factor 1) into sums and
        // accumulator rows / columns that are never read (i >= d).
        auto step = [&](auto full_c, bool valid, double yA, double yB, const double (&yab)[NB], const double (&yba)[NB]) {
            constexpr bool FULL = decltype(full_c)::value;
            if constexpr (!TSRC) sample_sums(FULL || valid, yA, yB);   // (tile source: added once per sample by the lane that evaluated it)
            double pv[NB], qv[NB];
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                const bool lv = FULL || (valid && ok[k]);
                const double df = lv ? yab[k] - yA : 0.0;
                V[k] = fma(yB, df, V[k]);            // :192
                // total-order term: (yA - yAB)^2 (Jansen1999, :213), yA (yA - yAB) (Sobol2007, :211) or yA yAB (Homma1996, :207)
                if constexpr (JANSEN) E[k] = fma(df, df, E[k]);
                else E[k] = fma(lv ? yA : 0.0, sobol2007 ? -df : yab[k], E[k]);
                pv[k] = lv ? yba[k] - yB : 0.0;
                qv[k] = df;
                U[k] = fma(yA, pv[k], U[k]);
            }
            int p = 0;
#pragma unroll
            for (int rb = 0; rb < NB; ++rb)
#pragma unroll
                for (int cb = rb; cb < NB; ++cb, ++p) pm_dmma(acc[p][0], acc[p][1], pv[rb], qv[cb]);   // :197
        };

        if constexpr (YSRC) {
            // ---- Y source: warp-private tiles of YCH samples, the warps take chunks round-robin ----
            const int64_t nchunks = (cnt + YCH - 1) / YCH;
            int rab[NB], rba[NB];   // tile rows of this lane's yAB / yBA values
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                rab[k] = ok[k] ? 2 + 8 * k + g8 : 0;
                rba[k] = ok[k] ? 2 + d + 8 * k + g8 : 0;
            }
            // 16-byte copies (two samples of a block, around L1) when every block of this tile starts 16-byte aligned
            const bool wide = a.nout == 1 && (a.n & 1) == 0 && (reinterpret_cast<uintptr_t>(ybase) & 15) == 0;
            auto issue = [&](int64_t ch, int st) {
                if (ch < nchunks) {
                    double* dst = wtile + (size_t)st * nblk * YCH;
                    if (wide) {
                        constexpr int LPR = YCH / 2, RPI = 32 / LPR;          // lanes per row, rows per instruction
                        const int sl = 2 * (lane % LPR);
                        const int64_t s = ch * YCH + sl;
                        const double* yp = ybase + s;
                        if (s + 1 < cnt) {
#pragma unroll 4
                            for (int b = lane / LPR; b < nblk; b += RPI)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst + pm_yidx(b, sl, YCH))), "l"(yp + ybs * b) : "memory");
                        } else if (s < cnt) {
                            for (int b = lane / LPR; b < nblk; b += RPI) pm_cp_async8(dst + pm_yidx(b, sl, YCH), yp + ybs * b);
                        }
                    } else {
                        constexpr int RPI = 32 / YCH;
                        const int sl = lane % YCH;
                        const int64_t s = ch * YCH + sl;
                        if (s < cnt) {
                            const double* yp = ybase + (size_t)a.nout * s;
#pragma unroll 6
                            for (int b = lane / YCH; b < nblk; b += RPI) pm_cp_async8(dst + pm_yidx(b, sl, YCH), yp + ybs * b);
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
#pragma unroll
            for (int st = 0; st < ST - 1; ++st) issue(warp + (int64_t)st * PMMA_NW, st);
            int st = 0;
            for (int64_t ch = warp; ch < nchunks; ch += PMMA_NW) {
                __syncwarp();   // every lane has finished reading the stage that is refilled now
                issue(ch + (int64_t)(ST - 1) * PMMA_NW, st == 0 ? ST - 1 : st - 1);
                asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory");
                __syncwarp();   // the other lanes' copies of this chunk are visible
                const double* tl = wtile + (size_t)st * nblk * YCH;
                if ((ch + 1) * YCH <= cnt) {   // a full chunk: no selects (padded coordinates read row 0)
#pragma unroll 2
                    for (int q = 0; q < YCH / 4; ++q) {
                        const int sl = 4 * q + t4;
                        double yab[NB], yba[NB];
#pragma unroll
                        for (int k = 0; k < NB; ++k) {
                            yab[k] = tl[pm_yidx(rab[k], sl, YCH)];
                            yba[k] = tl[pm_yidx(rba[k], sl, YCH)];
                        }
                        step(std::true_type{}, true, tl[pm_yidx(0, sl, YCH)], tl[pm_yidx(1, sl, YCH)], yab, yba);
                    }
                } else {
#pragma unroll 2
                    for (int q = 0; q < YCH / 4; ++q) {
                        const int sl = 4 * q + t4;
                        const bool valid = ch * YCH + sl < cnt;
                        double yab[NB], yba[NB];
                        const double yA = valid ? tl[pm_yidx(0, sl, YCH)] : c, yB = valid ? tl[pm_yidx(1, sl, YCH)] : c;
#pragma unroll
                        for (int k = 0; k < NB; ++k) {
                            const bool lv = valid && ok[k];
                            yab[k] = lv ? tl[pm_yidx(rab[k], sl, YCH)] : yA;
                            yba[k] = lv ? tl[pm_yidx(rba[k], sl, YCH)] : c;
                        }
                        step(std::false_type{}, valid, yA, yB, yab, yba);
                    }
                }
                st = st + 1 == ST ? 0 : st + 1;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else if constexpr (TSRC) {
            // ---- product source, thread-per-sample evaluation (even d = TD <= 8, the UN template slot): lane L evaluates sample L of a
            //      32-sample chunk -- all 2d+2 model values from prefix/suffix products, no shuffles -- and writes them to the warp's
            //      tile; the mma fragments are read back from there.  The chunk's raw rows (32*d contiguous doubles of A and of B)
            //      arrive by 16-byte cp.async copies, fully coalesced, ST chunks ahead; pieces are xor-swizzled so that a lane reads
            //      its own row conflict-free.  Full chunks take a path without any validity selects. ----
            constexpr int TD = UN, HP = TD / 2;                      // 16-byte pieces per row
            static_assert(TD % 2 == 0 && TD >= 2 && TD <= 8, "tile source: even d <= 8");
            const int gi = g8 < TD ? g8 : 0;                        // lanes of padded coordinates read a valid row and are masked
            const int64_t nchunks = (cnt + PM_CH - 1) / PM_CH;
            double* wbase = pm_sm + (size_t)warp * ((size_t)ST * 2 * PM_CH * TD + (2 + 2 * TD) * PM_PITCH);
            double* ytile = wbase + (size_t)ST * 2 * PM_CH * TD;
            auto swz = [](int row, int j) { return HP == 4 ? (j ^ ((row >> 1) & 3)) : HP == 2 ? (j ^ ((row >> 2) & 1)) : j; };
            auto issue = [&](int64_t ch, int st) {
                if (ch < nchunks) {
                    const int rows = (int)min((int64_t)PM_CH, cnt - ch * PM_CH);
                    const double* ga = a.A + (size_t)(row0 + ch * PM_CH) * TD;
                    const double* gb = a.B + (size_t)(row0 + ch * PM_CH) * TD;
                    double* sa = wbase + (size_t)st * 2 * PM_CH * TD;
                    double* sb = sa + PM_CH * TD;
#pragma unroll
                    for (int i = 0; i < HP; ++i) {
                        const int p = i * 32 + lane, row = p / HP, j = p % HP, q = 2 * (row * HP + swz(row, j));
                        if (row < rows) {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sa + q)), "l"(ga + 2 * p) : "memory");
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sb + q)), "l"(gb + 2 * p) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            double2 cfd[TD];
#pragma unroll
            for (int i = 0; i < TD; ++i) cfd[i] = __ldg(a.coef + i);
            auto phase_a = [&](int st, bool valid) {   // this lane's sample
                const double* sa = wbase + (size_t)st * 2 * PM_CH * TD;
                const double* sb = sa + PM_CH * TD;
                double ga[TD], gb[TD];
#pragma unroll
                for (int j = 0; j < HP; ++j) {
                    const int q = 2 * (lane * HP + swz(lane, j));
                    double2 va = *reinterpret_cast<const double2*>(sa + q), vb = *reinterpret_cast<const double2*>(sb + q);
                    ga[2 * j] = fma(fabs(va.x - 0.5), cfd[2 * j].x, cfd[2 * j].y);
                    ga[2 * j + 1] = fma(fabs(va.y - 0.5), cfd[2 * j + 1].x, cfd[2 * j + 1].y);
                    gb[2 * j] = fma(fabs(vb.x - 0.5), cfd[2 * j].x, cfd[2 * j].y);
                    gb[2 * j + 1] = fma(fabs(vb.y - 0.5), cfd[2 * j + 1].x, cfd[2 * j + 1].y);
                }
                double PA[TD + 1], PB[TD + 1];
                PA[0] = PB[0] = 1.0;
#pragma unroll
                for (int i = 0; i < TD; ++i) { PA[i + 1] = PA[i] * ga[i]; PB[i + 1] = PB[i] * gb[i]; }
                double* yt = ytile + lane;
                yt[0] = PA[TD];
                yt[PM_PITCH] = PB[TD];
                double ra = 1.0, rb = 1.0;
#pragma unroll
                for (int k = TD - 1; k >= 0; --k) {
                    yt[(2 + k) * PM_PITCH] = PA[k] * (gb[k] * ra);        // AB_k: A with coordinate k from B (:96-99)
                    yt[(2 + TD + k) * PM_PITCH] = PB[k] * (ga[k] * rb);   // BA_k (:100-104)
                    ra *= ga[k];
                    rb *= gb[k];
                }
                sample_sums(valid, PA[TD], PB[TD]);
            };
#pragma unroll
            for (int st = 0; st < ST - 1; ++st) issue(warp + (int64_t)st * PMMA_NW, st);
            int st = 0;
            for (int64_t ch = warp; ch < nchunks; ch += PMMA_NW) {
                __syncwarp();   // the previous chunk's fragments have been read; the stage refilled now is free
                issue(ch + (int64_t)(ST - 1) * PMMA_NW, st == 0 ? ST - 1 : st - 1);
                asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory");
                __syncwarp();
                const bool full = (ch + 1) * PM_CH <= cnt;
                if (full) {
                    phase_a(st, true);
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < PM_CH / 4; ++q) {   // the same 32 samples as 8 mma steps (padded lanes are masked by ok[0])
                        const int sl = 4 * q + t4;
                        double yab[1], yba[1];
                        yab[0] = ytile[(2 + gi) * PM_PITCH + sl];
                        yba[0] = ytile[(2 + TD + gi) * PM_PITCH + sl];
                        step(std::true_type{}, true, ytile[sl], ytile[PM_PITCH + sl], yab, yba);
                    }
                } else {
                    phase_a(st, ch * PM_CH + lane < cnt);   // (lanes beyond the tile read stale rows: finite or not, they are masked)
                    __syncwarp();
#pragma unroll 2
                    for (int q = 0; q < PM_CH / 4; ++q) {
                        const int sl = 4 * q + t4;
                        const bool valid = ch * PM_CH + sl < cnt;
                        const bool lv = valid && g8 < TD;
                        const double yA = valid ? ytile[sl] : c, yB = valid ? ytile[PM_PITCH + sl] : c;   // (0 * NaN would poison V)
                        double yab[1], yba[1];
                        yab[0] = lv ? ytile[(2 + gi) * PM_PITCH + sl] : yA;
                        yba[0] = lv ? ytile[(2 + TD + gi) * PM_PITCH + sl] : c;
                        step(std::false_type{}, valid, yA, yB, yab, yba);
                    }
                }
                st = st + 1 == ST ? 0 : st + 1;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            // ---- product source: per-thread ring, sample 4j + t of a group of 4*UN samples ----
            const int64_t ngroups = (cnt + 4 * UN - 1) / (4 * UN);
            auto slot_of = [&](int st, int j, int v) -> double* { return ring + (size_t)((st * UN + j) * 2 * NB + v) * PMMA_THREADS; };
            auto issue = [&](int64_t G, int st) {
                if (G < ngroups) {
                    const int64_t s0 = G * (4 * UN) + t4;
#pragma unroll
                    for (int j = 0; j < UN; ++j) {
                        const int64_t s = s0 + 4 * j;
                        if (s < cnt) {
#pragma unroll
                            for (int k = 0; k < NB; ++k)
                                if (ok[k]) {
                                    const size_t e = (size_t)(row0 + s) * d + 8 * k + g8;
                                    pm_cp_async8(slot_of(st, j, 2 * k), a.A + e);
                                    pm_cp_async8(slot_of(st, j, 2 * k + 1), a.B + e);
                                }
                        }
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            // (the slots of padded coordinates are never copied into: they hold 0.5, and the factor there is the constant 1)
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (!ok[k])
                    for (int q = 0; q < ST * UN; ++q) { *slot_of(0, q, 2 * k) = 0.5; *slot_of(0, q, 2 * k + 1) = 0.5; }
            auto consume = [&](auto full_c, int64_t G, int st) {
                constexpr bool FULL = decltype(full_c)::value;
#pragma unroll
                for (int j = 0; j < UN; ++j) {
                    const bool valid = FULL || G * (4 * UN) + t4 + 4 * j < cnt;
                    double ga[NB], gb[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k) {
                        const bool lv = FULL || (valid && ok[k]);
                        const double xa = lv ? *slot_of(st, j, 2 * k) : 0.5, xb = lv ? *slot_of(st, j, 2 * k + 1) : 0.5;
                        ga[k] = fma(fabs(xa - 0.5), cf[k].x, cf[k].y);
                        gb[k] = fma(fabs(xb - 0.5), cf[k].x, cf[k].y);
                    }
                    // products of this thread's OTHER factors (prefix * suffix), then of the other 7 lanes of the sample (xor butterfly)
                    double ea[NB], eb[NB];
                    double pa = 1.0, pb = 1.0;
#pragma unroll
                    for (int k = 0; k < NB; ++k) { ea[k] = pa; eb[k] = pb; pa *= ga[k]; pb *= gb[k]; }
                    double sa = 1.0, sb = 1.0;
#pragma unroll
                    for (int k = NB - 1; k >= 0; --k) { ea[k] *= sa; eb[k] *= sb; sa *= ga[k]; sb *= gb[k]; }
                    double oa = 1.0, ob = 1.0;   // pa, pb = this lane's full product; oa, ob = product over the other lanes
#pragma unroll
                    for (int x = 4; x < 32; x <<= 1) {
                        const double ta = __shfl_xor_sync(0xffffffffu, pa, x), tb = __shfl_xor_sync(0xffffffffu, pb, x);
                        oa *= ta; pa *= ta;
                        ob *= tb; pb *= tb;
                    }
                    double yab[NB], yba[NB];
#pragma unroll
                    for (int k = 0; k < NB; ++k) {
                        yab[k] = gb[k] * (ea[k] * oa);   // A with coordinate i from B (:96-99)
                        yba[k] = ga[k] * (eb[k] * ob);   // B with coordinate i from A (:100-104)
                    }
                    step(full_c, valid, pa, pb, yab, yba);
                }
            };
#pragma unroll
            for (int st = 0; st < ST - 1; ++st) issue(warp + (int64_t)st * PMMA_NW, st);
            int st = 0;
            for (int64_t G = warp; G < ngroups; G += PMMA_NW) {
                issue(G + (int64_t)(ST - 1) * PMMA_NW, st == 0 ? ST - 1 : st - 1);
                asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory");
                if ((G + 1) * (4 * UN) <= cnt) consume(std::true_type{}, G, st);
                else consume(std::false_type{}, G, st);
                st = st + 1 == ST ? 0 : st + 1;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }

        // ---- flush: lanes of a coordinate (xor 1, 2), then the warps in a fixed order through shared memory ----
#pragma unroll
        for (int x = 1; x < 4; x <<= 1) {
#pragma unroll
            for (int k = 0; k < NB; ++k) {
                V[k] += __shfl_xor_sync(0xffffffffu, V[k], x);
                E[k] += __shfl_xor_sync(0xffffffffu, E[k], x);
                U[k] += __shfl_xor_sync(0xffffffffu, U[k], x);
            }
            sfa += __shfl_xor_sync(0xffffffffu, sfa, x); sfb += __shfl_xor_sync(0xffffffffu, sfb, x);
            qfa += __shfl_xor_sync(0xffffffffu, qfa, x); qfb += __shfl_xor_sync(0xffffffffu, qfb, x);
        }
        if constexpr (TSRC) {   // every lane added its own samples: finish the reduction over the 8 lane groups
#pragma unroll
            for (int x = 4; x < 32; x <<= 1) {
                sfa += __shfl_xor_sync(0xffffffffu, sfa, x); sfb += __shfl_xor_sync(0xffffffffu, sfb, x);
                qfa += __shfl_xor_sync(0xffffffffu, qfa, x); qfb += __shfl_xor_sync(0xffffffffu, qfb, x);
            }
        }
        __syncthreads();   // nobody still reads the shared arrays of the previous tile
        for (int i = threadIdx.x; i < DU * DU + 4 * DU + 8; i += PMMA_THREADS) Csm[i] = 0.0;
        __syncthreads();
        for (int w = 0; w < PMMA_NW; ++w) {
            if (warp == w) {
                int p = 0;
#pragma unroll
                for (int rb = 0; rb < NB; ++rb)
#pragma unroll
                    for (int cb = rb; cb < NB; ++cb, ++p) {
                        double* e = Csm + (8 * rb + g8) * DU + 8 * cb + 2 * t4;   // C fragment: row g, columns 2t, 2t+1
                        e[0] += acc[p][0];
                        e[1] += acc[p][1];
                    }
                if (t4 == 0) {
#pragma unroll
                    for (int k = 0; k < NB; ++k) {
                        const int i = 8 * k + g8;
                        Vsm[i] += V[k]; Vsm[DU + i] += E[k]; Vsm[2 * DU + i] += U[k];
                    }
                }
                if (lane == 0) {
                    double* q = Vsm + 4 * DU;
                    q[0] += sfa; q[1] += sfb; q[2] += qfa; q[3] += qfb;
                }
            }
            __syncthreads();
        }
        {
            const double* q = Vsm + 4 * DU;
            const int pbase = stat_pair_base(d, a.est);
            for (int i = threadIdx.x; i < d; i += PMMA_THREADS) {
                rec[stat_v(i)] = Vsm[i];
                rec[stat_e(d, 0, i)] = Vsm[DU + i];
            }
            for (int e = threadIdx.x; e < d * d; e += PMMA_THREADS) {
                const int k = e / d, j = e - k * d;
                if (k < j) rec[pbase + pair_index(d, k, j)] = Csm[k * DU + j] + (Vsm[j] + Vsm[2 * DU + k]);   // P_kj + V_j + U_k
            }
            if (threadIdx.x == 0) {
                // un-shift exactly (double-double): sum y = s' + m c ; sum y^2 = q' + 2 c s' + m c^2
                auto unshift = [&](double s1, double q2, double m, dd& S, dd& Q) {
                    double p = m * c, pe = __fma_rn(m, c, -p);
                    S = dd{s1, 0.0};
                    dd_add_dd(S, dd{p, pe});
                    double c2 = c * c, c2e = __fma_rn(c, c, -c2);
                    double m1 = m * c2, m1e = __fma_rn(m, c2, -m1);
                    m1e = __fma_rn(m, c2e, m1e);
                    double m2 = 2.0 * c * s1, m2e = __fma_rn(2.0 * c, s1, -m2);
                    Q = dd{q2, 0.0};
                    dd_add_dd(Q, quick_two_sum(m1, m1e));
                    dd_add_dd(Q, quick_two_sum(m2, m2e));
                };
                dd S, Q, SA, QA;
                dd sab2 = two_sum(q[0], q[1]), qab2 = two_sum(q[2], q[3]);
                // [yA; yB]: shifted sums added first (in double-double), then un-shifted with 2m values
                {
                    const double m = 2.0 * (double)cnt;
                    double p = m * c, pe = __fma_rn(m, c, -p);
                    S = sab2;
                    dd_add_dd(S, dd{p, pe});
                    double c2 = c * c, c2e = __fma_rn(c, c, -c2);
                    double m1 = m * c2, m1e = __fma_rn(m, c2, -m1);
                    m1e = __fma_rn(m, c2e, m1e);
                    double m2 = 2.0 * c * sab2.hi, m2e = __fma_rn(2.0 * c, sab2.hi, -m2);
                    m2e = __fma_rn(2.0 * c, sab2.lo, m2e);
                    Q = qab2;
                    dd_add_dd(Q, quick_two_sum(m1, m1e));
                    dd_add_dd(Q, quick_two_sum(m2, m2e));
                }
                unshift(q[0], q[2], (double)cnt, SA, QA);
                rec[0] = S.hi; rec[1] = S.lo; rec[2] = Q.hi; rec[3] = Q.lo;
                rec[4] = SA.hi; rec[5] = SA.lo; rec[6] = QA.hi; rec[7] = QA.lo;
            }
        }
        // (no fused tail here: carrying it cost the main loop 8 % on config 3 -- 98 -> 107 us -- for a ~10 us saving on the host;
        //  stage 2 and the exchange follow as launches of their own)
    }
}
#endif  // __CUDACC__

}  // namespace gsa
