// fused_small.cuh -- thread-per-sample kernels for SMALL d (compile-time D), everything in registers.
//
// One thread owns one sample: it obtains the sample's STEP = D+2 (or 2D+2) model values y[] from a "source"
//   * YSource        : reads them from a precomputed Y = f(all_points)      (K3, estimator-only; coalesced, since for
//                      a fixed block b consecutive samples are consecutive addresses)
//   * IshigamiSource : reads the sample's A and B rows and evaluates the Ishigami function, sharing sin/pow4 terms
//                      between the 2D+2 evaluations (the swapped designs AB_k / BA_k exist only as register selects)
//   * SobolGSource   : same for the Sobol G-function with prefix/suffix products
// and accumulates the estimator terms of reference src/sobol_sensitivity.jl:192 (V_k), :197 (V_kj), :211-235 (E_k)
// in per-thread registers over a grid-stride loop.  Accumulators are combined once per tile in a fixed order
// (shared-memory transpose, 16 accumulators per pass).  No atomics, bitwise reproducible.
#pragma once
#include "gsa_common.cuh"
#include "models.cuh"

namespace gsa {

constexpr int SMALL_THREADS = 256;
constexpr int SMALL_CH = 16;  // accumulators reduced per shared-memory pass

template <int D, bool SECOND>
struct SmallShape {
    static constexpr int STEP = SECOND ? 2 * D + 2 : D + 2;
    static constexpr int NPAIR = SECOND ? D * (D - 1) / 2 : 0;
};

// ---- per-thread asynchronous staging (cp.async / LDGSTS) ---------------------------------------------------------------
// Every thread copies the data of ITS OWN future samples into a private slot of a shared-memory ring and later reads only
// that slot back, so the pipeline needs no barrier at all -- only the per-thread cp.async group counter.  This keeps
// SMALL_NST-1 samples per thread in flight without spending registers on them (the register version of this kernel was
// latency-bound at 8 warps/SM: 59-65 % long-scoreboard stalls, profiles/r1_notes.md).
constexpr int SMALL_NST = 3;

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- sources -------------------------------------------------------------------------------------------------------
// issue(st, ...) : start the asynchronous copy of one sample's raw data into the thread's slot of stage `st`
// fetch(st, y)   : read the slot back and produce the STEP model values
template <int D, bool SECOND>
struct YSource {
    static constexpr bool kIsY = true;
    static constexpr bool kAsync = true;    // cp.async ring: 4.4 -> 6.1 TB/s at d = 8 (profiles/r1_notes.md)
    static constexpr int STEP = SmallShape<D, SECOND>::STEP;
    static constexpr int kStageDoubles = STEP * SMALL_THREADS;   // [block b][thread]
    const double* Y;  // element (o, col) at Y[o + nout*col]
    int64_t n;
    int nout, step;
    // r = replicate, s = sample position inside the replicate, o = output
    __device__ __forceinline__ void issue(double* st, int r, int64_t s, int o) const {
        const double* p = Y + o + (int64_t)nout * ((int64_t)r * step * n + s);
        const int64_t bs = (int64_t)nout * n;
#pragma unroll
        for (int b = 0; b < STEP; ++b) cp_async8(st + b * SMALL_THREADS + threadIdx.x, p + bs * b);
    }
    __device__ __forceinline__ void fetch(const double* st, double (&y)[STEP]) const {
#pragma unroll
        for (int b = 0; b < STEP; ++b) y[b] = st[b * SMALL_THREADS + threadIdx.x];
    }
};

// raw A/B rows of a sample: row stride RS doubles in shared memory, chosen so that a quarter-warp's 16-byte reads (even D)
// or a half-warp's 8-byte reads (odd D) fall in distinct banks
template <int D>
struct RowStage {
    static constexpr int RS = (D % 2) ? D : (((D / 2) % 2) ? D : D + 2);
    static constexpr int kStageDoubles = 2 * RS * SMALL_THREADS;
    __device__ __forceinline__ static void issue(double* st, const double* A, const double* B, int64_t row) {
        double* sa = st + threadIdx.x * RS;
        double* sb = sa + RS * SMALL_THREADS;
        const double* pa = A + row * D;
        const double* pb = B + row * D;
        if constexpr (D % 2 == 0) {
#pragma unroll
            for (int i = 0; i < D; i += 2) { cp_async16(sa + i, pa + i); cp_async16(sb + i, pb + i); }
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) { cp_async8(sa + i, pa + i); cp_async8(sb + i, pb + i); }
        }
    }
    __device__ __forceinline__ static void fetch(const double* st, double (&a)[D], double (&b)[D]) {
        const double* sa = st + threadIdx.x * RS;
        const double* sb = sa + RS * SMALL_THREADS;
        if constexpr (D % 2 == 0) {
#pragma unroll
            for (int i = 0; i < D; i += 2) {
                const double2 va = *reinterpret_cast<const double2*>(sa + i), vb = *reinterpret_cast<const double2*>(sb + i);
                a[i] = va.x; a[i + 1] = va.y; b[i] = vb.x; b[i + 1] = vb.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) { a[i] = sa[i]; b[i] = sb[i]; }
        }
    }
};

template <int D>
__device__ __forceinline__ void load_rows(const double* A, const double* B, int64_t row, double (&a)[D], double (&b)[D]) {
    const double* pa = A + row * D;
    const double* pb = B + row * D;
    if constexpr (D % 2 == 0) {  // rows are 16-byte aligned: vector loads
#pragma unroll
        for (int i = 0; i < D; i += 2) {
            double2 va = ldg_stream2(pa + i), vb = ldg_stream2(pb + i);
            a[i] = va.x; a[i + 1] = va.y; b[i] = vb.x; b[i + 1] = vb.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < D; ++i) { a[i] = ldg_stream(pa + i); b[i] = ldg_stream(pb + i); }
    }
}

template <int D, bool SECOND>
struct IshigamiSource {
    static constexpr bool kIsY = false;
    static constexpr bool kAsync = false;   // compute-heavy: the ring did not pay off here (0.140 -> 0.159 ms on config 3)
    static constexpr int STEP = SmallShape<D, SECOND>::STEP;
    static constexpr int kStageDoubles = RowStage<D>::kStageDoubles;
    const double* A;
    const double* B;
    int64_t buf_col0;  // global column of row 0 of A/B
    double pa, pb;     // the model's a, b
    __device__ __forceinline__ void issue(double* st, int64_t col) const { RowStage<D>::issue(st, A, B, col - buf_col0); }
    __device__ __forceinline__ void fetch(const double* st, double (&y)[STEP]) const {
        double a[D], b[D];
        RowStage<D>::fetch(st, a, b);
        eval(a, b, y);
    }
    __device__ __forceinline__ void load(int64_t col, double (&y)[STEP]) const {
        double a[D], b[D];
        load_rows<D>(A, B, col - buf_col0, a, b);
        eval(a, b, y);
    }
    __device__ __forceinline__ void eval(double (&a)[D], double (&b)[D], double (&y)[STEP]) const {
        // shared terms: f = s1 + pa*s2^2 + pb*x3^4*s1   (test/sobol_method.jl:4-8)
        const double s1a = sin(a[0]), s1b = sin(b[0]);
        double t = sin(a[1]); const double q2a = pa * (t * t);
        t = sin(b[1]); const double q2b = pa * (t * t);
        const double p3a = pb * pow4(a[2]), p3b = pb * pow4(b[2]);
        auto f = [](double s1, double q2, double p3) { return s1 + q2 + p3 * s1; };
        y[0] = f(s1a, q2a, p3a);
        y[1] = f(s1b, q2b, p3b);
        y[2] = f(s1b, q2a, p3a); y[3] = f(s1a, q2b, p3a); y[4] = f(s1a, q2a, p3b);
#pragma unroll
        for (int k = 3; k < D; ++k) y[2 + k] = y[0];  // inputs beyond the third do not enter the model
        if constexpr (SECOND) {
            y[2 + D] = f(s1a, q2b, p3b); y[3 + D] = f(s1b, q2a, p3b); y[4 + D] = f(s1b, q2b, p3a);
#pragma unroll
            for (int k = 3; k < D; ++k) y[2 + D + k] = y[1];
        }
    }
};

template <int D, bool SECOND>
struct SobolGSource {
    static constexpr bool kIsY = false;
    static constexpr bool kAsync = false;
    static constexpr int STEP = SmallShape<D, SECOND>::STEP;
    static constexpr int kStageDoubles = RowStage<D>::kStageDoubles;
    const double* A;
    const double* B;
    const double2* coef;  // (c1_k, c0_k): g_k(x) = |x-0.5|*c1_k + c0_k
    int64_t buf_col0;
    __device__ __forceinline__ void issue(double* st, int64_t col) const { RowStage<D>::issue(st, A, B, col - buf_col0); }
    __device__ __forceinline__ void fetch(const double* st, double (&y)[STEP]) const {
        double a[D], b[D];
        RowStage<D>::fetch(st, a, b);
        eval(a, b, y);
    }
    __device__ __forceinline__ void load(int64_t col, double (&y)[STEP]) const {
        double a[D], b[D];
        load_rows<D>(A, B, col - buf_col0, a, b);
        eval(a, b, y);
    }
    __device__ __forceinline__ void eval(double (&a)[D], double (&b)[D], double (&y)[STEP]) const {
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double2 c = __ldg(coef + i);
            a[i] = fma(fabs(a[i] - 0.5), c.x, c.y);
            b[i] = fma(fabs(b[i] - 0.5), c.x, c.y);
        }
        // prefix products PA[k] = prod_{i<k} gA_i (PA[D] = yA), same for B; suffixes built on the fly
        double PA[D + 1], PB[D + 1];
        PA[0] = PB[0] = 1.0;
#pragma unroll
        for (int i = 0; i < D; ++i) { PA[i + 1] = PA[i] * a[i]; PB[i + 1] = PB[i] * b[i]; }
        y[0] = PA[D];
        y[1] = PB[D];
        double sa = 1.0, sb = 1.0;
#pragma unroll
        for (int k = D - 1; k >= 0; --k) {
            y[2 + k] = PA[k] * (b[k] * sa);                          // AB_k: A with coordinate k from B (:96-99)
            if constexpr (SECOND) y[2 + D + k] = PB[k] * (a[k] * sb);   // BA_k (:100-104)
            sa *= a[k];
            sb *= b[k];
        }
    }
};

// ---- the kernel ---------------------------------------------------------------------------------------------------
template <class Src>
struct SmallArgs {
    Src src;
    double* tile_rec;   // [ntiles][nout][nstat]
    TileGeom g;         // fused sources: global columns; Y source: col_lo = 0, col_hi = nboot*n, rep_first = 0
    int nout, nstat, est, t_begin, t_end;
};

// dynamic shared memory: the staging ring, re-used by the tile flush ([SMALL_CH][SMALL_THREADS+1] doubles), + 8 dd slots per warp
template <class Src>
constexpr size_t small_smem_bytes() {
    size_t ring = Src::kAsync ? (size_t)SMALL_NST * Src::kStageDoubles : 0, flush = (size_t)SMALL_CH * (SMALL_THREADS + 1);
    return sizeof(double) * ((ring > flush ? ring : flush) + (SMALL_THREADS / 32) * 8);
}

template <class Src, int D, bool SECOND, bool JANSEN>
__global__ void __launch_bounds__(SMALL_THREADS) small_kernel(SmallArgs<Src> a) {
    using Sh = SmallShape<D, SECOND>;
    constexpr int NE = JANSEN ? 1 : 3;
    constexpr int NACC = D + NE * D + Sh::NPAIR;
    constexpr size_t RING = Src::kAsync ? (size_t)SMALL_NST * Src::kStageDoubles : 0, FLUSH = (size_t)SMALL_CH * (SMALL_THREADS + 1);
    extern __shared__ __align__(16) double small_sm[];
    double* ring = small_sm;
    double (*sm)[SMALL_THREADS + 1] = reinterpret_cast<double (*)[SMALL_THREADS + 1]>(small_sm);
    double (*sdd)[8] = reinterpret_cast<double (*)[8]>(small_sm + (RING > FLUSH ? RING : FLUSH));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int est = a.est;
    const int ntile_units = (a.t_end - a.t_begin) * a.nout;

    for (int u = blockIdx.x; u < ntile_units; u += gridDim.x) {
        const int tile = a.t_begin + u / a.nout, o = u % a.nout;
        int64_t c0, c1;
        tile_range(a.g, tile, c0, c1);
        const int r = a.g.rep_first + tile / a.g.tpr;
        double acc[NACC];
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = 0.0;
        ShiftSums accA, accB;   // running sums of yA, yA^2 and yB, yB^2 (shifted + compensated; un-shifted exactly at the flush)

        auto issue = [&](int64_t col, int stage) {
            if (col < c1) {
                double* st = ring + (size_t)stage * Src::kStageDoubles;
                if constexpr (Src::kIsY) a.src.issue(st, r, col - (int64_t)r * a.g.n, o);
                else a.src.issue(st, col);
            }
            cp_async_commit();   // always commit (possibly empty) so that the group count stays uniform
        };
        // prologue: SMALL_NST-1 samples in flight
        if constexpr (Src::kAsync) {
#pragma unroll
            for (int p = 0; p < SMALL_NST - 1; ++p) issue(c0 + tid + (int64_t)p * SMALL_THREADS, p);
        }
        int stage = 0;
        for (int64_t col = c0 + tid; col < c1; col += SMALL_THREADS) {
            double y[Sh::STEP];
            if constexpr (Src::kAsync) {
                int pstage = stage + SMALL_NST - 1;
                if (pstage >= SMALL_NST) pstage -= SMALL_NST;
                issue(col + (int64_t)(SMALL_NST - 1) * SMALL_THREADS, pstage);
                cp_async_wait<SMALL_NST - 1>();   // this thread's copy for `col` has landed
                a.src.fetch(ring + (size_t)stage * Src::kStageDoubles, y);
                if (++stage == SMALL_NST) stage = 0;
            } else {
                a.src.load(col, y);
            }
            const double fa = y[0], fb = y[1];
            accA.add(fa);
            accB.add(fb);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const double fab = y[2 + k];
                acc[k] = fma(fb, fab - fa, acc[k]);                                        // :192
                if constexpr (JANSEN) {
                    const double df = fa - fab;
                    acc[D + k] = fma(df, df, acc[D + k]);                                  // :213
                } else {
                    if (est == GSA_EI_SOBOL2007) acc[D + k] = fma(fa, fa - fab, acc[D + k]);   // :211
                    else {
                        acc[D + k] = fma(fa, fab, acc[D + k]);                             // :207 / :224
                        acc[2 * D + k] += fa * fa + fab * fab;                             // :219
                        acc[3 * D + k] += fa + fab;                                        // :220
                    }
                }
            }
            if constexpr (SECOND) {
                const double fafb = fa * fb;
                int p = D + NE * D;
#pragma unroll
                for (int k = 0; k < D; ++k)
#pragma unroll
                    for (int j = k + 1; j < D; ++j, ++p) acc[p] += y[2 + D + k] * y[2 + j] - fafb;   // :197
            }
        }
        if constexpr (Src::kAsync) cp_async_wait<0>();

        // ---- tile flush ----
        double* rec = a.tile_rec + ((size_t)tile * a.nout + o) * a.nstat;
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
#pragma unroll
        for (int base = 0; base < NACC; base += SMALL_CH) {
            __syncthreads();
#pragma unroll
            for (int c = 0; c < SMALL_CH; ++c)
                if (base + c < NACC) sm[c][tid] = acc[base + c];
            __syncthreads();
            // thread (c = tid/16, part = tid%16) sums 16 consecutive columns, then a 16-lane butterfly
            const int c = tid >> 4, part = tid & 15;
            double s = 0.0;
            if (base + c < NACC) {
#pragma unroll
                for (int j = 0; j < 16; ++j) s += sm[c][part * 16 + j];
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (part == 0 && base + c < NACC) {
                // accumulator index -> record index.  Accumulators: [V (D) | E terms (NE*D) | pairs]; the record has
                // est_nterms(est) E terms (3 only for Janon2014), so the pair block moves when NE != est_nterms.
                const int i = base + c;
                if (i < 2 * D) rec[STAT_DD + i] = s;
                else if (i < D + NE * D) { if (est == GSA_EI_JANON2014) rec[STAT_DD + i] = s; }
                else rec[stat_pair_base(D, est) + (i - (D + NE * D))] = s;
            }
        }
        __syncthreads();
        if (tid < 4) {
            dd s{0.0, 0.0};
            for (int w = 0; w < SMALL_THREADS / 32; ++w) dd_add_dd(s, dd{sdd[w][2 * tid], sdd[w][2 * tid + 1]});
            rec[2 * tid] = s.hi;
            rec[2 * tid + 1] = s.lo;
        }
        __syncthreads();
        // (no fused tail here: these kernels live on their register budget -- 80 registers, three CTAs per SM -- and are
        //  launch-latency-sized anyway; stage 2 and the exchange follow as launches of their own)
    }
}

}  // namespace gsa
