// dispatch.cuh -- picks the fused K2+K3 kernel for (model, d, nout, options) and launches it.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "gsa_internal.h"
#include "fused_generic.cuh"
#include "fused_product.cuh"
#include "small_launch.cuh"
#include "fused_linear.cuh"
#include "fused_pairs_mma.cuh"

namespace gsa {

enum { PLAN_GENERIC = 0, PLAN_PRODUCT = 1, PLAN_SMALL_ISHIGAMI = 2, PLAN_SMALL_SOBOLG = 3, PLAN_LINEAR = 4, PLAN_PAIRS_MMA = 5 };

struct FusedPlan {
    int kind = PLAN_GENERIC;
    int model = 0, d = 0, nout = 1, np = 0, est = 0, nstat = 0;
    bool second = false;
    int threads = 256;
    size_t smem = 0;
    int resident = 148, tiles_per_cta = 2, ts_min = 256, ts_max = 1 << 20, ts_gran = 32;
    const void* lin_kernel = nullptr;
    size_t lin_smem = 0;
    const void* kernel = nullptr;

    int prod_T = 0, prod_CPL = 0;
    const void* kernel_v2 = nullptr;   // pair-load variant of the product kernel (even d, 16-byte aligned rows; chosen at launch)
    bool user_sep = false;       // product kernel linked against a user's per-coordinate function (gsa_model_register_separable)
    bool gen = false;            // the design is generated inside the kernel (gsa_sobol_*_generated)
    uint64_t gen_first = 0;
    int gen_nboot = 1, gen_unit_box = 0;
    double pa = 0.0, pb = 0.0;   // Ishigami parameters (small kernel)
    bool pairs_tile = false;     // pairs kernel with thread-per-sample evaluation (even d <= 8)
    const void* kernel_alt = nullptr;
    size_t smem_alt = 0;
    int lin_NB = 0, lin_un = 2;  // 8-coordinate blocks of the stacked [a; b] vector (linear Gram kernel); warp steps per load group
    bool writes_partials() const { return kind == PLAN_LINEAR; }   // the plan adds its records to `out_partials` itself
    double* out_partials = nullptr;
    double* rec_base = nullptr;   // where the tile records go: h->tile_rec, or straight into `partials` when a replicate is one tile
    TailArgs tail{};              // tail.counters != nullptr: stage 2 + exchange + host copy happen inside the launch (tail.cuh)
    // (the linear path has its own moment reduction; the thread-per-sample register kernels cannot spare the registers; the pair
    //  kernels lost 8 % of their main loop to it)
    bool supports_tail() const { return kind == PLAN_PRODUCT || kind == PLAN_GENERIC; }

    int prepare(gsa_handle* h, const double* params, int np);
    int launch(gsa_handle* h, const double* A, const double* B, const TileGeom& g, int t0, int t1);
};

using GenericKernel = void (*)(GenericArgs);

template <bool SECOND, bool LARGE>
static GenericKernel generic_kernel_for(int model) {
    switch (model) {
        case GSA_MODEL_ISHIGAMI: return fused_generic_kernel<IshigamiModel, SECOND, LARGE>;
        case GSA_MODEL_SOBOL_G: return fused_generic_kernel<SobolGModel, SECOND, LARGE>;
        case GSA_MODEL_LINEAR: return fused_generic_kernel<LinearModel, SECOND, LARGE>;
        case GSA_MODEL_ISHI_LINEAR: return fused_generic_kernel<IshiLinearModel, SECOND, LARGE>;
        default: return nullptr;
    }
}

static int plan_generic(gsa_handle* h, FusedPlan* p) {
    if (p->d > GENERIC_DMAX)
        return fail(GSA_ERR_UNSUPPORTED, "generic fused kernel supports d <= %d (got %d)", GENERIC_DMAX, p->d);
    if (p->nout > GENERIC_NOMAX)
        return fail(GSA_ERR_UNSUPPORTED, "generic fused kernel supports nout <= %d (got %d)", GENERIC_NOMAX, p->nout);
    if (p->d * p->nout > GENERIC_ACC_L)
        return fail(GSA_ERR_UNSUPPORTED, "generic fused kernel supports d*nout <= %d (got %d)", GENERIC_ACC_L, p->d * p->nout);
    if (p->second && p->d * p->nout > GENERIC_YMAX)
        return fail(GSA_ERR_UNSUPPORTED, "generic second-order kernel supports d*nout <= %d (got %d)", GENERIC_YMAX, p->d * p->nout);
    const bool large = p->d > 32 || p->d * p->nout > GENERIC_ACC_S;   // which instantiation: capacity of the per-thread arrays
    const void* k = nullptr;
    const bool user = p->model >= GSA_MODEL_USER_BASE;
    if (user) {
        const int idx = p->model - GSA_MODEL_USER_BASE;
        if (idx >= (int)h->user_models.size()) return fail(GSA_ERR_INVALID, "unknown user model id %d for this handle", p->model);
        k = user_generic_kernel(h, p->model, p->second, large);
    } else if (p->second) {
        k = large ? (const void*)generic_kernel_for<true, true>(p->model) : (const void*)generic_kernel_for<true, false>(p->model);
    } else {
        k = large ? (const void*)generic_kernel_for<false, true>(p->model) : (const void*)generic_kernel_for<false, false>(p->model);
    }
    if (!k) return fail(GSA_ERR_INVALID, "unknown model id %d", p->model);
    const size_t rec_bytes = sizeof(double) * (size_t)p->nstat * p->nout, per_warp = rec_bytes + sizeof(double) * GENERIC_TILE_DOUBLES;
    // Four warps per CTA (one per SM sub-partition) whenever they fit: every warp of a CTA then runs at the same speed and the two
    // barriers that end a tile cost nothing.  With five warps (what 48 KB allowed) an SM held 15 warps on 4 sub-partitions, the warps
    // on the crowded one ran 4/3 slower, and 38 % of all stall samples were the others waiting for them at the tile barrier
    // (profiles/r2_user_models_ncu.txt).
    int warps = (int)std::min<size_t>(4, ((user ? 48u : 160u) << 10) / per_warp);   // linked kernels: stay within the default 48 KB
    if (warps < 1) return fail(GSA_ERR_UNSUPPORTED, "statistics record of %zu bytes does not fit in shared memory", rec_bytes);
    if (warps == 3) warps = 2;
    p->kind = PLAN_GENERIC;
    p->kernel = k;
    p->threads = 32 * warps;
    p->smem = per_warp * warps;
    int nb = 2;
    if (!user && p->smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, p->kernel, p->threads, p->smem) != cudaSuccess) {   // (also answers for a linked cudaKernel_t)
        cudaGetLastError();
        nb = 2;
    }
    p->resident = std::max(1, nb) * h->sm_count;
    p->tiles_per_cta = 1;    // every tile ends with two CTA barriers: as few, as long tiles as the wave allows
    p->ts_min = 32 * warps;
    p->ts_gran = 32 * warps;
    return GSA_OK;
}

// ---- product-form (Sobol G) fast path -------------------------------------------------------------
using ProductKernel = void (*)(ProductArgs);
static const int kProdCpl[] = {1, 2, 3, 4, 6, 8, 10, 13};

template <int T, bool GEN, bool JANSEN>
static ProductKernel product_kernel_T(int cpl) {
    switch (cpl) {
        case 8: return fused_product_kernel<T, 8, GEN, JANSEN>;
        case 10: return fused_product_kernel<T, 10, GEN, JANSEN>;
        case 13: return fused_product_kernel<T, 13, GEN, JANSEN>;
        default: return nullptr;
    }
}
template <bool GEN, bool JANSEN>
static ProductKernel product_kernel_sel(int T, int cpl) {
    switch (T) {
        case 1:
            switch (cpl) {
                case 1: return fused_product_kernel<1, 1, GEN, JANSEN>;
                case 2: return fused_product_kernel<1, 2, GEN, JANSEN>;
                case 3: return fused_product_kernel<1, 3, GEN, JANSEN>;
                case 4: return fused_product_kernel<1, 4, GEN, JANSEN>;
                case 6: return fused_product_kernel<1, 6, GEN, JANSEN>;
                default: return product_kernel_T<1, GEN, JANSEN>(cpl);
            }
        case 2: return product_kernel_T<2, GEN, JANSEN>(cpl);
        case 4: return product_kernel_T<4, GEN, JANSEN>(cpl);
        case 8: return product_kernel_T<8, GEN, JANSEN>(cpl);
        case 16: return product_kernel_T<16, GEN, JANSEN>(cpl);
        case 32: return product_kernel_T<32, GEN, JANSEN>(cpl);
        default: return nullptr;
    }
}
static ProductKernel product_kernel_for(int T, int cpl, bool gen = false, bool jansen = true) {
    if (!jansen) return gen ? nullptr : product_kernel_sel<false, false>(T, cpl);
    return gen ? product_kernel_sel<true, true>(T, cpl) : product_kernel_sel<false, true>(T, cpl);
}

static bool product_shape(int d, int* T, int* cpl) {
    int t = 1;
    const char* env = getenv("GSA_PRODUCT_T");   // tuning knob: force the lanes-per-sample
    if (env && atoi(env) > 0) t = atoi(env);
    for (; t <= 32; t <<= 1) {
        const int need = (d + t - 1) / t;
        if (need > 13) continue;
        for (int c : kProdCpl)
            if (c >= need && product_kernel_for(t, c)) {
                *T = t;
                *cpl = c;
                return true;
            }
    }
    return false;
}

static int plan_product(gsa_handle* h, FusedPlan* p) {
    p->kind = PLAN_PRODUCT;
    p->threads = PROD_THREADS;
    const int slots = p->prod_T * p->prod_CPL;
    p->smem = sizeof(double) * (2 * slots + (PROD_THREADS / 32) * (2 * slots + 8) + (p->gen ? 2 * slots + (32 * (2 * slots + 1) + 1) / 2 : 0));
    int nb = 0;
    if (p->user_sep) {
        int rc = user_separable_kernel(h, p->model, p->prod_T, p->prod_CPL, p->est == GSA_EI_JANSEN1999, &p->kernel);
        if (rc) return rc;
        if (p->smem > (48u << 10)) return fail(GSA_ERR_UNSUPPORTED, "separable user kernel: d = %d needs %zu bytes of shared memory (max 48 KB)", p->d, p->smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, p->kernel, p->threads, p->smem) != cudaSuccess) {
            cudaGetLastError();
            nb = 2;
        }
    } else {
        p->kernel = (const void*)product_kernel_for(p->prod_T, p->prod_CPL, p->gen, p->est == GSA_EI_JANSEN1999);
        if (p->smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
        GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, p->kernel, p->threads, p->smem));
        // pair loads: T = 8, 14 slots per lane covers 96 < d <= 112 (config 5: d = 100)   [GSA_PRODUCT_VEC2=0 turns it off]
        const char* ev = getenv("GSA_PRODUCT_VEC2");
        if (!p->gen && p->d % 2 == 0 && p->d > 96 && p->d <= 112 && p->est == GSA_EI_JANSEN1999 && !(ev && atoi(ev) == 0)) {
            p->kernel_v2 = (const void*)fused_product_kernel_v2<8, 14, true>;
            int nb2 = 0;
            const size_t smem2 = sizeof(double) * (2 * 112 + (PROD_THREADS / 32) * (2 * 112 + 8));
            GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, p->kernel_v2, p->threads, smem2));
            if (nb2 < nb) p->kernel_v2 = nullptr;   // (must not cost residency)
        }
    }
    p->resident = std::max(1, nb) * h->sm_count;
    if (const char* e = getenv("GSA_TILES_PER_CTA")) p->tiles_per_cta = std::max(1, atoi(e));   // tuning knob
    p->ts_min = p->ts_gran = PROD_THREADS / p->prod_T;
    return GSA_OK;
}

// ---- second-order pair sums on the FP64 tensor-core path (fused_pairs_mma.cuh) ----------------------
constexpr int PMMA_UN = 2, PMMA_ST = 3;   // (ring depth 3 beat 4 and 6 in both sweeps, profiles/r2_pairs_sweep3.txt)
template <bool J>
static const void* pairs_mma_kernel_product_j(int NB, size_t* smem) {
    switch (NB) {
#define GSA_PM(NBV) case NBV: *smem = pm_smem_bytes(pm_ring_doubles_product<NBV, PMMA_UN, PMMA_ST>(), NBV); return (const void*)pairs_mma_kernel<PM_SRC_PRODUCT, NBV, PMMA_UN, PMMA_ST, J>;
        GSA_PM(1) GSA_PM(2) GSA_PM(3) GSA_PM(4)
#undef GSA_PM
        default: return nullptr;
    }
}
static const void* pairs_mma_kernel_product(int NB, size_t* smem, bool jansen = true) {
    if (const char* e = getenv("GSA_PAIRS_VARIANT")) {   // tuning experiment (NB = 3, Jansen): "UN,ST"
        int un = 0, st = 0;
        if (NB == 3 && jansen && sscanf(e, "%d,%d", &un, &st) == 2) {
#define GSA_PMV(UNV, STV) if (un == UNV && st == STV) { *smem = pm_smem_bytes(pm_ring_doubles_product<3, UNV, STV>(), 3); return (const void*)pairs_mma_kernel<PM_SRC_PRODUCT, 3, UNV, STV, true>; }
            GSA_PMV(4, 2) GSA_PMV(1, 6) GSA_PMV(2, 4) GSA_PMV(1, 8)
#undef GSA_PMV
        }
    }
    return jansen ? pairs_mma_kernel_product_j<true>(NB, smem) : pairs_mma_kernel_product_j<false>(NB, smem);
}
constexpr int PMMA_TILE_ST = 3;
template <bool J>
static const void* pairs_mma_kernel_tile_j(int d, size_t* smem) {
    *smem = pm_smem_bytes(pm_ring_doubles_tile(d, PMMA_TILE_ST), 1);
    switch (d) {
        case 2: return (const void*)pairs_mma_kernel<PM_SRC_PRODUCT_TILE, 1, 2, PMMA_TILE_ST, J>;
        case 4: return (const void*)pairs_mma_kernel<PM_SRC_PRODUCT_TILE, 1, 4, PMMA_TILE_ST, J>;
        case 6: return (const void*)pairs_mma_kernel<PM_SRC_PRODUCT_TILE, 1, 6, PMMA_TILE_ST, J>;
        case 8: return (const void*)pairs_mma_kernel<PM_SRC_PRODUCT_TILE, 1, 8, PMMA_TILE_ST, J>;
        default: return nullptr;
    }
}
static const void* pairs_mma_kernel_tile(int d, size_t* smem, bool jansen = true) {
    return jansen ? pairs_mma_kernel_tile_j<true>(d, smem) : pairs_mma_kernel_tile_j<false>(d, smem);
}
// Y source: (samples per chunk, stages) of the warp tiles.  One 8-coordinate block keeps the round-1 shape (32, 3); wider rows
// take small chunks and deep rings (fused_pairs_mma.cuh).  GSA_PAIRS_YCH / GSA_PAIRS_YST: tuning knobs.
template <bool J>
static const void* pairs_mma_kernel_y_j(int NB, int ch, int st) {
#define GSA_PMY(NBV, CHV, STV) if (NB == NBV && ch == CHV && st == STV) return (const void*)pairs_mma_kernel<PM_SRC_Y, NBV, CHV, STV, J>;
#define GSA_PMY_NB(NBV) GSA_PMY(NBV, 32, 2) GSA_PMY(NBV, 32, 3) GSA_PMY(NBV, 16, 3) GSA_PMY(NBV, 16, 4) GSA_PMY(NBV, 16, 6)
    GSA_PMY_NB(1) GSA_PMY_NB(2) GSA_PMY_NB(3) GSA_PMY_NB(4)
#undef GSA_PMY_NB
#undef GSA_PMY
    return nullptr;
}
static const void* pairs_mma_kernel_y(int d, size_t* smem, bool jansen = true) {
    const int NB = (d + 7) / 8;
    // measured at d = 20 (profiles/r2_pairs_sweep2.txt): (16, 4) 85 %, (32, 2) 81 %, (16, 3) 79 %, (32, 3) 65 %, (16, 6) 59 % of the HBM
    // peak -- the deepest ring that still leaves two CTAs per SM
    int ch = NB == 1 ? 32 : 16, st = NB == 1 ? 3 : 4;
    if (NB > 1 && pm_smem_bytes(pm_ring_doubles_y(d, 4, 16), NB) > (112u << 10)) st = 3;
    if (const char* e = getenv("GSA_PAIRS_YCH")) ch = atoi(e) == 32 ? 32 : 16;
    if (const char* e = getenv("GSA_PAIRS_YST")) st = atoi(e);
    if (ch == 32) st = st <= 2 ? 2 : 3;
    else st = st <= 3 ? 3 : st <= 4 ? 4 : 6;
    *smem = pm_smem_bytes(pm_ring_doubles_y(d, st, ch), NB);
    return jansen ? pairs_mma_kernel_y_j<true>(NB, ch, st) : pairs_mma_kernel_y_j<false>(NB, ch, st);
}

static int plan_fused(gsa_handle* h, int model, int d, int nout, int np, bool second, int est, FusedPlan* p, bool gen = false) {
    p->gen = gen;
    p->model = model; p->d = d; p->nout = nout; p->np = np; p->second = second; p->est = est;
    p->nstat = stat_count(d, est, second);
    const char* force = getenv("GSA_FORCE_GENERIC");   // testing knob: always take the generic kernel
    const bool generic_only = force && atoi(force) != 0;
    const bool one_term = est == GSA_EI_JANSEN1999 || (!gen && (est == GSA_EI_SOBOL2007 || est == GSA_EI_HOMMA1996));
    if (!generic_only && model == GSA_MODEL_SOBOL_G && !second && one_term && product_shape(d, &p->prod_T, &p->prod_CPL))
        return plan_product(h, p);
    if (!generic_only && !gen && model >= GSA_MODEL_USER_BASE && model - GSA_MODEL_USER_BASE < (int)h->user_models.size() &&
        h->user_models[model - GSA_MODEL_USER_BASE].kind != GSA_USER_GENERIC && !second && one_term && product_shape(d, &p->prod_T, &p->prod_CPL)) {
        p->user_sep = true;   // product-form / additive user model: the O(d) kernel linked against its per-coordinate function
        return plan_product(h, p);
    }
    if (gen)
        return fail(GSA_ERR_UNSUPPORTED, "fused design generation is implemented for product-form models (sobol_g), S1/ST, Jansen1999, d <= 416; "
                                         "use gsa_sobol_design + gsa_sobol_run for this request");
    if (!generic_only && model == GSA_MODEL_LINEAR && d <= LIN_DMAX) {   // every estimator and S2 are moment forms of the same Gram matrix
        p->kind = PLAN_LINEAR;
        p->threads = LIN_THREADS;
        const int NB = (2 * d + 7) / 8;
        int un = 2;
        int st = NB <= 5 ? 6 : 4;            // groups in flight per thread (ring stages)
        if (const char* e = getenv("GSA_LIN_ST")) st = atoi(e) == 2 ? 2 : atoi(e) == 3 ? 3 : (atoi(e) == 4 ? 4 : 6);   // tuning knobs
        if (const char* e = getenv("GSA_LIN_UN")) un = atoi(e) == 4 ? 4 : 2;
        if (un == 4 && st > 3) st = 3;
        if (un == 2 && st == 2) st = 3;
        const void* kern = nullptr;
#define GSA_LIN_CASE(NBV)                                                                                                                  \
    if (NB == NBV)                                                                                                                         \
        kern = un == 4 ? (st == 3 ? (const void*)linear_gram_kernel<NBV, 4, 3> : (const void*)linear_gram_kernel<NBV, 4, 2>)               \
                       : (st == 6 ? (const void*)linear_gram_kernel<NBV, 2, 6> : st == 4 ? (const void*)linear_gram_kernel<NBV, 2, 4> : (const void*)linear_gram_kernel<NBV, 2, 3>);
        GSA_LIN_CASE(1) GSA_LIN_CASE(2) GSA_LIN_CASE(3) GSA_LIN_CASE(4) GSA_LIN_CASE(5) GSA_LIN_CASE(6) GSA_LIN_CASE(7) GSA_LIN_CASE(8)
#undef GSA_LIN_CASE
        if (!kern) return fail(GSA_ERR_UNSUPPORTED, "linear kernel: no instantiation for %d blocks", NB);
        p->lin_NB = NB;
        GSA_CUDA(cudaFuncSetAttribute((const void*)linear_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_rec_smem_bytes(LIN_DMAX)));
        p->lin_smem = lin_smem_bytes(NB, un, st);
        if (p->lin_smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->lin_smem));
        int nbl = 0;
        GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbl, kern, LIN_THREADS, p->lin_smem));
        p->lin_kernel = kern;
        p->resident = h->sm_count * std::max(1, nbl);
        p->tiles_per_cta = 1;               // few, large tiles: every tile ends with a cross-warp combine and a moments write-out
        p->ts_min = p->ts_gran = 4 * un * LIN_NW;
        p->lin_un = un;
        return GSA_OK;
    }
    {
        // Second order, product-form model, Jansen: the pair sums run as FP64 mma over the samples (fused_pairs_mma.cuh).
        // 8 < d <= 32: one coordinate block per lane group, 7x faster than the lane-per-coordinate kernel with shuffled pair sums it
        // replaced (d = 20, N = 2^21: 1.30 -> 0.19 ms).  Even d <= 8: thread-per-sample evaluation into a warp tile (with a single
        // 8-coordinate block the per-step butterfly of the lane-per-coordinate source dominates: 0.158 vs 0.133 ms for the register
        // kernel on config 3).  Odd d <= 8 stays on the register kernel.
        const char* em = getenv("GSA_PAIRS_MMA");   // tuning knob: 0 = never, 1 = lane-per-coordinate source, 2 = tile source (even d <= 8)
        const int mode = em ? atoi(em) : (d > SMALL_DMAX2 ? 1 : (d % 2 == 0 ? 2 : 0));
        if (!generic_only && model == GSA_MODEL_SOBOL_G && second && est != GSA_EI_JANON2014 && d <= PMMA_DMAX && mode != 0) {
            p->kind = PLAN_PAIRS_MMA;
            p->threads = PMMA_THREADS;
            p->pairs_tile = mode == 2 && d <= 8 && d % 2 == 0;   // (rows must be 16-byte aligned: checked at launch)
            const bool jn = est == GSA_EI_JANSEN1999;
            p->kernel = p->pairs_tile ? pairs_mma_kernel_tile(d, &p->smem, jn) : pairs_mma_kernel_product((d + 7) / 8, &p->smem, jn);
            if (p->smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
            int nb = 0;
            GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, p->kernel, p->threads, p->smem));
            p->resident = std::max(1, nb) * h->sm_count;
            p->tiles_per_cta = 1;
            p->ts_min = p->ts_gran = p->pairs_tile ? PM_CH * PMMA_NW : 4 * PMMA_UN * PMMA_NW;
            if (p->pairs_tile) {   // designs that are not 16-byte aligned take the lane-per-coordinate source at launch
                p->kernel_alt = pairs_mma_kernel_product(1, &p->smem_alt, jn);
                if (p->smem_alt > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(p->kernel_alt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_alt));
            }
            return GSA_OK;
        }
    }
    const bool small_ishi = model == GSA_MODEL_ISHIGAMI && (d == 3 || d == 4);
    const bool small_g = model == GSA_MODEL_SOBOL_G && d <= (second ? SMALL_DMAX2 : SMALL_DMAX);
    if (!generic_only && (small_ishi || small_g)) {
        // thread-per-sample register kernels: few, large tiles (the per-tile flush costs ~300 instructions per thread)
        p->kind = small_ishi ? PLAN_SMALL_ISHIGAMI : PLAN_SMALL_SOBOLG;
        p->threads = SMALL_THREADS;
        int occ = 1, rc;
        TileGeom g0{};
        if (small_ishi) rc = launch_small_ishigami(h, d, second, est, nullptr, nullptr, 0.0, 0.0, g0, 0, 0, 0, nullptr, &occ);
        else rc = launch_small_sobolg(h, d, second, est, nullptr, nullptr, nullptr, g0, 0, 0, 0, nullptr, &occ);
        if (rc) return rc;
        p->resident = h->sm_count * occ;
        p->tiles_per_cta = 1;               // the per-tile flush costs ~300 instructions per thread: one big tile per CTA
        p->ts_min = p->ts_gran = SMALL_THREADS;
        return GSA_OK;
    }
    return plan_generic(h, p);
}

inline int FusedPlan::launch(gsa_handle* h, const double* A, const double* B, const TileGeom& g, int t0, int t1) {
    if (t1 <= t0) return GSA_OK;
    const int grid = std::min(t1 - t0, resident);
    double* recs = rec_base ? rec_base : h->tile_rec.as<double>();
    GSA_CUDA(cudaEventRecord(h->evk0, h->stream));
    if (kind == PLAN_GENERIC) {
        GenericArgs a;
        a.A = A; a.B = B; a.params = h->params.as<double>(); a.tile_rec = recs;
        a.g = g; a.d = d; a.np = np; a.nout = nout; a.nstat = nstat; a.est = est; a.t_begin = t0; a.t_end = t1;
        a.tail = tail;
        void* kargs[] = {&a};
        GSA_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(threads), kargs, smem, h->stream));   // also takes cudaKernel_t handles
    }
    else if (kind == PLAN_SMALL_ISHIGAMI)
        return launch_small_ishigami(h, d, second, est, A, B, pa, pb, g, t0, t1, nstat, recs, nullptr);
    else if (kind == PLAN_SMALL_SOBOLG)
        return launch_small_sobolg(h, d, second, est, A, B, h->aux.as<double2>(), g, t0, t1, nstat, recs, nullptr);
    else if (kind == PLAN_LINEAR) {
        // moments per tile -> moments per replicate of this launch -> records added to `partials` (no tile records, no stage 2)
        const int gs = lin_gram_size(lin_NB);
        const int rl0 = t0 / g.tpr, nrl = (t1 - 1) / g.tpr - rl0 + 1;
        int rc = h->gram.reserve(sizeof(double) * ((size_t)g.ntiles + g.ntiles / g.tpr) * gs);
        if (rc) return rc;
        double* red = h->gram.as<double>() + (size_t)g.ntiles * gs;
        LinearArgs a;
        a.A = A; a.B = B; a.gram = h->gram.as<double>(); a.g = g; a.d = d; a.t_begin = t0; a.t_end = t1;
        void* kargs[] = {&a};
        GSA_CUDA(cudaLaunchKernel(lin_kernel, dim3(grid), dim3(LIN_THREADS), kargs, lin_smem, h->stream));
        GSA_CUDA(cudaEventRecord(h->evk1, h->stream));
        h->kernel_timed = true;
        linear_gram_reduce_kernel<<<dim3((gs + LGR_X - 1) / LGR_X, nrl), dim3(LGR_X, LGR_Y), 0, h->stream>>>(a.gram, red, gs, g.tpr, t0, t1, rl0);
        LinearRecArgs r;
        r.red = red; r.A = A; r.W = h->params.as<double>(); r.partials = out_partials;
        r.g = g; r.d = d; r.NB = lin_NB; r.nout = nout; r.nstat = nstat; r.rl_first = rl0; r.est = est; r.second = second ? 1 : 0;
        linear_records_kernel<<<dim3(nrl, (nout + LREC_OB - 1) / LREC_OB), LREC_OB * LREC_J, lin_rec_smem_bytes(d), h->stream>>>(r);
        h->last_launches += 3;
        GSA_CUDA(cudaGetLastError());
        return GSA_OK;
    }
    else if (kind == PLAN_PAIRS_MMA) {
        PairsArgs a{};
        a.A = A; a.B = B; a.coef = h->aux.as<double2>(); a.tile_rec = recs; a.nout = 1;
        a.g = g; a.d = d; a.nstat = nstat; a.est = est; a.t_begin = t0; a.t_end = t1;
        void* kargs[] = {&a};
        const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
        if (pairs_tile && !aligned) GSA_CUDA(cudaLaunchKernel(kernel_alt, dim3(grid), dim3(threads), kargs, smem_alt, h->stream));
        else GSA_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(threads), kargs, smem, h->stream));
    }
    else if (kind == PLAN_PRODUCT) {
        ProductArgs a;
        a.A = A; a.B = B; a.coef = h->aux.as<double2>(); a.tile_rec = recs;
        a.params = h->params.as<double>(); a.np = np;
        a.g = g; a.d = d; a.nstat = nstat; a.t_begin = t0; a.t_end = t1; a.est = est;
        a.V = h->vtab.as<uint32_t>(); a.box = h->lbub.as<double2>(); a.first = gen_first; a.nboot = gen_nboot; a.unit_box = gen_unit_box;
        a.tail = tail;
        void* kargs[] = {&a};
        const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
        if (kernel_v2 && aligned)
            GSA_CUDA(cudaLaunchKernel(kernel_v2, dim3(grid), dim3(threads), kargs, sizeof(double) * (2 * 112 + (PROD_THREADS / 32) * (2 * 112 + 8)), h->stream));
        else
            GSA_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(threads), kargs, smem, h->stream));   // (also takes the cudaKernel_t of a linked user kernel)
    }
    GSA_CUDA(cudaEventRecord(h->evk1, h->stream));
    h->kernel_timed = true;
    h->last_launches++;
    GSA_CUDA(cudaGetLastError());
    return GSA_OK;
}

inline int FusedPlan::prepare(gsa_handle* h, const double* params, int np_) {
    if (kind == PLAN_SMALL_ISHIGAMI) {
        pa = params[0];
        pb = params[1];
    }
    if ((kind == PLAN_PRODUCT || kind == PLAN_SMALL_SOBOLG || kind == PLAN_PAIRS_MMA) && !user_sep && h->aux_kind != PLAN_PRODUCT) {
        // (upload_params resets aux_kind whenever the parameters change)
        // g_k(x) = (|4x-2| + a_k)/(1 + a_k) = |x - 0.5| * 4/(1+a_k) + a_k/(1+a_k)
        std::vector<double> c(2 * (size_t)d);
        for (int k = 0; k < d; ++k) {
            c[2 * k] = 4.0 / (1.0 + params[k]);
            c[2 * k + 1] = params[k] / (1.0 + params[k]);
        }
        int rc = h->aux.reserve(sizeof(double) * c.size());
        if (rc) return rc;
        GSA_CUDA(cudaMemcpyAsync(h->aux.p, c.data(), sizeof(double) * c.size(), cudaMemcpyHostToDevice, h->stream));
        GSA_CUDA(cudaStreamSynchronize(h->stream));   // c is a stack-lifetime staging vector
        GSA_CUDA(cudaEventRecord(h->ev_params, h->stream));
        h->aux_kind = PLAN_PRODUCT;
    }
    (void)np_;
    return GSA_OK;
}

}  // namespace gsa
