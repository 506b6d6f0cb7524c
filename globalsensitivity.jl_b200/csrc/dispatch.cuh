// dispatch.cuh -- picks the fused K2+K3 kernel for (model, d, nout, options) and launches it.
#pragma once
#include <algorithm>

#include "gsa_internal.h"
#include "fused_generic.cuh"

namespace gsa {

enum { PLAN_GENERIC = 0 };

struct FusedPlan {
    int kind = PLAN_GENERIC;
    int model = 0, d = 0, nout = 1, np = 0, est = 0, nstat = 0;
    bool second = false;
    int threads = 256;
    size_t smem = 0;
    int resident = 148, ts_min = 256, ts_max = 1 << 20, ts_gran = 32;
    const void* kernel = nullptr;

    int prepare(gsa_handle*, const double*, int) { return GSA_OK; }
    int launch(gsa_handle* h, const double* A, const double* B, const TileGeom& g, int t0, int t1);
};

using GenericKernel = void (*)(GenericArgs);

template <bool SECOND>
static GenericKernel generic_kernel_for(int model) {
    switch (model) {
        case GSA_MODEL_ISHIGAMI: return fused_generic_kernel<IshigamiModel, SECOND>;
        case GSA_MODEL_SOBOL_G: return fused_generic_kernel<SobolGModel, SECOND>;
        case GSA_MODEL_LINEAR: return fused_generic_kernel<LinearModel, SECOND>;
        case GSA_MODEL_ISHI_LINEAR: return fused_generic_kernel<IshiLinearModel, SECOND>;
        default: return nullptr;
    }
}

static int plan_generic(gsa_handle* h, FusedPlan* p) {
    if (p->d > GENERIC_DMAX)
        return fail(GSA_ERR_UNSUPPORTED, "generic fused kernel supports d <= %d (got %d)", GENERIC_DMAX, p->d);
    if (p->nout > GENERIC_NOMAX)
        return fail(GSA_ERR_UNSUPPORTED, "generic fused kernel supports nout <= %d (got %d)", GENERIC_NOMAX, p->nout);
    if (p->second && p->d * p->nout > GENERIC_YMAX)
        return fail(GSA_ERR_UNSUPPORTED, "generic second-order kernel supports d*nout <= %d (got %d)", GENERIC_YMAX, p->d * p->nout);
    GenericKernel k = p->second ? generic_kernel_for<true>(p->model) : generic_kernel_for<false>(p->model);
    if (!k) return fail(GSA_ERR_INVALID, "unknown model id %d", p->model);
    const size_t rec_bytes = sizeof(double) * (size_t)p->nstat * p->nout;
    int warps = (int)std::min<size_t>(8, (160u << 10) / rec_bytes);
    if (warps < 1) return fail(GSA_ERR_UNSUPPORTED, "statistics record of %zu bytes does not fit in shared memory", rec_bytes);
    p->kind = PLAN_GENERIC;
    p->kernel = (const void*)k;
    p->threads = 32 * warps;
    p->smem = rec_bytes * warps;
    if (p->smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
    int nb = 0;
    GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, p->kernel, p->threads, p->smem));
    p->resident = std::max(1, nb) * h->sm_count;
    p->ts_min = 32 * warps;
    p->ts_gran = 32 * warps;
    return GSA_OK;
}

static int plan_fused(gsa_handle* h, int model, int d, int nout, int np, bool second, int est, FusedPlan* p) {
    p->model = model; p->d = d; p->nout = nout; p->np = np; p->second = second; p->est = est;
    p->nstat = stat_count(d, est, second);
    return plan_generic(h, p);
}

inline int FusedPlan::launch(gsa_handle* h, const double* A, const double* B, const TileGeom& g, int t0, int t1) {
    if (t1 <= t0) return GSA_OK;
    const int grid = std::min(t1 - t0, resident);
    if (kind == PLAN_GENERIC) {
        GenericArgs a;
        a.A = A; a.B = B; a.params = h->params.as<double>(); a.tile_rec = h->tile_rec.as<double>();
        a.g = g; a.d = d; a.np = np; a.nout = nout; a.nstat = nstat; a.est = est; a.t_begin = t0; a.t_end = t1;
        ((GenericKernel)kernel)<<<grid, threads, smem, h->stream>>>(a);
    }
    h->last_launches++;
    GSA_CUDA(cudaGetLastError());
    return GSA_OK;
}

}  // namespace gsa
