// gsa_api.cu -- the C ABI of libgsa_cuda.so (include/gsa_cuda.h): handle, model registry, launch logic.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdint>
#include <algorithm>
#include <thread>

#include "gsa_internal.h"
#include "gsa_common.cuh"
#include "models.cuh"
#include "sobol_design.cuh"
#include "estimator_y.cuh"
#include "fused_generic.cuh"
#include "reduce_tiles.cuh"
#include "dispatch.cuh"
#include "peer_state.h"
#include <emmintrin.h>

namespace gsa {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return GSA_ERR_CUDA;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return GSA_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        want = bytes;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        return fail(GSA_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = want;
    return GSA_OK;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// ------------------------------------------------------------------------------------------------
// Joe-Kuo direction numbers (host): Bratley-Fox recurrence on the embedded table.
// ------------------------------------------------------------------------------------------------
extern "C" const unsigned char gsa_joe_kuo_blob[];
extern "C" const unsigned char gsa_joe_kuo_blob_end[];

const uint32_t* joe_kuo_table() { return reinterpret_cast<const uint32_t*>(gsa_joe_kuo_blob); }

void direction_numbers(int dim_first, int ndim, uint32_t* V) {
    const uint32_t* tab = joe_kuo_table();
    for (int jj = 0; jj < ndim; ++jj) {
        const int j = dim_first + jj;
        uint64_t m[32];
        if (j == 0) {
            for (int i = 0; i < 32; ++i) m[i] = 1;
        } else {
            const uint32_t poly = tab[19 * j];
            int deg = 0;
            while ((poly >> (deg + 1)) != 0) ++deg;
            for (int i = 0; i < deg; ++i) m[i] = tab[19 * j + 1 + i];
            for (int i = deg; i < 32; ++i) {
                uint64_t v = m[i - deg] ^ (m[i - deg] << deg);
                for (int k = 1; k < deg; ++k)
                    if ((poly >> (deg - k)) & 1u) v ^= m[i - k] << k;
                m[i] = v;
            }
        }
        for (int i = 0; i < 32; ++i) V[32 * jj + i] = (uint32_t)((m[i] << (31 - i)) & 0xFFFFFFFFull);
    }
}

static bool is_device_or_pinned(const void* p, bool* is_device) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        *is_device = false;
        return false;
    }
    *is_device = at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
    return at.type != cudaMemoryTypeUnregistered;
}

// Parallel host memcpy for staging pageable designs into pinned buffers: one core copies at ~10 GB/s, PCIe Gen5 takes 55 GB/s.
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    unsigned hw = std::thread::hardware_concurrency();
    // (measured on the 16-core host of the B200 box: 1 thread 13.5 GB/s end to end, 8 threads 36.7, 16 threads 42.3; pinned 55.5)
    int nt = (int)std::min<size_t>(std::max(1u, std::min(16u, hw > 2 ? hw - 2 : 1u)), bytes / (4u << 20) + 1);
    if (const char* e = getenv("GSA_STAGING_THREADS")) nt = std::max(1, atoi(e));
    if (nt <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (bytes / nt + 4095) & ~(size_t)4095;
    for (int i = 0; i < nt; ++i) {
        const size_t off = (size_t)i * per;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        th.emplace_back([=]() { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
    }
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------------------
// option / model validation
// ------------------------------------------------------------------------------------------------
static int check_opts(const gsa_sobol_opts* o) {
    if (!o) return fail(GSA_ERR_INVALID, "opts is NULL");
    if (o->nboot < 1) return fail(GSA_ERR_INVALID, "nboot must be >= 1 (got %d)", o->nboot);
    if (o->ei_estimator < 0 || o->ei_estimator > 3)
        return fail(GSA_ERR_INVALID, "unknown Ei_estimator %d (0 Homma1996, 1 Sobol2007, 2 Jansen1999, 3 Janon2014)", o->ei_estimator);
    if (o->input_location != GSA_LOC_HOST && o->input_location != GSA_LOC_DEVICE)
        return fail(GSA_ERR_INVALID, "input_location must be GSA_LOC_HOST or GSA_LOC_DEVICE");
    if (o->nboot > 1 && !(o->conf_level > 0.0 && o->conf_level < 1.0))
        return fail(GSA_ERR_INVALID, "conf_level must be in (0,1) (got %g)", o->conf_level);
    return GSA_OK;
}

static int model_nout(gsa_handle* h, int model, int np, int d, int* nout) {
    if (model >= GSA_MODEL_USER_BASE) {
        const int idx = model - GSA_MODEL_USER_BASE;
        if (!h || idx >= (int)h->user_models.size()) return fail(GSA_ERR_INVALID, "unknown user model id %d for this handle", model);
        *nout = h->user_models[idx].nout;
        return GSA_OK;
    }
    switch (model) {
        case GSA_MODEL_ISHIGAMI:
            if (d < 3) return fail(GSA_ERR_INVALID, "ishigami needs d >= 3 (got %d)", d);
            if (np != 2) return fail(GSA_ERR_INVALID, "ishigami takes 2 parameters {a, b} (got %d)", np);
            *nout = 1;
            return GSA_OK;
        case GSA_MODEL_SOBOL_G:
            if (np != d) return fail(GSA_ERR_INVALID, "sobol_g takes d parameters a[d] (got %d, d=%d)", np, d);
            *nout = 1;
            return GSA_OK;
        case GSA_MODEL_LINEAR:
            if (np < d || np % d) return fail(GSA_ERR_INVALID, "linear takes nout*d parameters W (got %d, d=%d)", np, d);
            *nout = np / d;
            return GSA_OK;
        case GSA_MODEL_ISHI_LINEAR:
            if (d < 3) return fail(GSA_ERR_INVALID, "ishi_linear needs d >= 3 (got %d)", d);
            *nout = 2;
            return GSA_OK;
        default: return fail(GSA_ERR_INVALID, "unknown model id %d", model);
    }
}

// Tiles: aim at ~2 tiles per resident CTA; a tile never straddles a replicate.
static void choose_tiles(int64_t n, int nrep, int resident, int tiles_per_cta, int ts_min, int ts_max, int gran, int* ts, int* tpr) {
    int64_t want_tpr = std::max<int64_t>(1, ((int64_t)tiles_per_cta * resident + nrep - 1) / std::max(nrep, 1));
    int64_t t = (n + want_tpr - 1) / want_tpr;
    t = (t + gran - 1) / gran * gran;
    t = std::max<int64_t>(t, ts_min);
    t = std::min<int64_t>(t, ts_max);
    *ts = (int)t;
    *tpr = (int)std::max<int64_t>(1, (n + t - 1) / t);
}

static int upload_params(gsa_handle* h, const double* params, int np) {
    int rc = h->params.reserve(sizeof(double) * std::max(np, 1));
    if (rc) return rc;
    if (np > 0) {
        // parameters are host memory (include/gsa_cuda.h); identical parameters are not uploaded again
        if ((int)h->params_cache.size() == np && memcmp(h->params_cache.data(), params, sizeof(double) * np) == 0) {
            // the cached upload may have been issued on another stream (gsa_set_stream): order this call after it
            GSA_CUDA(cudaStreamWaitEvent(h->stream, h->ev_params, 0));
            return GSA_OK;
        }
        GSA_CUDA(cudaStreamSynchronize(h->stream));   // params_cache is about to be overwritten: no copy from it may be in flight
        h->params_cache.assign(params, params + np);
        h->aux_kind = -1;
        GSA_CUDA(cudaMemcpyAsync(h->params.p, h->params_cache.data(), sizeof(double) * np, cudaMemcpyHostToDevice, h->stream));
        GSA_CUDA(cudaEventRecord(h->ev_params, h->stream));
    }
    return GSA_OK;
}

static int run_reduce(gsa_handle* h, int recsz, int nstat, int tpr, int rep_first, int nrep, double* partials) {
    if (tpr <= 16) {
        const int64_t n = (int64_t)nrep * recsz;
        reduce_tiles_flat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->tile_rec.as<double>(), partials, recsz, nstat, tpr, rep_first, nrep);
    } else {
        dim3 grid((recsz + RT_X - 1) / RT_X, nrep), block(RT_X, RT_Y);
        reduce_tiles_kernel<<<grid, block, 0, h->stream>>>(h->tile_rec.as<double>(), partials, recsz, nstat, tpr, rep_first);
    }
    h->last_launches++;
    GSA_CUDA(cudaGetLastError());
    return GSA_OK;
}

// ------------------------------------------------------------------------------------------------
// K2..K4 driver
// ------------------------------------------------------------------------------------------------
// design generated inside the fused kernel instead of read from A/B (gsa_sobol_*_generated)
struct GenSpec {
    const double* lb;
    const double* ub;
    int64_t samples;
};

// K1 launch: `nm` matrices (<= 8) of d x ncols starting at sequence index `first`; matrix m takes the dimensions
// [voff[m], voff[m] + d) of the table.  scale/lb: device vectors [d] (ub - lb, lb).
static int launch_design(gsa_handle* h, const double* scale, const double* lb, double* const* out, const int* voff, int nm, int d,
                         uint64_t first, int64_t ncols) {
    if (ncols <= 0) return GSA_OK;
    DesignArgs a;
    a.V = h->vtab.as<uint32_t>();
    a.scale = scale;
    a.lb = lb;
    for (int m = 0; m < 8; ++m) { a.out[m] = m < nm ? out[m] : nullptr; a.voff[m] = m < nm ? voff[m] : 0; }
    a.first = first;
    a.ncols = ncols;
    a.d = d;
    a.dT = d * nm;
    // two dimensions per thread (16-byte stores) when d is even and the outputs are 16-byte aligned
    bool vec2 = d % 2 == 0 && !getenv("GSA_DESIGN_SCALAR");
    for (int m = 0; m < nm; ++m) vec2 = vec2 && (reinterpret_cast<uintptr_t>(a.out[m]) % 16 == 0);
    const int64_t units = vec2 ? a.dT / 2 : a.dT;     // work items per point
    const int tpb = vec2 ? DESIGN_THREADS2 : DESIGN_THREADS;
    // P*units threads ~ 2 full waves of 2048 threads per SM
    const int64_t want = (int64_t)h->sm_count * 2048 * 2;
    int lp = 0;
    while (((int64_t)1 << lp) * units < want && ((int64_t)1 << lp) < ncols) ++lp;
    a.lp = lp;
    const int64_t threads = ((int64_t)1 << lp) * units;
    if (vec2) sobol_design_kernel2<<<(unsigned)((threads + tpb - 1) / tpb), tpb, 0, h->stream>>>(a);
    else sobol_design_kernel<<<(unsigned)((threads + tpb - 1) / tpb), tpb, 0, h->stream>>>(a);
    h->last_launches++;
    GSA_CUDA(cudaGetLastError());
    return GSA_OK;
}

static int ensure_directions(gsa_handle* h, int dT) {
    if (h->vdims >= dT) return GSA_OK;
    h->vhost.resize((size_t)dT * 32);
    direction_numbers(0, dT, h->vhost.data());
    int rc = h->vtab.reserve(sizeof(uint32_t) * 32 * (size_t)dT);
    if (rc) return rc;
    GSA_CUDA(cudaMemcpyAsync(h->vtab.p, h->vhost.data(), sizeof(uint32_t) * 32 * (size_t)dT, cudaMemcpyHostToDevice, h->stream));
    h->vdims = dT;
    return GSA_OK;
}

// What should happen to the partial sums at the end of the call (beyond landing in `partials`): an all-reduce over the handle's
// peer group and / or a copy into the handle's result mailbox in pinned host memory.  Where the plan allows, the kernel's last
// CTA does all of it (tail.cuh: one launch per step); otherwise the classic sequence of launches follows.
struct TailRequest {
    bool exchange = false;   // in: all-reduce over the peer group (gsa_peer_connect)
    bool to_host = false;    // in: final sums -> h->tail_host
    bool fused = false;      // out: done by the kernel's last CTA, h->tail_flag is raised to h->tail_flag_value at the end
};

int peer_allreduce_launch(gsa_handle* h, double* data_dev, size_t count, int nstat, double* host_out_mapped_dev);   // peer.cu

static int ensure_tail(gsa_handle* h, size_t count) {
    if (!h->tail_flag) {
        GSA_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->tail_flag), sizeof(unsigned long long), cudaHostAllocMapped));
        *h->tail_flag = 0;
        h->tail_flag_value = 0;
        GSA_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->tail_flag_dev), h->tail_flag, 0));
    }
    if (h->tail_host_cap < count) {
        GSA_CUDA(cudaStreamSynchronize(h->stream));   // nothing in flight may still write the old mailbox
        if (h->tail_host) cudaFreeHost(h->tail_host);
        h->tail_host = nullptr;
        h->tail_host_cap = 0;
        const size_t cap = count + count / 4 + 64;
        GSA_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->tail_host), sizeof(double) * cap, cudaHostAllocMapped));
        GSA_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->tail_host_dev), h->tail_host, 0));
        h->tail_host_cap = cap;
    }
    return GSA_OK;
}

// Block until the result mailbox of the last call holds the final sums.  Fused tail: spin on the flag the kernel's last CTA
// raises (the host sees the result ~2 us after it is written, without waiting for the stream to drain); otherwise synchronise.
static int wait_tail(gsa_handle* h, const TailRequest& tr) {
    if (tr.fused) {
        volatile unsigned long long* f = h->tail_flag;
        const unsigned long long want = h->tail_flag_value;
        for (unsigned spins = 0; *f < want; ++spins) {
            if ((spins & 0x3fff) == 0x3fff) {   // a failed launch / timed-out exchange never raises the flag: look at the stream now and then
                const cudaError_t q = cudaStreamQuery(h->stream);
                if (q != cudaErrorNotReady) {
                    if (q != cudaSuccess) return cuda_fail(q, "fused kernel");
                    break;
                }
            }
            _mm_pause();
        }
        if (*f < want) {
            GSA_CUDA(cudaStreamSynchronize(h->stream));
            if (*f < want && !(tr.exchange && h->peer && h->peer->status_host && *h->peer->status_host))
                return fail(GSA_ERR_CUDA, "the fused kernel finished without raising its result flag (internal error: tile sign-off incomplete)");
        }
    } else {
        GSA_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (tr.exchange && h->peer && h->peer->status_host && *h->peer->status_host)
        return fail(GSA_ERR_CUDA, "peer all-reduce timed out waiting for a rank");
    return GSA_OK;
}

static int partials_impl(gsa_handle* h, const gsa_sobol_opts* o, int model, const double* params, int np, const double* A,
                         const double* B, int d, int64_t N, int64_t col0, int64_t ncols, double* partials, const GenSpec* gen = nullptr,
                         TailRequest* tr = nullptr) {
    int rc = check_opts(o);
    if (rc) return rc;
    if (d < 1) return fail(GSA_ERR_INVALID, "d must be >= 1");
    if (N < 0 || col0 < 0 || ncols < 0 || col0 + ncols > N) return fail(GSA_ERR_INVALID, "column range [%lld, %lld) outside the d x %lld design", (long long)col0, (long long)(col0 + ncols), (long long)N);
    if (!partials) return fail(GSA_ERR_INVALID, "partials is NULL");
    if (ncols > 0 && !gen && (!A || !B)) return fail(GSA_ERR_INVALID, "A or B is NULL");
    int nout = 0;
    if ((rc = model_nout(h, model, np, d, &nout))) return rc;
    const bool second = o->second_order != 0;
    const int est = o->ei_estimator;
    const int nstat = stat_count(d, est, second), recsz = nstat * nout;
    const int64_t n = N / o->nboot;

    FusedPlan plan;
    // Generated design (gsa(f, ::Sobol, p_range; samples)): product-form models have a kernel that generates its points in registers;
    // every other model / order / estimator takes its regular kernel on column CHUNKS that K1 writes into a scratch buffer just
    // ahead of it (32 MB per matrix: the chunk is consumed out of L2) -- A and B are never materialised as a whole either way.
    bool gen_chunked = false;
    rc = plan_fused(h, model, d, nout, np, second, est, &plan, gen != nullptr);
    if (rc == GSA_ERR_UNSUPPORTED && gen) {
        plan = FusedPlan();
        rc = plan_fused(h, model, d, nout, np, second, est, &plan, false);
        gen_chunked = true;
    }
    if (rc) return rc;
    if (!gen && (plan.kind == PLAN_SMALL_ISHIGAMI || plan.kind == PLAN_SMALL_SOBOLG) && d % 2 == 0 && o->input_location == GSA_LOC_DEVICE &&
        ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15)) {
        // the register kernels read even-d rows with 16-byte vector loads; a device view that starts at an odd element takes
        // the generic kernel instead of faulting (the C ABI accepts any device pointer)
        plan.tiles_per_cta = 2;
        if ((rc = plan_generic(h, &plan))) return rc;
    }
    if (gen) {
        if ((int64_t)2 * o->nboot * d > 21201)
            return fail(GSA_ERR_INVALID, "2*nboot*d = %lld exceeds the 21201 dimensions of the Joe-Kuo table", (long long)2 * o->nboot * d);
        if (gen->samples < 1 || gen->samples >= ((int64_t)1 << 31)) return fail(GSA_ERR_INVALID, "samples must be in [1, 2^31)");
        if ((rc = ensure_directions(h, 2 * o->nboot * d))) return rc;
        if ((rc = h->lbub.reserve(sizeof(double) * 2 * (size_t)d))) return rc;
        std::vector<double> box(2 * (size_t)d);
        plan.gen_unit_box = 1;
        for (int i = 0; i < d; ++i) {
            // (ub - lb, lb): interleaved pairs for the generating kernel, two vectors for K1 (chunked path)
            box[gen_chunked ? i : 2 * i] = gen->ub[i] - gen->lb[i];
            box[gen_chunked ? d + i : 2 * i + 1] = gen->lb[i];
            if (gen->lb[i] != 0.0 || gen->ub[i] != 1.0) plan.gen_unit_box = 0;
        }
        GSA_CUDA(cudaMemcpyAsync(h->lbub.p, box.data(), sizeof(double) * box.size(), cudaMemcpyHostToDevice, h->stream));
        GSA_CUDA(cudaStreamSynchronize(h->stream));   // box is a stack-lifetime staging vector
        uint64_t first = 1;
        while ((first << 1) <= (uint64_t)gen->samples) first <<= 1;   // 2^floor(log2 samples)
        plan.gen_first = first;
        plan.gen_nboot = o->nboot;
    }
    const uint64_t gen_first = plan.gen_first;

    h->last_launches = 0;
    h->kernel_timed = false;
    GSA_CUDA(cudaEventRecord(h->ev0, h->stream));
    const size_t count = (size_t)o->nboot * recsz;
    const bool want_tail = tr && (tr->exchange || tr->to_host);
    if (want_tail && (rc = ensure_tail(h, count))) return rc;
    bool zeroed = false;
    const int64_t valid_hi = std::min<int64_t>(col0 + ncols, (int64_t)o->nboot * n);
    if (n > 0 && valid_hi > col0) {
        if ((rc = upload_params(h, params, np))) return rc;
        if ((rc = plan.prepare(h, params, np))) return rc;
        TileGeom g;
        g.n = n;
        g.col_lo = col0;
        g.col_hi = valid_hi;
        g.buf_col0 = col0;
        g.rep_first = (int)(col0 / n);
        const int rep_last = (int)((valid_hi - 1) / n);
        const int nrep = rep_last - g.rep_first + 1;
        const size_t col_bytes = sizeof(double) * (size_t)d;
        bool devp = false;
        const bool from_host = !gen && o->input_location != GSA_LOC_DEVICE;
        const bool host_in = from_host || gen_chunked;   // the design reaches the kernels chunk by chunk through a scratch buffer
        const bool pinned = from_host && is_device_or_pinned(A, &devp) && is_device_or_pinned(B, &devp);
        // Host designs are streamed in column chunks (256 MB per matrix from pinned memory, 64 MB through the staging
        // buffers otherwise).  The kernel of a chunk only sees that chunk's tiles, so the tiles must be small enough that one
        // chunk holds at least a full wave of them -- with shard-sized tiles each chunk kernel ran on ONE CTA and the whole
        // pipeline sat at 25 GB/s instead of the 55 GB/s PCIe delivers (profiles/r1_notes.md).
        const int64_t chunk_cols_target = std::max<int64_t>(plan.ts_min, (int64_t)(((gen_chunked ? 32u : pinned ? 256u : 64u) << 20) / col_bytes));
        if (host_in) {
            int64_t ts = (chunk_cols_target + plan.resident - 1) / plan.resident;
            ts = (ts + plan.ts_gran - 1) / plan.ts_gran * plan.ts_gran;
            ts = std::max<int64_t>(plan.ts_min, std::min<int64_t>(ts, plan.ts_max));
            g.ts = (int)std::min<int64_t>(ts, std::max<int64_t>(plan.ts_min, (n + plan.ts_gran - 1) / plan.ts_gran * plan.ts_gran));
            g.tpr = (int)std::max<int64_t>(1, (n + g.ts - 1) / g.ts);
        } else {
            // Size the tiles from the columns THIS SHARD holds (a shard can be a thin slice of a huge replicate: sizing them
            // from the replicate length left 3/4 of the CTAs idle on 8 GPUs, profiles/r1_notes.md): about tiles_per_cta * resident
            // tiles over the shard's columns -- and, because tiles are handed out round-robin to `resident` CTAs, a tile count
            // that fills whole waves: 100 replicates x 9 tiles on 444 CTAs is 2.03 waves, i.e. three passes for two passes' worth
            // of work (config-5 shape with nboot = 100: 5.1 ms instead of 4.1).  Among the tile sizes down to a quarter of the
            // first choice (and not below 16 granules: every tile ends with a flush) take the one with the best wave efficiency.
            const int64_t cols_local = valid_hi - col0;
            auto round_ts = [&](int64_t t) {
                t = (t + plan.ts_gran - 1) / plan.ts_gran * plan.ts_gran;
                t = std::max<int64_t>(plan.ts_min, std::min<int64_t>(t, plan.ts_max));
                return std::min<int64_t>(t, std::max<int64_t>(plan.ts_min, (n + plan.ts_gran - 1) / plan.ts_gran * plan.ts_gran));
            };
            auto ntiles_for = [&](int64_t t) { return nrep == 1 ? (cols_local + t - 1) / t : (int64_t)nrep * ((n + t - 1) / t); };
            const int64_t ts0 = round_ts((cols_local + (int64_t)plan.tiles_per_cta * plan.resident - 1) / ((int64_t)plan.tiles_per_cta * plan.resident));
            int64_t best_ts = ts0;
            double best_eff = -1.0;
            const int64_t tpr0 = (n + ts0 - 1) / ts0;
            for (int64_t tpr = tpr0; tpr <= 4 * tpr0 + 1; ++tpr) {
                const int64_t t = round_ts((n + tpr - 1) / tpr);
                if (tpr > tpr0 && (t * 4 < ts0 || t < 16 * (int64_t)plan.ts_gran)) break;   // every tile ends with a flush: keep them long
                const int64_t T = ntiles_for(t), waves = (T + plan.resident - 1) / plan.resident;
                const double eff = (double)T / ((double)waves * plan.resident);
                if (eff > best_eff + 0.01) { best_eff = eff; best_ts = t; }
                if (best_eff >= 0.985) break;
            }
            g.ts = (int)best_ts;
            g.tpr = (int)std::max<int64_t>(1, (n + g.ts - 1) / g.ts);
        }
        g.ntiles = nrep * g.tpr;
        if ((int64_t)nrep * g.tpr > (int64_t)INT32_MAX / 2) return fail(GSA_ERR_UNSUPPORTED, "too many tiles (%lld)", (long long)nrep * g.tpr);
        bool direct = plan.writes_partials();   // the plan's own kernels add per-replicate records to `partials`
        plan.out_partials = partials;
        if (!direct && !host_in && g.tpr == 1) {
            // one tile per replicate and one launch (many short replicates: BASELINE config 2): the tile record IS the replicate's
            // partial sum, so the kernel writes it straight into `partials` (zeroed above) -- no tile records, no stage-2 launch
            plan.rec_base = partials + (size_t)g.rep_first * recsz;
            direct = true;
        }
        if (!direct && (rc = h->tile_rec.reserve(sizeof(double) * (size_t)g.ntiles * recsz))) return rc;
        // tiles that intersect the shard form one contiguous index range [t_lo, t_hi); the others are all-zero records
        const int t_lo = (int)((col0 - (int64_t)g.rep_first * n) / g.ts);
        const int t_hi = (rep_last - g.rep_first) * g.tpr + (int)((valid_hi - 1 - (int64_t)rep_last * n) / g.ts) + 1;
        // one launch per step: stage 2, the exchange and the host copy run in the kernel's last CTA (tail.cuh)
        // (a shard with more than TAIL_G groups = 1024 launched tiles in a replicate, or an exchange of more than 4096 doubles --
        // the fused exchange is the work of ONE CTA -- take the classic launches)
        // Measured (tools/tail_probe.py, profiles/r2_tail_probe.txt): the in-kernel tree costs ~20-30 us for one replicate of
        // 444-888 tiles -- about what reduce_tiles + the copy cost as launches of their own, minus their launch gaps and the
        // host's stream synchronisation -- but 2-3x more than the classic sequence for many replicates of few tiles (every
        // replicate's finisher pays its own fences and host write).  So: few replicates only.
        // (groups with launched tiles per replicate: a shard is a slice of the replicate's tiles)
        const int launched_groups = std::min(tail_groups_per_replicate(g.tpr), (t_hi - t_lo + TAIL_G - 1) / TAIL_G + 1);
        const bool fuse = want_tail && plan.supports_tail() && !host_in && t_hi > t_lo && !getenv("GSA_NO_TAIL") &&
                          launched_groups <= TAIL_G && (!tr->exchange || count <= 4096) &&
                          (nrep <= 4 || getenv("GSA_FORCE_TAIL"));
        if (fuse) {
            TailArgs& t = plan.tail;
            const size_t cbytes = sizeof(unsigned int) * tail_counter_count(nrep, g.tpr);
            if (h->tail_counters.cap < cbytes) h->tail_counters_zeroed = 0;
            if ((rc = h->tail_counters.reserve(cbytes))) return rc;
            if (h->tail_counters_zeroed < cbytes) {   // the counters clean up after themselves: zeroed once per (re)allocation
                GSA_CUDA(cudaMemsetAsync(h->tail_counters.p, 0, h->tail_counters.cap, h->stream));
                h->tail_counters_zeroed = h->tail_counters.cap;
            }
            t.ngr = tail_groups_per_replicate(g.tpr);
            if (!direct && (rc = h->tail_grec.reserve(sizeof(double) * (size_t)nrep * t.ngr * recsz))) return rc;
            t.counters = h->tail_counters.as<unsigned int>();
            t.tile_rec = direct ? nullptr : h->tile_rec.as<double>();
            t.grec = direct ? nullptr : h->tail_grec.as<double>();
            t.partials = partials;
            t.reduce = direct ? 0 : 1;
            t.recsz = recsz; t.nstat = nstat; t.tpr = g.tpr; t.rep_first = g.rep_first; t.nrep = nrep; t.nboot = o->nboot;
            t.t_lo = t_lo; t.t_hi = t_hi;
            t.dbg = getenv("GSA_TAIL_DBG") ? atoi(getenv("GSA_TAIL_DBG")) : 0;
            t.exchange = 0;
            t.host_out = nullptr;
            if (tr->exchange) {
                if ((rc = peer_next_exchange(h, partials, count, nstat, h->tail_host_dev, &t.peer))) return rc;
                t.exchange = 1;
            } else {
                t.host_out = h->tail_host_dev;
            }
            t.host_flag = h->tail_flag_dev;
            t.flag_value = ++h->tail_flag_value;
            tr->fused = true;
        }
        if (!fuse || direct) {   // (the fused stage 2 writes every element of `partials` itself)
            GSA_CUDA(cudaMemsetAsync(partials, 0, sizeof(double) * count, h->stream));
            zeroed = true;
        }
        // tile records outside the launched range are zero records: all of them for the classic stage 2, under the fused tail
        // only those that share a group with a launched tile (the tail sums whole groups; the others are never read)
        int z_lo = 0, z_hi = g.ntiles;
        if (fuse) {
            const int r0 = t_lo / g.tpr * g.tpr, r1 = (t_hi - 1) / g.tpr * g.tpr;
            z_lo = r0 + (t_lo - r0) / TAIL_G * TAIL_G;
            z_hi = std::min(r1 + g.tpr, r1 + ((t_hi - 1 - r1) / TAIL_G + 1) * TAIL_G);
        }
        if (!direct && t_lo > z_lo)
            GSA_CUDA(cudaMemsetAsync(h->tile_rec.as<double>() + (size_t)z_lo * recsz, 0, sizeof(double) * (size_t)(t_lo - z_lo) * recsz, h->stream));
        if (!direct && t_hi < z_hi)
            GSA_CUDA(cudaMemsetAsync(h->tile_rec.as<double>() + (size_t)t_hi * recsz, 0, sizeof(double) * (size_t)(z_hi - t_hi) * recsz, h->stream));

        if (!host_in) {
            if ((rc = plan.launch(h, A, B, g, t_lo, t_hi))) return rc;
        } else {
            // Stream the host design through two device buffers: the H2D copy of chunk i+1 (copy stream)
            // overlaps the kernel of chunk i.  Chunk boundaries are tile boundaries.
            int tiles_per_chunk = (int)std::max<int64_t>(1, chunk_cols_target / g.ts);
            const size_t buf_bytes = 2 * col_bytes * (size_t)tiles_per_chunk * g.ts;
            for (int b = 0; b < (gen_chunked ? 1 : 2); ++b)
                if ((rc = h->dev_in[b].reserve(buf_bytes))) return rc;
            if (from_host && !pinned && h->pinned_cap < buf_bytes) {
                for (int b = 0; b < 2; ++b) {
                    if (h->pinned[b]) cudaFreeHost(h->pinned[b]);
                    h->pinned[b] = nullptr;
                }
                h->pinned_cap = 0;
                for (int b = 0; b < 2; ++b) GSA_CUDA(cudaMallocHost(&h->pinned[b], buf_bytes));
                h->pinned_cap = buf_bytes;
            }
            int ci = 0;
            for (int t0 = t_lo; t0 < t_hi; t0 += tiles_per_chunk, ++ci) {
                const int t1 = std::min(t_hi, t0 + tiles_per_chunk), b = ci & 1;
                // columns needed by this chunk = hull of its NON-EMPTY tiles (tiles clipped away by the shard bounds are empty)
                int64_t c0 = INT64_MAX, c1 = INT64_MIN;
                for (int t = t0; t < t1; ++t) {
                    int64_t a0, a1;
                    tile_range(g, t, a0, a1);
                    if (a1 > a0) { c0 = std::min(c0, a0); c1 = std::max(c1, a1); }
                }
                if (c1 <= c0) {   // nothing of this chunk lies in the shard: its tile records are all-zero
                    if (!direct) GSA_CUDA(cudaMemsetAsync(h->tile_rec.as<double>() + (size_t)t0 * recsz, 0, sizeof(double) * (size_t)(t1 - t0) * recsz, h->stream));
                    --ci;         // keep the buffer parity in step with the chunks that really use a buffer
                    continue;
                }
                const size_t bytes = col_bytes * (size_t)(c1 - c0);
                if (getenv("GSA_DEBUG")) fprintf(stderr, "[gsa] chunk %d tiles [%d,%d) cols [%lld,%lld) bytes %zu buf %zu pinned %d ts %d tpr %d\n", ci, t0, t1, (long long)c0, (long long)c1, bytes, buf_bytes, (int)pinned, g.ts, g.tpr);
                double* dA = h->dev_in[gen_chunked ? 0 : b].as<double>();
                double* dB = dA + (size_t)d * (size_t)tiles_per_chunk * g.ts;
                if (gen_chunked) {
                    // K1 writes the chunk (same stream: the previous chunk's kernel has finished with the buffer): column c is
                    // sample c mod samples of replicate c / samples, A from dimension block r, B from block nboot + r (:414-415)
                    for (int64_t r = c0 / n; r * n < c1; ++r) {
                        const int64_t s0 = std::max<int64_t>(c0, r * n), s1 = std::min<int64_t>(c1, (r + 1) * n);
                        double* outs[2] = {dA + (size_t)d * (size_t)(s0 - c0), dB + (size_t)d * (size_t)(s0 - c0)};
                        const int voff[2] = {(int)r * d, (o->nboot + (int)r) * d};
                        if ((rc = launch_design(h, h->lbub.as<double>(), h->lbub.as<double>() + d, outs, voff, 2, d,
                                                gen_first + (uint64_t)(s0 - r * n), s1 - s0))) return rc;
                    }
                    TileGeom gc = g;
                    gc.buf_col0 = c0;
                    if ((rc = plan.launch(h, dA, dB, gc, t0, t1))) return rc;
                    continue;
                }
                const double* hA = A + (size_t)d * (size_t)(c0 - col0);
                const double* hB = B + (size_t)d * (size_t)(c0 - col0);
                // buffer b free again?  Also on the first two chunks: the previous (asynchronous) call on this handle may still be
                // reading it.  Never-recorded events are complete, so the first call does not wait.
                GSA_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));
                if (!pinned) GSA_CUDA(cudaEventSynchronize(h->ev_copy[b]));            // staging buffer drained?
                if (pinned) {
                    GSA_CUDA(cudaMemcpyAsync(dA, hA, bytes, cudaMemcpyHostToDevice, h->copy_stream));
                    GSA_CUDA(cudaMemcpyAsync(dB, hB, bytes, cudaMemcpyHostToDevice, h->copy_stream));
                } else {
                    char* st = static_cast<char*>(h->pinned[b]);
                    parallel_memcpy(st, hA, bytes);
                    parallel_memcpy(st + bytes, hB, bytes);
                    GSA_CUDA(cudaMemcpyAsync(dA, st, bytes, cudaMemcpyHostToDevice, h->copy_stream));
                    GSA_CUDA(cudaMemcpyAsync(dB, st + bytes, bytes, cudaMemcpyHostToDevice, h->copy_stream));
                }
                GSA_CUDA(cudaEventRecord(h->ev_copy[b], h->copy_stream));
                GSA_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy[b], 0));
                TileGeom gc = g;
                gc.buf_col0 = c0;
                if ((rc = plan.launch(h, dA, dB, gc, t0, t1))) return rc;
                GSA_CUDA(cudaEventRecord(h->ev_done[b], h->stream));
            }
        }
        if (!direct && !(tr && tr->fused) && (rc = run_reduce(h, recsz, nstat, g.tpr, g.rep_first, nrep, partials))) return rc;
    }
    if (!(tr && tr->fused)) {
        if (!zeroed && !(n > 0 && valid_hi > col0)) GSA_CUDA(cudaMemsetAsync(partials, 0, sizeof(double) * count, h->stream));   // empty shard
        if (want_tail) {
            if (tr->exchange) {
                if ((rc = peer_allreduce_launch(h, partials, count, nstat, h->tail_host_dev))) return rc;
            } else {
                GSA_CUDA(cudaMemcpyAsync(h->tail_host, partials, sizeof(double) * count, cudaMemcpyDeviceToHost, h->stream));
            }
        }
    }
    GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
    return GSA_OK;
}

// the estimator kernels on a DEVICE all_y of `nboot` replicates of n samples each (block order [A | B | AB_1.. | BA_1..] per
// replicate); the per-replicate sums are ADDED to partials[rep_first .. rep_first + nboot)
static int y_kernels(gsa_handle* h, const gsa_sobol_opts* o, const double* Yd, int nout, int d, int64_t n, int nboot, int rep_first,
                     double* partials) {
    int rc;
    const bool second = o->second_order != 0;
    const int est = o->ei_estimator, step = second ? 2 * d + 2 : d + 2;
    const int nstat = stat_count(d, est, second), recsz = nstat * nout;
    if (n <= 0 || nboot <= 0) return GSA_OK;
        const char* force = getenv("GSA_FORCE_GENERIC");
        // register kernel: d <= 8 with second order; first order d <= 24 (Jansen: 6.4-6.5 TB/s; at d = 32 it spills and drops to 4.0 TB/s)
        // or 16 (the three-term estimators need 4d accumulators)
        const bool small = d <= (second ? SMALL_DMAX2 : (est == GSA_EI_JANSEN1999 ? SMALL_DMAX_Y : 16)) && !(force && atoi(force) != 0);
        int tpr = 1;
        const char* em = getenv("GSA_PAIRS_MMA");   // tuning knob: 0 = never, 1 = whenever possible; default: d > 8 (the register kernel covers d <= 8)
        const bool pairs_mma = second && est != GSA_EI_JANON2014 && d <= PMMA_DMAX && !(force && atoi(force) != 0) &&
                               (em ? atoi(em) != 0 : !small);
        if (pairs_mma) {
            // second order, d <= 32: the pair sums run as FP64 mma over the samples (fused_pairs_mma.cuh), Y is read exactly once
            size_t smem = 0;
            const void* kern = pairs_mma_kernel_y(d, &smem, est == GSA_EI_JANSEN1999);
            if (smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int nb = 0;
            GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, PMMA_THREADS, smem));
            const int resident = h->sm_count * std::max(1, nb);
            PairsArgs a{};
            a.Y = Yd; a.n = n; a.nout = nout; a.d = d; a.nstat = nstat; a.est = est;
            a.g.n = n; a.g.col_lo = 0; a.g.col_hi = (int64_t)nboot * n; a.g.buf_col0 = 0; a.g.rep_first = 0;
            choose_tiles(n, nboot, std::max(1, resident / nout), 1, PM_CH * PMMA_NW, 1 << 30, PM_CH * PMMA_NW, &a.g.ts, &a.g.tpr);
            a.g.ntiles = a.g.tpr * nboot;
            tpr = a.g.tpr;
            if ((rc = h->tile_rec.reserve(sizeof(double) * (size_t)a.g.ntiles * recsz))) return rc;
            a.tile_rec = h->tile_rec.as<double>();
            a.t_begin = 0; a.t_end = a.g.ntiles;
            const int64_t units = (int64_t)a.g.ntiles * nout;
            void* kargs[] = {&a};
            GSA_CUDA(cudaEventRecord(h->evk0, h->stream));
            GSA_CUDA(cudaLaunchKernel(kern, dim3((unsigned)std::min<int64_t>(units, resident)), dim3(PMMA_THREADS), kargs, smem, h->stream));
            GSA_CUDA(cudaEventRecord(h->evk1, h->stream));
            h->kernel_timed = true;
            h->last_launches++;
            GSA_CUDA(cudaGetLastError());
        } else if (small) {
            // thread-per-sample register kernel: every Y value is read exactly once, straight into registers
            TileGeom g;
            g.n = n; g.col_lo = 0; g.col_hi = (int64_t)nboot * n; g.buf_col0 = 0; g.rep_first = 0;
            int occ = 1;
            if ((rc = launch_small_y(h, d, second, est, Yd, n, nout, step, g, nstat, nullptr, &occ))) return rc;
            choose_tiles(n, nboot, std::max(1, h->sm_count * occ / nout), 1, SMALL_THREADS, 1 << 30, SMALL_THREADS, &g.ts, &g.tpr);
            g.ntiles = g.tpr * nboot;
            tpr = g.tpr;
            if ((rc = h->tile_rec.reserve(sizeof(double) * (size_t)g.ntiles * recsz))) return rc;
            if ((rc = launch_small_y(h, d, second, est, Yd, n, nout, step, g, nstat, h->tile_rec.as<double>()))) return rc;
        } else {
            EstYArgs a;
            a.Y = Yd;
            a.n = n;
            a.nout = nout; a.d = d; a.nstat = nstat; a.est = est; a.step = step;
            int ts_want = KY_TS_DEFAULT, threads = 256;
            if (const char* e = getenv("GSA_KY_TS")) ts_want = std::max(32, std::min(KY_TS_LIMIT, atoi(e) / 32 * 32));   // tuning knobs
            if (const char* e = getenv("GSA_KY_THREADS")) threads = std::max(32, std::min(1024, atoi(e) / 32 * 32));
            a.ts = (int)std::min<int64_t>(ts_want, (n + 31) / 32 * 32);
            a.tpr = (int)((n + a.ts - 1) / a.ts);
            a.ntiles = a.tpr * nboot;
            tpr = a.tpr;
            if ((rc = h->tile_rec.reserve(sizeof(double) * (size_t)a.ntiles * recsz))) return rc;
            a.tile_rec = h->tile_rec.as<double>();
            const size_t smem = sizeof(double) * 2 * (size_t)a.ts;
            const void* kern = second ? (const void*)estimator_y_kernel<true> : (const void*)estimator_y_kernel<false>;
            if (smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int nb = 0;
            GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
            dim3 grid(std::min(a.ntiles, h->sm_count * std::max(1, nb)), nout);
            GSA_CUDA(cudaEventRecord(h->evk0, h->stream));
            if (second) estimator_y_kernel<true><<<grid, threads, smem, h->stream>>>(a);
            else estimator_y_kernel<false><<<grid, threads, smem, h->stream>>>(a);
            GSA_CUDA(cudaEventRecord(h->evk1, h->stream));
            h->kernel_timed = true;
            h->last_launches++;
            GSA_CUDA(cudaGetLastError());
        }
    return run_reduce(h, recsz, nstat, tpr, rep_first, nboot, partials);
}

static int partials_y_impl(gsa_handle* h, const gsa_sobol_opts* o, const double* Y, int nout, int d, int64_t n, double* partials) {
    int rc = check_opts(o);
    if (rc) return rc;
    if (d < 1 || nout < 1 || n < 0) return fail(GSA_ERR_INVALID, "bad shape: nout=%d d=%d n=%lld", nout, d, (long long)n);
    if (!partials || (!Y && n > 0)) return fail(GSA_ERR_INVALID, "Y or partials is NULL");
    const bool second = o->second_order != 0;
    const int est = o->ei_estimator, step = second ? 2 * d + 2 : d + 2;
    const int nstat = stat_count(d, est, second), recsz = nstat * nout;
    h->last_launches = 0;
    h->kernel_timed = false;
    const size_t rep_bytes = sizeof(double) * (size_t)nout * (size_t)n * step;   // one replicate of all_y
    const size_t ybytes = rep_bytes * o->nboot;
    GSA_CUDA(cudaEventRecord(h->ev0, h->stream));
    GSA_CUDA(cudaMemsetAsync(partials, 0, sizeof(double) * (size_t)o->nboot * recsz, h->stream));
    if (o->input_location != GSA_LOC_HOST || ybytes == 0) {
        if ((rc = y_kernels(h, o, Y, nout, d, n, o->nboot, 0, partials))) return rc;
        GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
        return GSA_OK;
    }
    // A host all_y (the reference's `all_y = f(all_points)`, :143) is STREAMED: chunks of whole replicates, or -- when one
    // replicate exceeds the chunk size (config 4: 189 GB in one replicate) -- of sample ranges of a replicate, each copied as
    // `step` contiguous runs into one of two device buffers so that the copy of chunk i+1 overlaps the kernels of chunk i.  A
    // chunk is itself a small all_y (its own n), and the sums are additive, so the same kernels serve it.
    size_t chunk_bytes = (size_t)512 << 20;
    if (const char* e = getenv("GSA_Y_CHUNK_MB")) chunk_bytes = (size_t)std::max(1, atoi(e)) << 20;   // testing knob
    bool devp = false;
    const bool pinned = is_device_or_pinned(Y, &devp);
    if (!pinned) chunk_bytes = std::min(chunk_bytes, (size_t)128 << 20);
    const bool whole = rep_bytes <= chunk_bytes;
    const int reps_per_chunk = whole ? (int)std::min<size_t>((size_t)o->nboot, chunk_bytes / rep_bytes) : 1;
    const int64_t n_chunk = whole ? n : std::max<int64_t>(1, (int64_t)(chunk_bytes / (sizeof(double) * (size_t)nout * step)));
    const size_t buf_bytes = whole ? rep_bytes * reps_per_chunk : sizeof(double) * (size_t)nout * step * (size_t)n_chunk;
    for (int b = 0; b < 2; ++b)
        if ((rc = h->y_in[b].reserve(buf_bytes))) return rc;
    if (!pinned && h->pinned_cap < buf_bytes) {
        for (int b = 0; b < 2; ++b) {
            if (h->pinned[b]) cudaFreeHost(h->pinned[b]);
            h->pinned[b] = nullptr;
        }
        h->pinned_cap = 0;
        for (int b = 0; b < 2; ++b) GSA_CUDA(cudaMallocHost(&h->pinned[b], buf_bytes));
        h->pinned_cap = buf_bytes;
    }
    int ci = 0;
    auto copy_run = [&](int b, size_t dst_off, const double* src, size_t bytes) -> int {
        char* dst = static_cast<char*>(h->y_in[b].p) + dst_off;
        if (pinned) {
            GSA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->copy_stream));
        } else {
            char* st = static_cast<char*>(h->pinned[b]) + dst_off;
            parallel_memcpy(st, src, bytes);
            GSA_CUDA(cudaMemcpyAsync(dst, st, bytes, cudaMemcpyHostToDevice, h->copy_stream));
        }
        return GSA_OK;
    };
    for (int r0 = 0; r0 < o->nboot; r0 += reps_per_chunk) {
        const int nr = std::min(reps_per_chunk, o->nboot - r0);
        for (int64_t s0 = 0; s0 < n; s0 += n_chunk, ++ci) {
            const int64_t ns = std::min(n_chunk, n - s0);
            const int b = ci & 1;
            GSA_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));   // buffer b free again (also across calls)?
            if (!pinned) GSA_CUDA(cudaEventSynchronize(h->ev_copy[b]));        // staging buffer drained?
            if (whole) {
                if ((rc = copy_run(b, 0, Y + (size_t)r0 * (rep_bytes / sizeof(double)), rep_bytes * nr))) return rc;
            } else {
                for (int blk = 0; blk < step; ++blk)
                    if ((rc = copy_run(b, sizeof(double) * (size_t)nout * (size_t)ns * blk,
                                       Y + (size_t)nout * ((size_t)r0 * step * n + (size_t)blk * n + s0), sizeof(double) * (size_t)nout * ns)))
                        return rc;
            }
            GSA_CUDA(cudaEventRecord(h->ev_copy[b], h->copy_stream));
            GSA_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy[b], 0));
            if ((rc = y_kernels(h, o, h->y_in[b].as<double>(), nout, d, ns, nr, r0, partials))) return rc;
            GSA_CUDA(cudaEventRecord(h->ev_done[b], h->stream));
        }
    }
    GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
    return GSA_OK;
}

// entry points for multi.cu
int partials_entry(gsa_handle* h, const gsa_sobol_opts* o, int model, const double* params, int np, const double* A, const double* B,
                   int d, int64_t N, int64_t col0, int64_t ncols, double* partials) {
    return partials_impl(h, o, model, params, np, A, B, d, N, col0, ncols, partials);
}
int model_nout_entry(gsa_handle* h, int model, int np, int d, int* nout) { return model_nout(h, model, np, d, nout); }
int partials_generated_entry(gsa_handle* h, const gsa_sobol_opts* o, int model, const double* params, int np, const double* lb,
                             const double* ub, int d, int64_t samples, int64_t col0, int64_t ncols, double* partials) {
    GenSpec gs{lb, ub, samples};
    return partials_impl(h, o, model, params, np, nullptr, nullptr, d, (int64_t)o->nboot * samples, col0, ncols, partials, &gs);
}

// all_points writer (fuse_designs, reference :94-108 + :116-134)
__global__ void __launch_bounds__(256) all_points_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ P,
                                                         int d, int64_t n, int step, int64_t total /* d * cols */) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t col = e / d;
        const int i = (int)(e - col * d);
        const int64_t blk = col / n, s = col - blk * n;
        const int r = (int)(blk / step), b = (int)(blk - (int64_t)r * step);
        const int64_t src = (int64_t)d * ((int64_t)r * n + s) + i;
        bool fromA;
        if (b == 0) fromA = true;
        else if (b == 1) fromA = false;
        else if (b < 2 + d) fromA = (i != b - 2);       // AB_k: A with row k from B
        else fromA = (i == b - 2 - d);                  // BA_k: B with row k from A
        P[e] = fromA ? A[src] : B[src];
    }
}

}  // namespace gsa

using namespace gsa;

// ================================================================================================
// extern "C"
// ================================================================================================
extern "C" {

int gsa_version(void) { return GSA_VERSION; }

int gsa_last_error(char* buf, size_t n) {
    if (!buf || n == 0) return GSA_ERR_INVALID;
    snprintf(buf, n, "%s", g_err);
    return GSA_OK;
}

int gsa_create(int device, gsa_handle** out) {
    if (!out) return fail(GSA_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    GSA_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0) GSA_CUDA(cudaGetDevice(&device));
    if (device >= ndev) return fail(GSA_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    GSA_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GSA_ERR_UNSUPPORTED, "libgsa_cuda is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor, prop.name);
    DeviceGuard guard(device);
    gsa_handle* h = new (std::nothrow) gsa_handle();
    if (!h) return fail(GSA_ERR_NOMEM, "out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&h->evk0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->evk1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_params, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_switch, cudaEventDisableTiming);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaEventCreateWithFlags(&h->ev_copy[b], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done[b], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        gsa_destroy(h);
        return cuda_fail(e, "gsa_create");
    }
    h->stream = h->own_stream;
    *out = h;
    return GSA_OK;
}

int gsa_destroy(gsa_handle* h) {
    if (!h) return GSA_OK;
    {
        DeviceGuard guard(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        peer_release_handle(h);
        if (h->tail_host) cudaFreeHost(h->tail_host);
        if (h->tail_flag) cudaFreeHost(h->tail_flag);
        DevBuf* bufs[] = {&h->tile_rec, &h->partials, &h->params, &h->vtab, &h->lbub, &h->stage[0], &h->stage[1], &h->dev_in[0], &h->dev_in[1], &h->ybuf, &h->y_in[0], &h->y_in[1], &h->aux, &h->gram, &h->tail_counters, &h->tail_grec};
        for (DevBuf* b : bufs) b->release();
        for (auto& um : h->user_models) {
            if (um.lib) cudaLibraryUnload(um.lib);
            for (cudaLibrary_t l : um.lto_libs) cudaLibraryUnload(l);
        }
        for (int b = 0; b < 2; ++b) {
            if (h->pinned[b]) cudaFreeHost(h->pinned[b]);
            if (h->ev_copy[b]) cudaEventDestroy(h->ev_copy[b]);
            if (h->ev_done[b]) cudaEventDestroy(h->ev_done[b]);
        }
        if (h->ev0) cudaEventDestroy(h->ev0);
        if (h->ev1) cudaEventDestroy(h->ev1);
        if (h->evk0) cudaEventDestroy(h->evk0);
        if (h->evk1) cudaEventDestroy(h->evk1);
        if (h->ev_params) cudaEventDestroy(h->ev_params);
        if (h->ev_switch) cudaEventDestroy(h->ev_switch);
        if (h->own_stream) cudaStreamDestroy(h->own_stream);
        if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    }
    delete h;
    return GSA_OK;
}

int gsa_set_stream(gsa_handle* h, void* s, int external) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    cudaStream_t next = external ? static_cast<cudaStream_t>(s) : h->own_stream;
    if (next != h->stream) {
        // the handle's scratch (tile records, staging buffers, Gram moments) is shared by everything it runs: whatever was issued
        // on the old stream must finish before work on the new stream may reuse it
        DeviceGuard guard(h->device);
        GSA_CUDA(cudaEventRecord(h->ev_switch, h->stream));
        GSA_CUDA(cudaStreamWaitEvent(next, h->ev_switch, 0));
        h->stream = next;
    }
    return GSA_OK;
}

int gsa_synchronize(gsa_handle* h) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    DeviceGuard guard(h->device);
    GSA_CUDA(cudaStreamSynchronize(h->stream));
    return GSA_OK;
}

int gsa_model_lookup(const char* name, int* id) {
    if (!name || !id) return fail(GSA_ERR_INVALID, "name or id is NULL");
    static const struct { const char* n; int id; } tab[] = {
        {"ishigami", GSA_MODEL_ISHIGAMI}, {"sobol_g", GSA_MODEL_SOBOL_G}, {"linear", GSA_MODEL_LINEAR},
        {"linear_gaussian", GSA_MODEL_LINEAR}, {"ishi_linear", GSA_MODEL_ISHI_LINEAR}};
    for (auto& t : tab)
        if (!strcmp(t.n, name)) {
            *id = t.id;
            return GSA_OK;
        }
    return fail(GSA_ERR_INVALID, "unknown model '%s'", name);
}

int gsa_model_nout(gsa_handle* h, int model_id, int nparams, int d, int* nout) {
    if (!nout) return fail(GSA_ERR_INVALID, "nout is NULL");
    return model_nout(h, model_id, nparams, d, nout);
}

int gsa_model_register_fatbin(gsa_handle* h, const void* image, size_t len, const char* symbol, int nout, int* model_id) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    return register_user_model(h, image, len, symbol, GSA_USER_GENERIC, nout, model_id);
}

int gsa_model_register_separable(gsa_handle* h, const void* image, size_t len, const char* symbol, int kind, int* model_id) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    return register_user_model(h, image, len, symbol, kind, 1, model_id);
}

int gsa_sobol_nstat(const gsa_sobol_opts* o, int d, int* nstat) {
    int rc = check_opts(o);
    if (rc) return rc;
    if (!nstat || d < 1) return fail(GSA_ERR_INVALID, "bad arguments");
    *nstat = stat_count(d, o->ei_estimator, o->second_order != 0);
    return GSA_OK;
}

int gsa_sobol_design(gsa_handle* h, int64_t n, int d, int num_mats, const double* lb, const double* ub, int64_t col0,
                     int64_t ncols, double* const* out, int out_location) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    if (n < 1 || d < 1 || num_mats < 1 || !lb || !ub || !out) return fail(GSA_ERR_INVALID, "bad arguments to gsa_sobol_design");
    if (col0 < 0 || ncols < 0 || col0 + ncols > n) return fail(GSA_ERR_INVALID, "column range outside [0, n)");
    if ((int64_t)d * num_mats > 21201) return fail(GSA_ERR_INVALID, "num_mats*d = %lld exceeds the 21201 dimensions of the Joe-Kuo table", (long long)d * num_mats);
    if (n >= ((int64_t)1 << 31)) return fail(GSA_ERR_UNSUPPORTED, "n must be < 2^31");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    const int dT = d * num_mats;
    int rc;
    if ((rc = ensure_directions(h, dT))) return rc;
    if ((rc = h->lbub.reserve(sizeof(double) * 2 * (size_t)d))) return rc;
    std::vector<double> sc(2 * (size_t)d);
    for (int i = 0; i < d; ++i) {
        sc[i] = ub[i] - lb[i];
        sc[d + i] = lb[i];
    }
    GSA_CUDA(cudaMemcpyAsync(h->lbub.p, sc.data(), sizeof(double) * 2 * d, cudaMemcpyHostToDevice, h->stream));
    GSA_CUDA(cudaStreamSynchronize(h->stream));  // sc is a stack-lifetime staging vector
    h->last_launches = 0;
    uint64_t first = 1;
    while ((first << 1) <= (uint64_t)n) first <<= 1;   // 2^floor(log2 n)
    first += (uint64_t)col0;

    // host output: generate column chunks into device scratch and copy back
    const bool to_host = out_location == GSA_LOC_HOST;
    int64_t chunk = ncols;
    if (to_host) chunk = std::max<int64_t>(1, std::min<int64_t>(ncols, (int64_t)(256u << 20) / (sizeof(double) * dT)));
    if (to_host && ncols > 0)
        if ((rc = h->stage[0].reserve(sizeof(double) * (size_t)dT * chunk))) return rc;
    GSA_CUDA(cudaEventRecord(h->ev0, h->stream));
    for (int64_t cc = 0; cc < ncols; cc += chunk) {
        const int64_t nc = std::min(chunk, ncols - cc);
        for (int m0 = 0; m0 < num_mats; m0 += 8) {
            const int nm = std::min(8, num_mats - m0);
            double* outs[8];
            int voff[8];
            for (int m = 0; m < nm; ++m) {
                outs[m] = to_host ? h->stage[0].as<double>() + (size_t)(m0 + m) * d * nc : out[m0 + m] + (size_t)d * cc;
                voff[m] = (m0 + m) * d;
            }
            if ((rc = launch_design(h, h->lbub.as<double>(), h->lbub.as<double>() + d, outs, voff, nm, d, first + (uint64_t)cc, nc))) return rc;
        }
        if (to_host) {
            for (int m = 0; m < num_mats; ++m)
                GSA_CUDA(cudaMemcpyAsync(out[m] + (size_t)d * cc, h->stage[0].as<double>() + (size_t)m * d * nc, sizeof(double) * (size_t)d * nc,
                                         cudaMemcpyDeviceToHost, h->stream));
            GSA_CUDA(cudaStreamSynchronize(h->stream));
        }
    }
    GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
    return GSA_OK;
}

int gsa_sobol_partials(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* A,
                       const double* B, int d, int64_t N, int64_t col0, int64_t ncols, double* partials_dev) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    return partials_impl(h, opts, model_id, params, nparams, A, B, d, N, col0, ncols, partials_dev);
}

static int fetch_and_finalize(gsa_handle* h, const gsa_sobol_opts* opts, int nout, int d, int64_t n, gsa_sobol_result* out);

int gsa_sobol_partials_generated(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* lb,
                                 const double* ub, int d, int64_t samples, int64_t col0, int64_t ncols, double* partials_dev) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    if (!lb || !ub || !opts) return fail(GSA_ERR_INVALID, "lb, ub or opts is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    GenSpec gs{lb, ub, samples};
    return partials_impl(h, opts, model_id, params, nparams, nullptr, nullptr, d, (int64_t)opts->nboot * samples, col0, ncols, partials_dev, &gs);
}

int gsa_sobol_run_generated(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* lb,
                            const double* ub, int d, int64_t samples, gsa_sobol_result* out) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    if (!out || !lb || !ub) return fail(GSA_ERR_INVALID, "result, lb or ub is NULL");
    int rc = check_opts(opts);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    int nout = 0;
    if (d < 1) return fail(GSA_ERR_INVALID, "d must be >= 1");
    if ((rc = model_nout(h, model_id, nparams, d, &nout))) return rc;
    const int nstat = stat_count(d, opts->ei_estimator, opts->second_order != 0);
    if ((rc = h->partials.reserve(sizeof(double) * (size_t)opts->nboot * nout * nstat))) return rc;
    GenSpec gs{lb, ub, samples};
    const int64_t N = (int64_t)opts->nboot * samples;
    TailRequest tr;
    tr.to_host = true;
    if ((rc = partials_impl(h, opts, model_id, params, nparams, nullptr, nullptr, d, N, 0, N, h->partials.as<double>(), &gs, &tr))) return rc;
    if ((rc = wait_tail(h, tr))) return rc;
    return finalize_host(*opts, h->tail_host, nout, d, samples, out);
}

int gsa_sobol_partials_y(gsa_handle* h, const gsa_sobol_opts* opts, const double* Y, int nout, int d, int64_t n, double* partials_dev) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    return partials_y_impl(h, opts, Y, nout, d, n, partials_dev);
}

int gsa_sobol_combine(const double* parts_host, int nparts, size_t count, int nstat, double* out_host) {
    if (!parts_host || !out_host || nparts < 1 || nstat < STAT_DD || count % (size_t)nstat) return fail(GSA_ERR_INVALID, "bad arguments to gsa_sobol_combine");
    for (size_t i = 0; i < count; ++i) {
        const int e = (int)(i % (size_t)nstat);
        if (e < STAT_DD) {
            if (e & 1) continue;
            dd s{0.0, 0.0};
            for (int g = 0; g < nparts; ++g) dd_add_dd(s, dd{parts_host[(size_t)g * count + i], parts_host[(size_t)g * count + i + 1]});
            out_host[i] = s.hi;
            out_host[i + 1] = s.lo;
        } else {
            double s = 0.0;
            for (int g = 0; g < nparts; ++g) s += parts_host[(size_t)g * count + i];
            out_host[i] = s;
        }
    }
    return GSA_OK;
}

int gsa_sobol_finalize(const gsa_sobol_opts* opts, const double* partials_host, int nout, int d, int64_t n, gsa_sobol_result* out) {
    int rc = check_opts(opts);
    if (rc) return rc;
    if (!partials_host || !out || nout < 1 || d < 1) return fail(GSA_ERR_INVALID, "bad arguments to gsa_sobol_finalize");
    return finalize_host(*opts, partials_host, nout, d, n, out);
}

static int fetch_and_finalize(gsa_handle* h, const gsa_sobol_opts* opts, int nout, int d, int64_t n, gsa_sobol_result* out) {
    const int nstat = stat_count(d, opts->ei_estimator, opts->second_order != 0);
    std::vector<double> host((size_t)opts->nboot * nout * nstat);
    GSA_CUDA(cudaMemcpyAsync(host.data(), h->partials.p, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, h->stream));
    GSA_CUDA(cudaStreamSynchronize(h->stream));
    return finalize_host(*opts, host.data(), nout, d, n, out);
}

int gsa_sobol_run(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* A,
                  const double* B, int d, int64_t N, gsa_sobol_result* out) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    if (!out) return fail(GSA_ERR_INVALID, "result is NULL");
    int rc = check_opts(opts);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    int nout = 0;
    if (d < 1) return fail(GSA_ERR_INVALID, "d must be >= 1");
    if ((rc = model_nout(h, model_id, nparams, d, &nout))) return rc;
    const int nstat = stat_count(d, opts->ei_estimator, opts->second_order != 0);
    if ((rc = h->partials.reserve(sizeof(double) * (size_t)opts->nboot * nout * nstat))) return rc;
    TailRequest tr;
    tr.to_host = true;
    if ((rc = partials_impl(h, opts, model_id, params, nparams, A, B, d, N, 0, N, h->partials.as<double>(), nullptr, &tr))) return rc;
    if ((rc = wait_tail(h, tr))) return rc;
    return finalize_host(*opts, h->tail_host, nout, d, N / opts->nboot, out);
}

// One call per step of the sharded form (one process per GPU): partial sums of this rank's column shard, all-reduce over the
// peer group of the handle, epilogue.  One kernel launch where the plan has a fused tail (tail.cuh).
static int run_sharded(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* A,
                       const double* B, const GenSpec* gen, int d, int64_t N, int64_t col0, int64_t ncols, gsa_sobol_result* out) {
    int rc = check_opts(opts);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    int nout = 0;
    if (d < 1) return fail(GSA_ERR_INVALID, "d must be >= 1");
    if ((rc = model_nout(h, model_id, nparams, d, &nout))) return rc;
    const int nstat = stat_count(d, opts->ei_estimator, opts->second_order != 0);
    if ((rc = h->partials.reserve(sizeof(double) * (size_t)opts->nboot * nout * nstat))) return rc;
    TailRequest tr;
    tr.exchange = peer_connected(h);   // a group of one (or no group): the shard is the whole design
    tr.to_host = true;
    if ((rc = partials_impl(h, opts, model_id, params, nparams, A, B, d, N, col0, ncols, h->partials.as<double>(), gen, &tr))) return rc;
    if ((rc = wait_tail(h, tr))) return rc;
    return finalize_host(*opts, h->tail_host, nout, d, N / opts->nboot, out);
}

int gsa_sobol_run_sharded(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* A,
                          const double* B, int d, int64_t N, int64_t col0, int64_t ncols, gsa_sobol_result* out) {
    if (!h || !out) return fail(GSA_ERR_INVALID, "handle or result is NULL");
    return run_sharded(h, opts, model_id, params, nparams, A, B, nullptr, d, N, col0, ncols, out);
}

int gsa_sobol_run_sharded_generated(gsa_handle* h, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams,
                                    const double* lb, const double* ub, int d, int64_t samples, int64_t col0, int64_t ncols,
                                    gsa_sobol_result* out) {
    if (!h || !out || !lb || !ub || !opts) return fail(GSA_ERR_INVALID, "handle, result, lb, ub or opts is NULL");
    GenSpec gs{lb, ub, samples};
    return run_sharded(h, opts, model_id, params, nparams, nullptr, nullptr, &gs, d, (int64_t)opts->nboot * samples, col0, ncols, out);
}

int gsa_sobol_analyze_y(gsa_handle* h, const gsa_sobol_opts* opts, const double* Y, int nout, int d, int64_t n, gsa_sobol_result* out) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    if (!out) return fail(GSA_ERR_INVALID, "result is NULL");
    int rc = check_opts(opts);
    if (rc) return rc;
    if (d < 1 || nout < 1) return fail(GSA_ERR_INVALID, "bad shape");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    const int nstat = stat_count(d, opts->ei_estimator, opts->second_order != 0);
    if ((rc = h->partials.reserve(sizeof(double) * (size_t)opts->nboot * nout * nstat))) return rc;
    if ((rc = partials_y_impl(h, opts, Y, nout, d, n, h->partials.as<double>()))) return rc;
    return fetch_and_finalize(h, opts, nout, d, n, out);
}

int gsa_sobol_all_points(gsa_handle* h, const gsa_sobol_opts* opts, const double* A, const double* B, int d, int64_t N, double* P,
                         int p_location) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    int rc = check_opts(opts);
    if (rc) return rc;
    if (d < 1 || N < 0 || !A || !B || !P) return fail(GSA_ERR_INVALID, "bad arguments to gsa_sobol_all_points");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    const int64_t n = N / opts->nboot;
    const int step = opts->second_order ? 2 * d + 2 : d + 2;
    const int64_t total = (int64_t)d * n * step * opts->nboot;
    if (total == 0) return GSA_OK;
    h->last_launches = 0;
    GSA_CUDA(cudaEventRecord(h->ev0, h->stream));
    const bool a_host = opts->input_location == GSA_LOC_HOST, p_host = p_location == GSA_LOC_HOST;
    if (!a_host && !p_host) {   // everything on the device: one launch
        const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
        all_points_kernel<<<grid, 256, 0, h->stream>>>(A, B, P, d, n, step, total);
        h->last_launches++;
        GSA_CUDA(cudaGetLastError());
        GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
        return GSA_OK;
    }
    // Host designs and / or a host all_points (which is `step` times the size of A and B: 5.5 TB for config 5) go through the
    // device in chunks of a sample range of one replicate: upload the A/B columns, fuse them into a compact [step][ns] block
    // set, and scatter the `step` blocks to their places in P -- nothing of the size of P is ever held on the device.
    size_t chunk_bytes = (size_t)256 << 20;
    if (const char* e = getenv("GSA_P_CHUNK_MB")) chunk_bytes = (size_t)std::max(1, atoi(e)) << 20;   // testing knob
    const int64_t ns_max = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(chunk_bytes / (sizeof(double) * (size_t)d * step))));
    if (a_host)
        for (int b = 0; b < 2; ++b)
            if ((rc = h->dev_in[b].reserve(sizeof(double) * (size_t)d * ns_max))) return rc;
    if ((rc = h->ybuf.reserve(sizeof(double) * (size_t)d * ns_max * step))) return rc;
    double* scratch = h->ybuf.as<double>();
    for (int r = 0; r < opts->nboot; ++r)
        for (int64_t s0 = 0; s0 < n; s0 += ns_max) {
            const int64_t ns = std::min(ns_max, n - s0);
            const size_t in_off = (size_t)d * (size_t)(r * n + s0), in_bytes = sizeof(double) * (size_t)d * ns;
            const double *dA = A + in_off, *dB = B + in_off;
            if (a_host) {
                GSA_CUDA(cudaMemcpyAsync(h->dev_in[0].p, A + in_off, in_bytes, cudaMemcpyHostToDevice, h->stream));
                GSA_CUDA(cudaMemcpyAsync(h->dev_in[1].p, B + in_off, in_bytes, cudaMemcpyHostToDevice, h->stream));
                dA = h->dev_in[0].as<double>();
                dB = h->dev_in[1].as<double>();
            }
            const int64_t ctotal = (int64_t)d * ns * step;
            const int grid = (int)std::min<int64_t>((ctotal + 255) / 256, (int64_t)h->sm_count * 16);
            all_points_kernel<<<grid, 256, 0, h->stream>>>(dA, dB, scratch, d, ns, step, ctotal);   // the chunk as a 1-replicate design
            h->last_launches++;
            GSA_CUDA(cudaGetLastError());
            for (int blk = 0; blk < step; ++blk)
                GSA_CUDA(cudaMemcpyAsync(P + (size_t)d * (((size_t)r * step + blk) * n + s0), scratch + (size_t)d * ns * blk, in_bytes,
                                         p_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, h->stream));
            if (p_host || a_host) GSA_CUDA(cudaStreamSynchronize(h->stream));   // pageable host memory on either side: keep it simple and ordered
        }
    GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
    return GSA_OK;
}

int gsa_host_register(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail(GSA_ERR_INVALID, "bad arguments to gsa_host_register");
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return GSA_OK;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostRegister");
    return GSA_OK;
}

int gsa_host_unregister(void* ptr) {
    if (!ptr) return GSA_OK;
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return GSA_OK;
    }
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostUnregister");
    return GSA_OK;
}

int gsa_last_timing(gsa_handle* h, float* ms, float* kernel_ms, int* launches) {
    if (!h) return fail(GSA_ERR_INVALID, "handle is NULL");
    DeviceGuard guard(h->device);
    GSA_CUDA(cudaEventSynchronize(h->ev1));
    float t = 0.f, tk = 0.f;
    GSA_CUDA(cudaEventElapsedTime(&t, h->ev0, h->ev1));
    if (h->kernel_timed) GSA_CUDA(cudaEventElapsedTime(&tk, h->evk0, h->evk1));
    if (ms) *ms = t;
    if (kernel_ms) *kernel_ms = tk;
    if (launches) *launches = h->last_launches;
    return GSA_OK;
}

}  // extern "C"
