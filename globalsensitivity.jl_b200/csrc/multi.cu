// multi.cu -- single-process, multi-device form of the path (what a plain Julia `ccall` needs: SURVEY 8(b) "Threading").
//
// gsa_multi owns one gsa_handle per device and one NCCL communicator per device (ncclCommInitAll).  A run shards the
// columns of A/B contiguously over the devices, launches the fused kernel of every shard from its own host
// thread (each streams its part of the host design through its own PCIe link), then issues ONE grouped
// ncclAllReduce(sum, fp64) of the nboot*nout*nstat partial sums -- a few KB over NVLink -- and runs the host epilogue.
// NCCL is loaded with dlopen, so libgsa_cuda.so has no link-time dependency on it.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>

#include "gsa_common.cuh"
#include "gsa_internal.h"

namespace gsa {
int partials_entry(gsa_handle* h, const gsa_sobol_opts* o, int model, const double* params, int np, const double* A, const double* B,
                   int d, int64_t N, int64_t col0, int64_t ncols, double* partials);
int model_nout_entry(gsa_handle* h, int model, int np, int d, int* nout);
int partials_generated_entry(gsa_handle* h, const gsa_sobol_opts* o, int model, const double* params, int np, const double* lb,
                             const double* ub, int d, int64_t samples, int64_t col0, int64_t ncols, double* partials);
}  // namespace gsa

using namespace gsa;

#define MDBG(...) do { if (getenv("GSA_DEBUG")) { fprintf(stderr, "[gsa_multi] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)

namespace {
struct NcclApi {
    void* so = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.so) return GSA_OK;
    for (const char* n : {"libnccl.so.2", "libnccl.so"})
        if ((g_nccl.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!g_nccl.so) return fail(GSA_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))dlsym(g_nccl.so, "ncclCommInitAll");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.so, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.so, "ncclAllReduce");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(g_nccl.so, "ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(g_nccl.so, "ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.so, "ncclGetErrorString");
    if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.GroupStart || !g_nccl.GroupEnd)
        return fail(GSA_ERR_NCCL, "libnccl does not export the expected entry points");
    return GSA_OK;
}
int nccl_fail(int e, const char* what) {
    return fail(GSA_ERR_NCCL, "NCCL error %d (%s) in %s", e, g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?", what);
}
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;   // nccl.h: ncclFloat64 = 8, ncclSum = 0
}  // namespace

struct gsa_multi {
    std::vector<gsa_handle*> h;
    std::vector<void*> comm;
    std::mutex mu;
    float last_ms = 0.f;
};

extern "C" {

int gsa_multi_create(const int* devices, int ndev, gsa_multi** out) {
    if (!out || ndev < 1 || ndev > 64) return fail(GSA_ERR_INVALID, "bad arguments to gsa_multi_create");
    *out = nullptr;
    gsa_multi* m = new (std::nothrow) gsa_multi();
    if (!m) return fail(GSA_ERR_NOMEM, "out of host memory");
    std::vector<int> devs(ndev);
    for (int i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : i;
    int rc = GSA_OK;
    for (int i = 0; i < ndev && rc == GSA_OK; ++i) {
        gsa_handle* hh = nullptr;
        rc = gsa_create(devs[i], &hh);
        if (rc == GSA_OK) m->h.push_back(hh);
    }
    if (rc == GSA_OK && ndev > 1) {
        rc = load_nccl();
        if (rc == GSA_OK) {
            m->comm.assign(ndev, nullptr);
            int e = g_nccl.CommInitAll(m->comm.data(), ndev, devs.data());
            if (e) {
                m->comm.clear();
                rc = nccl_fail(e, "ncclCommInitAll");
            }
        }
    }
    if (rc != GSA_OK) {
        char keep[512];
        gsa_last_error(keep, sizeof(keep));
        gsa_multi_destroy(m);
        set_error("%s", keep);
        return rc;
    }
    *out = m;
    return GSA_OK;
}

int gsa_multi_destroy(gsa_multi* m) {
    if (!m) return GSA_OK;
    for (void* c : m->comm)
        if (c && g_nccl.CommDestroy) g_nccl.CommDestroy(c);
    for (gsa_handle* hh : m->h) gsa_destroy(hh);
    delete m;
    return GSA_OK;
}

int gsa_multi_device_count(gsa_multi* m, int* ndev) {
    if (!m || !ndev) return fail(GSA_ERR_INVALID, "NULL argument");
    *ndev = (int)m->h.size();
    return GSA_OK;
}

int gsa_multi_handle(gsa_multi* m, int i, gsa_handle** out) {
    if (!m || !out || i < 0 || i >= (int)m->h.size()) return fail(GSA_ERR_INVALID, "bad arguments to gsa_multi_handle");
    *out = m->h[i];
    return GSA_OK;
}

}  // extern "C"

// gsa(model, Sobol(...), A, B; batch=true) with the design sharded over the devices of `m`.  A, B: d x N in HOST memory
// (input_location must be GSA_LOC_HOST; pinned memory is streamed without staging); or, with lb/ub, the generated design.
static int multi_run(gsa_multi* m, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* A,
                     const double* B, const double* lb, const double* ub, int64_t samples, int d, int64_t N, gsa_sobol_result* out) {
    const bool generated = lb != nullptr;
    if (model_id >= GSA_MODEL_USER_BASE) return fail(GSA_ERR_UNSUPPORTED, "user models are registered per handle; not supported by the multi-device entry yet");
    MDBG("run: enter");
    std::lock_guard<std::mutex> lk(m->mu);
    const int G = (int)m->h.size();
    int nout = 0, rc;
    if (d < 1) return fail(GSA_ERR_INVALID, "d must be >= 1");
    if ((rc = model_nout_entry(m->h[0], model_id, nparams, d, &nout))) return rc;
    if (opts->nboot < 1) return fail(GSA_ERR_INVALID, "nboot must be >= 1");
    const int nstat = stat_count(d, opts->ei_estimator, opts->second_order != 0);
    const size_t cnt = (size_t)opts->nboot * nout * nstat;
    std::vector<int> rcs(G, GSA_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g]() {
            gsa_handle* hh = m->h[g];
            const int64_t base = N / G, rem = N % G;
            const int64_t col0 = g * base + std::min<int64_t>(g, rem), ncols = base + (g < rem ? 1 : 0);
            MDBG("worker %d: start", g);
            std::lock_guard<std::mutex> hl(hh->mu);
            cudaSetDevice(hh->device);
            MDBG("worker %d: device set", g);
            int r = hh->partials.reserve(sizeof(double) * cnt);
            if (r == GSA_OK && generated)
                r = partials_generated_entry(hh, opts, model_id, params, nparams, lb, ub, d, samples, col0, ncols, hh->partials.as<double>());
            else if (r == GSA_OK)
                r = partials_entry(hh, opts, model_id, params, nparams, A ? A + (size_t)d * col0 : nullptr, B ? B + (size_t)d * col0 : nullptr, d, N,
                                   col0, ncols, hh->partials.as<double>());
            if (r != GSA_OK) {
                char buf[512];
                gsa_last_error(buf, sizeof(buf));   // thread-local: carry it back to the caller's thread
                errs[g] = buf;
            }
            rcs[g] = r;
            MDBG("worker %d: done rc=%d", g, r);
        });
    }
    for (auto& t : th) t.join();
    MDBG("joined");
    for (int g = 0; g < G; ++g)
        if (rcs[g] != GSA_OK) return fail(rcs[g], "device %d: %s", m->h[g]->device, errs[g].c_str());
    if (G > 1) {   // the single collective of the path
        int e = g_nccl.GroupStart();
        if (e) return nccl_fail(e, "ncclGroupStart");
        for (int g = 0; g < G && !e; ++g)
            e = g_nccl.AllReduce(m->h[g]->partials.p, m->h[g]->partials.p, cnt, NCCL_DOUBLE, NCCL_SUM, m->comm[g], m->h[g]->stream);
        int e2 = g_nccl.GroupEnd();
        if (e || e2) return nccl_fail(e ? e : e2, "ncclAllReduce");
    }
    gsa_handle* h0 = m->h[0];
    cudaSetDevice(h0->device);
    if (getenv("GSA_DEBUG")) {
        for (int it = 0; it < 20; ++it) {
            cudaError_t q1 = cudaStreamQuery(h0->stream), q2 = cudaStreamQuery(h0->copy_stream);
            MDBG("query: stream %d copy %d", (int)q1, (int)q2);
            if (q1 == cudaSuccess && q2 == cudaSuccess) break;
            std::this_thread::sleep_for(std::chrono::milliseconds(250));
        }
    }
    std::vector<double> host(cnt);
    GSA_CUDA(cudaMemcpyAsync(host.data(), h0->partials.p, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h0->stream));
    GSA_CUDA(cudaStreamSynchronize(h0->stream));
    MDBG("partials on host");
    for (int g = 1; g < G; ++g) {
        cudaSetDevice(m->h[g]->device);
        GSA_CUDA(cudaStreamSynchronize(m->h[g]->stream));
    }
    return finalize_host(*opts, host.data(), nout, d, N / opts->nboot, out);
}

extern "C" {

int gsa_multi_sobol_run(gsa_multi* m, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams, const double* A,
                        const double* B, int d, int64_t N, gsa_sobol_result* out) {
    if (!m || !opts || !out) return fail(GSA_ERR_INVALID, "NULL argument to gsa_multi_sobol_run");
    if (opts->input_location != GSA_LOC_HOST) return fail(GSA_ERR_INVALID, "gsa_multi_sobol_run takes host designs (input_location = GSA_LOC_HOST)");
    return multi_run(m, opts, model_id, params, nparams, A, B, nullptr, nullptr, 0, d, N, out);
}

// gsa(model, Sobol(...), p_range; samples) (:407-417) over all devices with the design generated inside the fused kernel:
// nothing but lb/ub goes in and nothing but the partial sums comes out.
int gsa_multi_sobol_run_generated(gsa_multi* m, const gsa_sobol_opts* opts, int model_id, const double* params, int nparams,
                                  const double* lb, const double* ub, int d, int64_t samples, gsa_sobol_result* out) {
    if (!m || !opts || !out || !lb || !ub) return fail(GSA_ERR_INVALID, "NULL argument to gsa_multi_sobol_run_generated");
    if (opts->nboot < 1 || samples < 1) return fail(GSA_ERR_INVALID, "nboot and samples must be >= 1");
    return multi_run(m, opts, model_id, params, nparams, nullptr, nullptr, lb, ub, samples, d, (int64_t)opts->nboot * samples, out);
}

}  // extern "C"
