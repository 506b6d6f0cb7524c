// multi.cu -- single-process, multi-device form of the path (what a plain Julia `ccall` needs: SURVEY 8(b) "Threading").
//
// gsa_multi owns one gsa_handle per device.  A run shards the columns of A/B contiguously over the devices,
// launches the fused kernel of every shard from its own host thread (each streams its part of the host design through its own
// PCIe link), then ONE kernel on the first device pulls the nboot*nout*nstat partial sums of every device over NVLink peer
// memory (cudaDeviceEnablePeerAccess; all devices live in this process, so no IPC is needed), adds them in device order --
// double-double entries error-free -- and writes the result straight into pinned host memory for the host epilogue.  No
// collective library is involved.  Devices without peer access to the first one fall back to a copy of their few KB through
// the host, summed in the same order.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
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
constexpr int MULTI_MAX_DEV = 16;
struct GatherArgs {
    const double* part[MULTI_MAX_DEV];   // partial sums of every device that the first device can read (peer memory)
    double* out;                         // mapped pinned host memory
    size_t count;
    int ndev, nstat;
};
// sum over the devices, in device order; (hi, lo) pairs of the double-double entries are added error-free
__global__ void __launch_bounds__(256) multi_gather_sum_kernel(GatherArgs a) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(i % a.nstat);
        if (e < STAT_DD) {
            if (e & 1) continue;
            dd s{0.0, 0.0};
            for (int g = 0; g < a.ndev; ++g) dd_add_dd(s, dd{a.part[g][i], a.part[g][i + 1]});
            a.out[i] = s.hi;
            a.out[i + 1] = s.lo;
        } else {
            double s = 0.0;
            for (int g = 0; g < a.ndev; ++g) s += a.part[g][i];
            a.out[i] = s;
        }
    }
}
}  // namespace

struct gsa_multi {
    std::vector<gsa_handle*> h;
    std::vector<char> peer_ok;      // device g's memory is readable from the first device
    double* host_sum = nullptr;     // pinned + mapped, host_cap doubles
    size_t host_cap = 0;
    std::mutex mu;
    float last_ms = 0.f;
};

extern "C" {

int gsa_multi_create(const int* devices, int ndev, gsa_multi** out) {
    if (!out || ndev < 0 || ndev > 64) return fail(GSA_ERR_INVALID, "bad arguments to gsa_multi_create");
    *out = nullptr;
    if (ndev == 0) {   // all visible devices
        if (devices) return fail(GSA_ERR_INVALID, "gsa_multi_create: ndev == 0 (all devices) takes devices == NULL");
        GSA_CUDA(cudaGetDeviceCount(&ndev));
        if (ndev < 1) return fail(GSA_ERR_CUDA, "no CUDA device visible");
        ndev = std::min(ndev, MULTI_MAX_DEV);
    }
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
    if (rc == GSA_OK && ndev > MULTI_MAX_DEV) rc = fail(GSA_ERR_UNSUPPORTED, "gsa_multi supports up to %d devices", MULTI_MAX_DEV);
    if (rc == GSA_OK) {
        m->peer_ok.assign(ndev, 0);
        m->peer_ok[0] = 1;
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(devs[0]);
        for (int i = 1; i < ndev; ++i) {
            int can = 0;
            if (devs[i] == devs[0]) can = 1;
            else if (cudaDeviceCanAccessPeer(&can, devs[0], devs[i]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[i], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                can = e == cudaSuccess;
                if (!can) cudaGetLastError();
            }
            m->peer_ok[i] = (char)can;
        }
        cudaSetDevice(prev);
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
    if (m->host_sum) cudaFreeHost(m->host_sum);
    for (gsa_handle* hh : m->h) gsa_destroy(hh);
    delete m;
    return GSA_OK;
}

int gsa_multi_device_count(gsa_multi* m, int* ndev) {
    if (!m || !ndev) return fail(GSA_ERR_INVALID, "NULL argument");
    *ndev = (int)m->h.size();
    return GSA_OK;
}

int gsa_multi_model_register_fatbin(gsa_multi* m, const void* image, size_t len, const char* symbol, int nout, int* model_id) {
    if (!m || !model_id) return fail(GSA_ERR_INVALID, "NULL argument to gsa_multi_model_register_fatbin");
    std::lock_guard<std::mutex> lk(m->mu);
    int first = -1;
    for (gsa_handle* hh : m->h) {
        int id = -1;
        const int rc = gsa_model_register_fatbin(hh, image, len, symbol, nout, &id);
        if (rc != GSA_OK) return rc;
        if (first < 0) first = id;
        else if (id != first)   // ids are per-handle sequence numbers: they agree as long as models are only registered through `m`
            return fail(GSA_ERR_INVALID, "user model got id %d on device %d but %d on the first device: register models through gsa_multi only", id, hh->device, first);
    }
    *model_id = first;
    return GSA_OK;
}

int gsa_multi_model_register_separable(gsa_multi* m, const void* image, size_t len, const char* symbol, int kind, int* model_id) {
    if (!m || !model_id) return fail(GSA_ERR_INVALID, "NULL argument to gsa_multi_model_register_separable");
    std::lock_guard<std::mutex> lk(m->mu);
    int first = -1;
    for (gsa_handle* hh : m->h) {
        int id = -1;
        const int rc = gsa_model_register_separable(hh, image, len, symbol, kind, &id);
        if (rc != GSA_OK) return rc;
        if (first < 0) first = id;
        else if (id != first)
            return fail(GSA_ERR_INVALID, "user model got id %d on device %d but %d on the first device: register models through gsa_multi only", id, hh->device, first);
    }
    *model_id = first;
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
    MDBG("run: enter");
    std::lock_guard<std::mutex> lk(m->mu);
    const int G = (int)m->h.size();
    DeviceGuard guard(m->h[0]->device);   // the caller's current device is restored on every return path
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
    // ---- the single exchange of the path: the first device gathers and adds the partial sums over NVLink peer memory ----
    gsa_handle* h0 = m->h[0];
    cudaSetDevice(h0->device);
    if (m->host_cap < cnt) {
        if (m->host_sum) cudaFreeHost(m->host_sum);
        m->host_sum = nullptr;
        m->host_cap = 0;
        GSA_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&m->host_sum), sizeof(double) * cnt, cudaHostAllocMapped));
        m->host_cap = cnt;
    }
    GatherArgs ga{};
    ga.count = cnt;
    ga.nstat = nstat;
    std::vector<std::vector<double>> via_host;   // shards of devices without peer access to the first one
    std::vector<int> via_host_dev;
    for (int g = 0; g < G; ++g) {
        if (m->peer_ok[g]) {
            ga.part[ga.ndev++] = m->h[g]->partials.as<double>();
            if (g > 0) GSA_CUDA(cudaStreamWaitEvent(h0->stream, m->h[g]->ev1, 0));   // ev1: end of that device's partials call
        } else {
            via_host.emplace_back(cnt);
            via_host_dev.push_back(g);
        }
    }
    void* out_dev = nullptr;
    GSA_CUDA(cudaHostGetDevicePointer(&out_dev, m->host_sum, 0));
    ga.out = static_cast<double*>(out_dev);
    const int grid = (int)std::min<size_t>((cnt + 255) / 256, 64);
    multi_gather_sum_kernel<<<grid, 256, 0, h0->stream>>>(ga);
    GSA_CUDA(cudaGetLastError());
    for (size_t k = 0; k < via_host.size(); ++k) {
        gsa_handle* hh = m->h[via_host_dev[k]];
        cudaSetDevice(hh->device);
        GSA_CUDA(cudaMemcpyAsync(via_host[k].data(), hh->partials.p, sizeof(double) * cnt, cudaMemcpyDeviceToHost, hh->stream));
    }
    cudaSetDevice(h0->device);
    GSA_CUDA(cudaStreamSynchronize(h0->stream));
    MDBG("partials on host");
    for (int g = 1; g < G; ++g) {
        cudaSetDevice(m->h[g]->device);
        GSA_CUDA(cudaStreamSynchronize(m->h[g]->stream));
    }
    std::vector<double> host(m->host_sum, m->host_sum + cnt);
    for (const auto& v : via_host)
        for (size_t i = 0; i < cnt; ++i) {
            const int e = (int)(i % nstat);
            if (e < STAT_DD) {
                if (e & 1) continue;
                dd s{host[i], host[i + 1]};
                dd_add_dd(s, dd{v[i], v[i + 1]});
                host[i] = s.hi;
                host[i + 1] = s.lo;
            } else {
                host[i] += v[i];
            }
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
