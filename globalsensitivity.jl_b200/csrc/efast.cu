// efast.cu -- C ABI of the eFAST method (include/gsa_cuda.h, "eFAST"): design, fused design + model evaluation, spectral analysis.
// Reference: src/eFAST_sensitivity.jl:75-213.  Kernels in efast.cuh.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "efast.cuh"
#include "efast_fft.cuh"
#include "gsa_internal.h"

namespace gsa {

// omega (:83-96): omega[0] = (samples - 1) / (2 H) for the parameter of interest; the others get floor(range(1, stop = m,
// length = d - 1)) if m = omega[0] / (2 H) >= d - 1 (evaluated in exact integer arithmetic: 1 + k (m - 1) / (d - 2)), else
// (k mod m) + 1.
static int efast_frequencies(int d, int64_t S, int H, std::vector<int64_t>* omega) {
    if (d < 1 || H < 1 || S < 2) return fail(GSA_ERR_INVALID, "eFAST: bad d / num_harmonics / samples");
    const int64_t w0 = (S - 1) / (2 * H), m = w0 / (2 * H);
    if (w0 < 1 || (d > 1 && m < 1)) return fail(GSA_ERR_INVALID, "eFAST: samples = %lld is too small for num_harmonics = %d (needs samples - 1 >= 4 H^2)", (long long)S, H);
    if ((S & 1) && (int64_t)H * w0 > S / 2 - 1)
        return fail(GSA_ERR_INVALID, "eFAST: harmonic %d of omega = %lld lies beyond bin samples/2 - 1 (the reference raises a BoundsError here)", H, (long long)w0);
    omega->assign(d, w0);
    for (int k = 0; k < d - 1; ++k) {
        if (m >= d - 1) (*omega)[k + 1] = d > 2 ? 1 + ((int64_t)k * (m - 1)) / (d - 2) : 1;
        else (*omega)[k + 1] = (k % m) + 1;
    }
    return GSA_OK;
}

struct EfastBuffers {
    double *lb, *ub, *phi;
    int64_t* omega;
};
// lb, ub, phi, omega -> one device allocation (h->lbub)
static int efast_upload(gsa_handle* h, const double* lb, const double* ub, const double* phi, const std::vector<int64_t>& omega, int d, EfastBuffers* b) {
    int rc = h->lbub.reserve(sizeof(double) * 4 * (size_t)d);
    if (rc) return rc;
    std::vector<double> host(4 * (size_t)d);
    for (int i = 0; i < d; ++i) {
        host[i] = lb[i];
        host[d + i] = ub[i];
        host[2 * d + i] = phi ? phi[i] : 0.0;
        int64_t w = omega[i];
        memcpy(&host[3 * d + i], &w, sizeof(double));
    }
    GSA_CUDA(cudaMemcpyAsync(h->lbub.p, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, h->stream));
    GSA_CUDA(cudaStreamSynchronize(h->stream));   // host is a stack-lifetime staging vector
    b->lb = h->lbub.as<double>();
    b->ub = b->lb + d;
    b->phi = b->lb + 2 * d;
    b->omega = reinterpret_cast<int64_t*>(b->lb + 3 * d);
    return GSA_OK;
}

using EvalKernel = void (*)(EfastEvalArgs);
static const void* efast_eval_kernel_for(gsa_handle* h, int model) {
    if (model >= GSA_MODEL_USER_BASE) {
        const int idx = model - GSA_MODEL_USER_BASE;
        if (idx >= (int)h->user_models.size()) return nullptr;
        cudaKernel_t k = nullptr;
        if (cudaLibraryGetKernel(&k, h->user_models[idx].lib, "gsa_user_efast_eval") != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (const void*)k;
    }
    switch (model) {
        case GSA_MODEL_ISHIGAMI: return (const void*)(EvalKernel)efast_eval_kernel<IshigamiModel>;
        case GSA_MODEL_SOBOL_G: return (const void*)(EvalKernel)efast_eval_kernel<SobolGModel>;
        case GSA_MODEL_LINEAR: return (const void*)(EvalKernel)efast_eval_kernel<LinearModel>;
        case GSA_MODEL_ISHI_LINEAR: return (const void*)(EvalKernel)efast_eval_kernel<IshiLinearModel>;
        default: return nullptr;
    }
}

int model_nout_entry(gsa_handle* h, int model, int np, int d, int* nout);   // gsa_api.cu

// spectral analysis of a DEVICE all_y (nout x samples*d) -> S1, ST (host, nout x d column-major)
static int efast_analyze_device(gsa_handle* h, const double* Yd, int nout, int d, int64_t S, int H, gsa_efast_result* out) {
    std::vector<int64_t> omega;
    int rc = efast_frequencies(d, S, H, &omega);
    if (rc) return rc;
    EfastSpecArgs a{};
    a.g.d = d; a.g.nout = nout; a.g.H = H; a.g.S = S; a.g.omega0 = omega[0];
    a.g.nlow = (int)(omega[0] / 2);
    a.g.nbins = a.g.nlow + H + (int)(S & 1);
    const int nser = d * nout;
    // the spectrum: a Stockham FFT when the size factors into 2, 3, 5 (efast_fft.cuh), else the pruned DFT (any size).
    // GSA_EFAST_DFT=1 forces the DFT (tests compare the two).
    const int64_t M = (S & 1) ? S : S / 2;
    std::vector<int> radices;
    const char* force_dft = getenv("GSA_EFAST_DFT");
    const bool fft = efast_fft_plan(M, &radices) && !(force_dft && atoi(force_dft) != 0);
    // scratch: centred series, moments, powers, results, two complex work arrays -- one allocation
    const size_t nC = ((size_t)nser * S + 1) & ~(size_t)1, nM = 4 * (size_t)nser, nP = ((size_t)nser * a.g.nbins + 1) & ~(size_t)1, nR = 2 * (size_t)nser;
    const size_t nW = fft ? 2 * (size_t)nser * M : 0;
    if ((rc = h->ybuf.reserve(sizeof(double) * (nC + nM + nP + nR + 2 * nW)))) return rc;
    a.Y = Yd;
    a.C = h->ybuf.as<double>();
    a.mom = a.C + nC;
    a.power = a.mom + nM;
    a.S1 = a.power + nP;
    a.ST = a.S1 + nser;
    double* work[2] = {a.ST + nser, a.ST + nser + nW};
    efast_center_kernel<<<nser, 256, 0, h->stream>>>(a);
    GSA_CUDA(cudaEventRecord(h->evk0, h->stream));
    if (fft) {
        EfastFftArgs f{a.C, work[0], M, 1};
        bool first = true;
        int w = 0;
        for (int R : radices) {
            const dim3 grid((unsigned)((M / R + 255) / 256), nser);
            const bool real_in = first && (S & 1);
#define GSA_EF_PASS(RV) case RV: if (real_in) efast_fft_pass_kernel<RV, true><<<grid, 256, 0, h->stream>>>(f); else efast_fft_pass_kernel<RV, false><<<grid, 256, 0, h->stream>>>(f); break;
            switch (R) { GSA_EF_PASS(2) GSA_EF_PASS(3) GSA_EF_PASS(4) GSA_EF_PASS(5) }
#undef GSA_EF_PASS
            f.in = work[w];
            w ^= 1;
            f.out = work[w];
            f.Ls *= R;
            first = false;
            h->last_launches++;
        }
        EfastPowerArgs pa{f.in, a.power, a.g, M};
        efast_fft_power_kernel<<<dim3((a.g.nbins + 255) / 256, nser), 256, 0, h->stream>>>(pa);
    } else {
        efast_spectrum_kernel<<<dim3((a.g.nbins + EF_THREADS - 1) / EF_THREADS, nser), EF_THREADS, 0, h->stream>>>(a);
    }
    GSA_CUDA(cudaEventRecord(h->evk1, h->stream));
    h->kernel_timed = true;
    efast_indices_kernel<<<nser, 256, 0, h->stream>>>(a);
    h->last_launches += 3;
    GSA_CUDA(cudaGetLastError());
    GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
    std::vector<double> res(nR);
    GSA_CUDA(cudaMemcpyAsync(res.data(), a.S1, sizeof(double) * nR, cudaMemcpyDeviceToHost, h->stream));
    GSA_CUDA(cudaStreamSynchronize(h->stream));
    if (out->S1) memcpy(out->S1, res.data(), sizeof(double) * nser);
    if (out->ST) memcpy(out->ST, res.data() + nser, sizeof(double) * nser);
    return GSA_OK;
}

}  // namespace gsa

using namespace gsa;

extern "C" {

int gsa_efast_design(gsa_handle* h, const double* lb, const double* ub, int d, int64_t samples, int num_harmonics, const double* phi,
                     double* P, int p_location) {
    if (!h || !lb || !ub || !P) return fail(GSA_ERR_INVALID, "NULL argument to gsa_efast_design");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    std::vector<int64_t> omega;
    int rc = efast_frequencies(d, samples, num_harmonics, &omega);
    if (rc) return rc;
    EfastBuffers b;
    if ((rc = efast_upload(h, lb, ub, phi, omega, d, &b))) return rc;
    const size_t total = (size_t)d * (size_t)samples * d;
    double* dP = P;
    if (p_location == GSA_LOC_HOST) {
        if ((rc = h->ybuf.reserve(sizeof(double) * total))) return rc;
        dP = h->ybuf.as<double>();
    }
    h->last_launches = 0;
    GSA_CUDA(cudaEventRecord(h->ev0, h->stream));
    EfastDesignArgs a{b.lb, b.ub, b.phi, b.omega, dP, d, samples};
    const int grid = (int)std::min<int64_t>(((int64_t)samples * d + 255) / 256, (int64_t)h->sm_count * 16);
    efast_design_kernel<<<grid, 256, 0, h->stream>>>(a);
    h->last_launches++;
    GSA_CUDA(cudaGetLastError());
    GSA_CUDA(cudaEventRecord(h->ev1, h->stream));
    if (p_location == GSA_LOC_HOST) {
        GSA_CUDA(cudaMemcpyAsync(P, dP, sizeof(double) * total, cudaMemcpyDeviceToHost, h->stream));
        GSA_CUDA(cudaStreamSynchronize(h->stream));
    }
    return GSA_OK;
}

int gsa_efast_run(gsa_handle* h, int num_harmonics, int model_id, const double* params, int nparams, const double* lb, const double* ub,
                  const double* phi, int d, int64_t samples, gsa_efast_result* out) {
    if (!h || !lb || !ub || !out) return fail(GSA_ERR_INVALID, "NULL argument to gsa_efast_run");
    if (d > EF_DMAX) return fail(GSA_ERR_UNSUPPORTED, "eFAST supports d <= %d (got %d)", EF_DMAX, d);
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    int nout = 0, rc;
    if ((rc = model_nout_entry(h, model_id, nparams, d, &nout))) return rc;
    if (nout > EF_NOMAX) return fail(GSA_ERR_UNSUPPORTED, "eFAST device evaluation supports nout <= %d (got %d): evaluate the model yourself and call gsa_efast_analyze_y", EF_NOMAX, nout);
    std::vector<int64_t> omega;
    if ((rc = efast_frequencies(d, samples, num_harmonics, &omega))) return rc;
    EfastBuffers b;
    if ((rc = efast_upload(h, lb, ub, phi, omega, d, &b))) return rc;
    if ((rc = h->params.reserve(sizeof(double) * std::max(nparams, 1)))) return rc;
    if (nparams > 0) {
        GSA_CUDA(cudaStreamSynchronize(h->stream));
        h->params_cache.assign(params, params + nparams);
        h->aux_kind = -1;
        GSA_CUDA(cudaMemcpyAsync(h->params.p, h->params_cache.data(), sizeof(double) * nparams, cudaMemcpyHostToDevice, h->stream));
        GSA_CUDA(cudaEventRecord(h->ev_params, h->stream));
    }
    const void* kern = efast_eval_kernel_for(h, model_id);
    if (!kern) return fail(GSA_ERR_INVALID, "eFAST: no evaluation kernel for model id %d", model_id);
    const size_t cols = (size_t)samples * d;
    if ((rc = h->dev_in[0].reserve(sizeof(double) * cols * nout))) return rc;
    h->last_launches = 0;
    h->kernel_timed = false;
    GSA_CUDA(cudaEventRecord(h->ev0, h->stream));
    EfastEvalArgs a{b.lb, b.ub, b.phi, b.omega, h->params.as<double>(), h->dev_in[0].as<double>(), d, nparams, nout, samples};
    void* kargs[] = {&a};
    const int grid = (int)std::min<size_t>((cols + 127) / 128, (size_t)h->sm_count * 16);
    GSA_CUDA(cudaLaunchKernel(kern, dim3(grid), dim3(128), kargs, 0, h->stream));
    h->last_launches++;
    return efast_analyze_device(h, h->dev_in[0].as<double>(), nout, d, samples, num_harmonics, out);
}

int gsa_efast_analyze_y(gsa_handle* h, int num_harmonics, const double* Y, int y_location, int nout, int d, int64_t samples,
                        gsa_efast_result* out) {
    if (!h || !Y || !out) return fail(GSA_ERR_INVALID, "NULL argument to gsa_efast_analyze_y");
    if (nout < 1 || d < 1) return fail(GSA_ERR_INVALID, "bad shape");
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    const double* Yd = Y;
    h->last_launches = 0;
    h->kernel_timed = false;
    GSA_CUDA(cudaEventRecord(h->ev0, h->stream));
    if (y_location == GSA_LOC_HOST) {
        const size_t bytes = sizeof(double) * (size_t)nout * (size_t)samples * d;
        int rc = h->dev_in[0].reserve(bytes);
        if (rc) return rc;
        GSA_CUDA(cudaMemcpyAsync(h->dev_in[0].p, Y, bytes, cudaMemcpyHostToDevice, h->stream));
        Yd = h->dev_in[0].as<double>();
    }
    return efast_analyze_device(h, Yd, nout, d, samples, num_harmonics, out);
}

}  // extern "C"
