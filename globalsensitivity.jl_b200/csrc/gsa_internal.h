// gsa_internal.h -- host-side internals of libgsa_cuda (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/gsa_cuda.h"

struct gsa_handle;

namespace gsa {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define GSA_CUDA(call)                                                \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return ::gsa::cuda_fail(e__, #call);  \
    } while (0)

// growable device scratch
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T> T* as() { return static_cast<T*>(p); }
};

// sets the calling thread's current device for a scope and restores the previous one on every return path
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

int finalize_host(const gsa_sobol_opts& o, const double* partials, int nout, int d, int64_t n, gsa_sobol_result* out);
double host_norm_quantile(double p);

// Joe-Kuo table embedded in the library (directions_blob.S)
const uint32_t* joe_kuo_table();
void direction_numbers(int dim_first, int ndim, uint32_t* V /* ndim x 32 */);

}  // namespace gsa

namespace gsa {
struct UserModelEntry {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t k_generic[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [second order][large instantiation]
    int nout = 1;
    int kind = 0;   // GSA_USER_GENERIC | GSA_USER_PRODUCT | GSA_USER_ADDITIVE
    // separable models: the user's image is kept so that the kernel a launch needs can be LTO-linked on first use (the
    // per-coordinate function inlined); the kernels obtained so far, by name
    std::vector<char> image;
    bool try_lto = false;
    std::vector<std::pair<std::string, cudaKernel_t>> sep_kernels;
    std::vector<cudaLibrary_t> lto_libs;
};
int register_user_model(gsa_handle* h, const void* image, size_t len, const char* symbol, int kind, int nout, int* model_id);
int user_separable_kernel(gsa_handle* h, int model_id, int T, int cpl, bool jansen, const void** kernel);
const void* user_generic_kernel(gsa_handle* h, int model_id, bool second, bool large);   // LTO-linked (model inlined) when possible
struct PeerState;                          // peer.cu: mailboxes of the NVLink all-reduce
void peer_release_handle(gsa_handle* h);   // frees h->peer (no-op when absent)
}  // namespace gsa

struct gsa_handle {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr, ev_params = nullptr, ev_switch = nullptr, ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    std::mutex mu;
    gsa::DevBuf tile_rec, partials, params, vtab, lbub, stage[2], dev_in[2], ybuf, y_in[2], aux, gram;
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_cap = 0;
    std::vector<gsa::UserModelEntry> user_models;
    std::vector<double> params_cache;   // host copy of what h->params currently holds (skip re-uploads of identical parameters)
    int aux_kind = -1;                  // what h->aux holds for params_cache (plan kind), -1 = nothing
    std::vector<uint32_t> vhost;  // cached direction numbers [vdims][32]
    int vdims = 0;
    bool kernel_timed = false;   // evk0/evk1 bracket the dominant kernel of the last call
    int last_launches = 0;
    gsa::PeerState* peer = nullptr;
    // fused tail (tail.cuh): self-cleaning arrival counters and group records in device memory, and the result mailbox in mapped
    // pinned host memory: sums + a flag the last CTA raises
    gsa::DevBuf tail_counters, tail_grec;
    size_t tail_counters_zeroed = 0;        // bytes of tail_counters known to be zero
    double* tail_host = nullptr;            // [tail_host_cap] sums, host view
    double* tail_host_dev = nullptr;        // device view of the same memory
    size_t tail_host_cap = 0;
    unsigned long long* tail_flag = nullptr;      // host view
    unsigned long long* tail_flag_dev = nullptr;
    unsigned long long tail_flag_value = 0;       // value the last fused launch raises
};
