// peer.cu -- K5 without NCCL: all-reduce of the per-replicate partial sums over NVLink peer memory, one process per GPU.
//
// What crosses the links is tiny (config 5: 204 doubles, config 2: 17 000), so the cost of the collective is pure latency: a
// library all-reduce is a launch of its own, a protocol with several hops and, from Python, ~0.15 ms of the 2.3 ms step on 8
// GPUs.  Here every rank owns a MAILBOX in device memory, exported with cudaIpcGetMemHandle and mapped by all peers.  ONE small
// kernel per rank (one CTA per record, at most 64), launched on the handle's stream right behind the tile reduction,
//   1. pushes the rank's vector into slot [parity][rank] of EVERY peer's mailbox (plain remote stores through NVSwitch),
//   2. publishes it: __threadfence_system, then (by the CTA that finishes its push last) a release store of the call's epoch
//      into flag [parity][rank] of every peer,
//   3. waits (acquire loads) until its own flags of all ranks carry the epoch,
//   4. adds the world's vectors in RANK ORDER -- double-double (hi, lo) entries with an error-free add -- so every rank holds
//      bit-identical sums, and writes them to the device buffer and, if asked, straight into pinned host memory (no memcpy).
// Two slot sets alternate with the epoch's parity: a rank can be at most one call ahead of a peer that is still adding (it
// cannot pass step 3 of call e+1 before that peer has published e+1, which it does after it has finished reading call e).
// The wait is bounded (GSA_PEER_TIMEOUT_MS, default 10 s); on timeout the kernel raises a status word instead of hanging.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "gsa_common.cuh"
#include "gsa_internal.h"
#include "peer.cuh"
#include "peer_state.h"

namespace gsa {

constexpr int PEER_THREADS = 256;
constexpr int PEER_CTAS_MAX = 64;

// grid = min(records, PEER_CTAS_MAX) CTAs (always co-resident); every CTA owns a contiguous range of whole records
__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(PeerArgs a) {
    __shared__ int timed_out, is_last;
    const int par = (int)(a.epoch & 1), tid = threadIdx.x;
    // this CTA's contiguous range of whole records
    const size_t nrec = a.count / a.nstat, rpc = (nrec + gridDim.x - 1) / gridDim.x;
    const size_t lo = min(a.count, (size_t)blockIdx.x * rpc * a.nstat), hi = min(a.count, ((size_t)blockIdx.x + 1) * rpc * a.nstat);
    if (tid == 0) timed_out = 0;
    // 1. push this CTA's records into every peer
    for (int p = 0; p < a.world; ++p) {
        double* dst = static_cast<double*>(a.peer[p]) + ((size_t)par * PEER_WORLD_MAX + a.rank) * a.max_doubles;
        for (size_t i = lo + tid; i < hi; i += PEER_THREADS) dst[i] = a.data[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. the CTA that finishes its push last publishes the epoch to every peer
    if (tid == 0) is_last = atomicAdd(a.counter, 1ull) == a.counter_base + gridDim.x - 1;
    __syncthreads();
    if (is_last && tid < a.world) {
        __threadfence_system();
        unsigned long long* f = reinterpret_cast<unsigned long long*>(static_cast<char*>(a.peer[tid]) + peer_flags_offset(a.max_doubles));
        st_release_sys(f + par * PEER_WORLD_MAX + a.rank, a.epoch);
    }
    // 3. wait for everybody's vector of this epoch, 4. add in rank order
    peer_wait_and_add(a, lo, hi, &timed_out);
}

// unmap the peers' mailboxes (this rank's own mailbox stays allocated: peers may still have it mapped)
static void peer_unmap(PeerState* ps) {
    for (int r = 0; r < (int)ps->peer.size(); ++r)
        if (ps->peer[r] && r != ps->rank) cudaIpcCloseMemHandle(ps->peer[r]);
    ps->peer.clear();
    if (ps->peer_dev) cudaFree(ps->peer_dev);
    ps->peer_dev = nullptr;
}

static void peer_release(gsa_handle* h) {
    PeerState* ps = h->peer;
    if (!ps) return;
    peer_unmap(ps);
    if (ps->counter_dev) cudaFree(ps->counter_dev);
    if (ps->mailbox) cudaFree(ps->mailbox);
    if (ps->status_host) cudaFreeHost(ps->status_host);
    delete ps;
    h->peer = nullptr;
}

void peer_release_handle(gsa_handle* h) { peer_release(h); }

int peer_next_exchange(gsa_handle* h, double* data_dev, size_t count, int nstat, double* host_out_mapped_dev, PeerArgs* out);

// the stand-alone exchange: one launch on the handle's stream (the caller holds the handle's lock and has set the device)
int peer_allreduce_launch(gsa_handle* h, double* data_dev, size_t count, int nstat, double* host_out_mapped_dev) {
    PeerArgs a;
    int rc = peer_next_exchange(h, data_dev, count, nstat, host_out_mapped_dev, &a);
    if (rc) return rc;
    PeerState* ps = h->peer;
    const int grid = (int)std::min<size_t>(count / (size_t)nstat, (size_t)PEER_CTAS_MAX);
    a.counter = ps->counter_dev;
    a.counter_base = ps->ctas_launched;
    ps->ctas_launched += (unsigned long long)grid;
    peer_allreduce_kernel<<<grid, PEER_THREADS, 0, h->stream>>>(a);
    h->last_launches++;
    GSA_CUDA(cudaGetLastError());
    return GSA_OK;
}

bool peer_connected(const gsa_handle* h) { return h->peer && h->peer->peer_dev && h->peer->world > 1; }

int peer_next_exchange(gsa_handle* h, double* data_dev, size_t count, int nstat, double* host_out_mapped_dev, PeerArgs* out) {
    PeerState* ps = h->peer;
    if (!ps || !ps->peer_dev) return fail(GSA_ERR_INVALID, "peer exchange: peers are not connected (gsa_peer_export / gsa_peer_connect)");
    if (count > ps->max_doubles) return fail(GSA_ERR_INVALID, "peer exchange: %zu doubles exceed the mailbox (%zu)", count, ps->max_doubles);
    if (nstat < STAT_DD || count == 0 || count % (size_t)nstat) return fail(GSA_ERR_INVALID, "peer exchange: count %zu is not a multiple of nstat %d", count, nstat);
    if (*ps->status_host) return fail(GSA_ERR_CUDA, "peer exchange: an earlier exchange timed out waiting for a peer");
    PeerArgs a;
    a.peer = ps->peer_dev;
    a.data = data_dev;
    a.host_out = host_out_mapped_dev;
    a.status = ps->status_dev;
    a.counter = ps->counter_dev;
    a.counter_base = 0;
    a.count = count;
    a.max_doubles = ps->max_doubles;
    a.epoch = ++ps->epoch;
    const char* t = getenv("GSA_PEER_TIMEOUT_MS");
    a.timeout_ns = (long long)(t && atoi(t) > 0 ? atoi(t) : 10000) * 1000000LL;
    a.rank = ps->rank;
    a.world = ps->world;
    a.nstat = nstat;
    *out = a;
    return GSA_OK;
}

}  // namespace gsa

using namespace gsa;

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" {

int gsa_peer_export(gsa_handle* h, size_t max_doubles, void* handle_out) {
    if (!h || !handle_out || max_doubles == 0) return fail(GSA_ERR_INVALID, "bad arguments to gsa_peer_export");
    static_assert(sizeof(cudaIpcMemHandle_t) == GSA_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    std::lock_guard<std::mutex> lk(h->mu);
    DevGuard guard(h->device);
    GSA_CUDA(cudaStreamSynchronize(h->stream));
    peer_release(h);
    PeerState* ps = new (std::nothrow) PeerState();
    if (!ps) return fail(GSA_ERR_NOMEM, "out of host memory");
    h->peer = ps;
    ps->max_doubles = max_doubles;
    const size_t bytes = peer_mailbox_bytes(max_doubles);
    cudaError_t e = cudaMalloc(&ps->mailbox, bytes);
    if (e == cudaSuccess) e = cudaMemset(ps->mailbox, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ps->counter_dev), sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ps->counter_dev, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&ps->status_host), sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *ps->status_host = 0;
        e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&ps->status_dev), ps->status_host, 0);
    }
    cudaIpcMemHandle_t ih;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&ih, ps->mailbox);
    if (e != cudaSuccess) {
        peer_release(h);
        return cuda_fail(e, "gsa_peer_export");
    }
    memcpy(handle_out, &ih, sizeof(ih));
    return GSA_OK;
}

int gsa_peer_connect(gsa_handle* h, int rank, int world, const void* handles) {
    if (!h || !handles) return fail(GSA_ERR_INVALID, "bad arguments to gsa_peer_connect");
    if (!h->peer || !h->peer->mailbox) return fail(GSA_ERR_INVALID, "gsa_peer_connect: call gsa_peer_export first");
    if (world < 1 || world > PEER_WORLD_MAX || rank < 0 || rank >= world) return fail(GSA_ERR_INVALID, "rank %d / world %d out of range (max %d)", rank, world, PEER_WORLD_MAX);
    std::lock_guard<std::mutex> lk(h->mu);
    DevGuard guard(h->device);
    PeerState* ps = h->peer;
    if (ps->peer_dev) return fail(GSA_ERR_INVALID, "gsa_peer_connect: already connected (gsa_peer_export starts a new group)");
    ps->rank = rank;
    ps->world = world;
    ps->peer.assign(world, nullptr);
    ps->peer[rank] = ps->mailbox;
    for (int r = 0; r < world; ++r) {
        if (r == rank) continue;
        cudaIpcMemHandle_t ih;
        memcpy(&ih, static_cast<const char*>(handles) + (size_t)r * GSA_PEER_HANDLE_BYTES, sizeof(ih));
        cudaError_t e = cudaIpcOpenMemHandle(&ps->peer[r], ih, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            ps->peer[r] = nullptr;
            return cuda_fail(e, "cudaIpcOpenMemHandle (peer mailbox)");
        }
    }
    GSA_CUDA(cudaMalloc(reinterpret_cast<void**>(&ps->peer_dev), sizeof(void*) * world));
    GSA_CUDA(cudaMemcpy(ps->peer_dev, ps->peer.data(), sizeof(void*) * world, cudaMemcpyHostToDevice));
    return GSA_OK;
}

int gsa_peer_allreduce(gsa_handle* h, double* data_dev, size_t count, int nstat, double* host_out) {
    if (!h || !data_dev) return fail(GSA_ERR_INVALID, "bad arguments to gsa_peer_allreduce");
    std::lock_guard<std::mutex> lk(h->mu);
    DevGuard guard(h->device);
    double* host_dev = nullptr;
    if (host_out) {
        void* dp = nullptr;
        GSA_CUDA(cudaHostGetDevicePointer(&dp, host_out, 0));   // fails for pageable memory: the caller must pass pinned memory
        host_dev = static_cast<double*>(dp);
    }
    return peer_allreduce_launch(h, data_dev, count, nstat, host_dev);
}

int gsa_peer_status(gsa_handle* h, int* status) {
    if (!h || !status) return fail(GSA_ERR_INVALID, "bad arguments to gsa_peer_status");
    *status = (h->peer && h->peer->status_host) ? *h->peer->status_host : 0;
    return GSA_OK;
}

int gsa_peer_unmap(gsa_handle* h) {
    if (!h) return GSA_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DevGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->peer) peer_unmap(h->peer);
    return GSA_OK;
}

int gsa_peer_disconnect(gsa_handle* h) {
    if (!h) return GSA_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    DevGuard guard(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    peer_release(h);
    return GSA_OK;
}

}  // extern "C"
