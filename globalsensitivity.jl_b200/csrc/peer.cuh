// peer.cuh -- device side of K5 over NVLink peer memory, shared by the stand-alone all-reduce kernel (peer.cu) and by the
// fused tail of the K2+K3 kernels (tail.cuh): mailbox layout, system-scope loads/stores, and the exchange itself.
//
// Mailbox of a rank (device memory, exported with cudaIpcGetMemHandle and mapped by every peer):
//     doubles [2][PEER_WORLD_MAX][max_doubles]     slot [parity][source rank]
//     u64     [2][PEER_WORLD_MAX]                  flag [parity][source rank] = epoch of the vector in that slot
// Protocol of one exchange (epoch e, parity e & 1), per rank:
//     1. push the rank's vector into slot [parity][rank] of EVERY peer's mailbox (plain remote stores through NVSwitch)
//     2. __threadfence_system, then a release store of e into flag [parity][rank] of every peer
//     3. wait (acquire loads, bounded) until the own flags of all ranks carry e
//     4. add the world's vectors in RANK ORDER, double-double (hi, lo) entries error-free: bit-identical sums on every rank
// Two slot sets alternate with the parity: a rank can be at most one exchange ahead of a peer that is still adding (it cannot
// pass step 3 of e+1 before that peer has published e+1, which it does only after it has finished reading e).
#pragma once
#include "gsa_common.cuh"

namespace gsa {

constexpr int PEER_WORLD_MAX = 16;

__host__ __device__ inline size_t peer_flags_offset(size_t max_doubles) { return sizeof(double) * 2 * PEER_WORLD_MAX * max_doubles; }
inline size_t peer_mailbox_bytes(size_t max_doubles) { return peer_flags_offset(max_doubles) + sizeof(unsigned long long) * 2 * PEER_WORLD_MAX; }

struct PeerArgs {
    void* const* peer;          // [world] mailboxes
    double* data;               // in: this rank's vector, out: the sum
    double* host_out;           // optional mapped host copy of the sum
    int* status;
    unsigned long long* counter;   // CTAs of this rank that have finished pushing (monotonic over calls)
    unsigned long long counter_base;
    size_t count, max_doubles;
    unsigned long long epoch;
    long long timeout_ns;
    int rank, world, nstat;
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// steps 3 and 4 for the element range [lo, hi) of the vector, by all threads of the calling CTA (after the caller has pushed
// and published).  `timed_out` is a shared-memory word of the CTA, zeroed by the caller before a barrier.
__device__ __forceinline__ void peer_wait_and_add(const PeerArgs& a, size_t lo, size_t hi, int* timed_out) {
    const int par = (int)(a.epoch & 1), tid = threadIdx.x, nt = blockDim.x;
    if (tid < a.world) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(static_cast<const char*>(a.peer[a.rank]) + peer_flags_offset(a.max_doubles));
        const long long t0 = global_ns();
        while (ld_acquire_sys(f + par * PEER_WORLD_MAX + tid) != a.epoch) {
            if (global_ns() - t0 > a.timeout_ns) {
                *timed_out = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (*timed_out) {
        if (tid == 0) *a.status = 1;
        return;
    }
    const double* own = static_cast<const double*>(a.peer[a.rank]) + (size_t)par * PEER_WORLD_MAX * a.max_doubles;
    for (size_t i = lo + tid; i < hi; i += nt) {
        const int e = (int)(i % a.nstat);
        if (e < STAT_DD) {
            if (e & 1) continue;   // the lo word is handled with its hi word
            dd s{0.0, 0.0};
            for (int q = 0; q < a.world; ++q) {
                const double* v = own + (size_t)q * a.max_doubles + i;
                dd_add_dd(s, dd{ld_relaxed_sys(v), ld_relaxed_sys(v + 1)});
            }
            a.data[i] = s.hi;
            a.data[i + 1] = s.lo;
            if (a.host_out) { a.host_out[i] = s.hi; a.host_out[i + 1] = s.lo; }
        } else {
            double s = 0.0;
            for (int q = 0; q < a.world; ++q) s += ld_relaxed_sys(own + (size_t)q * a.max_doubles + i);
            a.data[i] = s;
            if (a.host_out) a.host_out[i] = s;
        }
    }
}

// the whole exchange by ONE CTA (the fused tail): push [0, count), publish, wait, add
__device__ __forceinline__ void peer_exchange_one_cta(const PeerArgs& a, int* timed_out) {
    const int par = (int)(a.epoch & 1), tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) *timed_out = 0;
    for (int p = 0; p < a.world; ++p) {
        double* dst = static_cast<double*>(a.peer[p]) + ((size_t)par * PEER_WORLD_MAX + a.rank) * a.max_doubles;
        for (size_t i = tid; i < a.count; i += nt) dst[i] = __ldcg(a.data + i);   // (written by other CTAs of this launch: read at L2)
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(static_cast<char*>(a.peer[tid]) + peer_flags_offset(a.max_doubles));
        st_release_sys(f + par * PEER_WORLD_MAX + a.rank, a.epoch);
    }
    peer_wait_and_add(a, 0, a.count, timed_out);
}
#endif  // __CUDACC__

}  // namespace gsa
