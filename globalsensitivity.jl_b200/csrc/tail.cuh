// tail.cuh -- the end of a K2+K3 launch, fused into the kernel itself (one launch per step):
//
// What used to follow the fused kernel -- a stage-2 launch that sums the tile records, a launch for the all-reduce, a copy to
// the host -- is done by the kernel's own CTAs as they finish, in a FIXED-ORDER TREE (results do not depend on who arrives when):
//   tile    a CTA that has written the record of tile t signs off at the tile's GROUP (TAIL_G consecutive tiles of a replicate);
//   group   the CTA whose arrival completes a group sums the group's tile records in tile order into a group record -- while the
//           other CTAs are still computing -- and signs off at the REPLICATE;
//   repl.   the CTA that completes a replicate sums its group records in group order into `partials` and signs off at the top;
//   top     the CTA that completes the launch zeroes the replicates this shard does not touch, runs K5 -- pushes the partial sums
//           into every peer's mailbox over NVLink, publishes, waits for the peers' and adds them in rank order (peer.cuh) -- or,
//           on a single GPU, copies them into mapped pinned host memory, and finally raises a flag in mapped pinned host memory:
//           the host sees the result without waiting for the stream to drain.
// Only the last group, the last replicate and the exchange are on the critical path; a single last CTA that summed all 888 x 208
// doubles of config 5 itself took ~100 us (profiles/r2_peer_probe_n2_first.json).  Without an exchange every replicate's finisher
// copies its own record to the host (a single CTA pushing 100 x 208 doubles through PCIe took ~80 us).
// Counters are self-cleaning (whoever completes a level resets its counter), so nothing is memset between launches.
// (SURVEY 5 / VERDICT r1 #3: "fold reduce_tiles and the peer push into the last-arriving CTA of the fused kernel".)
#pragma once
#include "gsa_common.cuh"
#include "peer.cuh"

namespace gsa {

constexpr int TAIL_G = 32;   // tiles per group

struct TailArgs {
    unsigned int* counters;        // nullptr: no fused tail.  [nrep*ngr] groups, then [nrep] replicates, then [1] top
    const double* tile_rec;        // [nrep*tpr][recsz]   (reduce != 0)
    double* grec;                  // [nrep*ngr][recsz] group records (reduce != 0)
    double* partials;              // [nboot][recsz]
    double* host_out;              // mapped pinned host copy of the final sums (may be null)
    unsigned long long* host_flag; // mapped pinned: set to flag_value after host_out is complete (may be null)
    unsigned long long flag_value;
    int reduce;                    // 1: tile records -> partials here; 0: the tiles were written straight into `partials` (tpr == 1)
    int exchange;                  // 1: peer all-reduce of `partials`
    int recsz, nstat, tpr, ngr, rep_first, nrep, nboot, t_lo, t_hi;   // launched tiles: [t_lo, t_hi) of the nrep*tpr local tiles
    int dbg;                       // timing experiments only (GSA_TAIL_DBG): 1 skip the sums, 2 skip the host copies, 4 no system fences
    PeerArgs peer;
};

inline int tail_groups_per_replicate(int tpr) { return (tpr + TAIL_G - 1) / TAIL_G; }
inline size_t tail_counter_count(int nrep, int tpr) { return (size_t)nrep * tail_groups_per_replicate(tpr) + nrep + 1; }

#ifdef __CUDACC__
// sum of `cnt` <= TAIL_G records (stride `stride` doubles, element i of record j at base[j*stride + i]) in record order ->
// out[0..recsz).  All loads of an element are issued before the first add (one L2 round trip per element, not one per record:
// with eight in flight the two levels of config 5 cost ~13 us, with one in flight -- the first version of the double-double
// entries -- ~25 us).
template <int MAXR>
__device__ __forceinline__ void tail_sum_records(const double* base, int cnt, size_t stride, int recsz, int nstat, double* out, double* out2) {
    constexpr int CH = 16;   // loads in flight per thread and round (two rounds cover a group of 32)
    for (int i = threadIdx.x; i < recsz; i += blockDim.x) {
        const int e = i % nstat;
        const bool is_dd = e < STAT_DD;
        if (is_dd && (e & 1)) continue;
        const double* b = base + i;
        if (is_dd) {
            dd s{0.0, 0.0};
#pragma unroll 1
            for (int j0 = 0; j0 < cnt; j0 += CH) {
                double vh[CH], vl[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    const bool ok = j0 + u < cnt;
                    vh[u] = ok ? __ldcg(b + (size_t)(j0 + u) * stride) : 0.0;
                    vl[u] = ok ? __ldcg(b + (size_t)(j0 + u) * stride + 1) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < CH; ++u)
                    if (j0 + u < cnt) dd_add_dd(s, dd{vh[u], vl[u]});
            }
            out[i] = s.hi;
            out[i + 1] = s.lo;
            if (out2) { out2[i] = s.hi; out2[i + 1] = s.lo; }
        } else {
            double s = 0.0;
#pragma unroll 1
            for (int j0 = 0; j0 < cnt; j0 += CH) {
                double v[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) v[u] = j0 + u < cnt ? __ldcg(b + (size_t)(j0 + u) * stride) : 0.0;
#pragma unroll
                for (int u = 0; u < CH; ++u) s += v[u];   // (absent records are +0.0)
            }
            out[i] = s;
            if (out2) out2[i] = s;
        }
    }
}

// arrival at a counter: true for the CTA whose arrival makes it `expected` (which then also resets it).  Every thread of the
// CTA calls this after its global writes; the winner may read what ALL earlier arrivals wrote.  sys: the CTA has written mapped
// host memory that the launch's last arrival will publish.
__device__ __forceinline__ bool tail_arrive(unsigned int* counter, unsigned int expected, int* s_flag, bool sys = false) {
    if (sys) __threadfence_system();
    else __threadfence();   // this thread's writes are visible device-wide before the CTA signs off
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool last = atomicAdd(counter, 1u) == expected - 1;
        if (last) *counter = 0u;   // self-cleaning: nobody else touches this counter any more in this launch
        *s_flag = last;
    }
    __syncthreads();
    const bool last = *s_flag != 0;
    if (last) __threadfence();
    return last;
}

// called by ALL threads of a CTA right after the record of `tile` has been written
__device__ __forceinline__ void fused_tail_tile_done(const TailArgs& t, int tile) {
    __shared__ int s_flag, s_timed_out;
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t count = (size_t)t.nboot * t.recsz;
    double* host = (t.exchange || (t.dbg & 2)) ? nullptr : t.host_out;   // without an exchange every replicate's finisher copies its own record out
    if (t.reduce) {
        const int rl = tile / t.tpr, tr0 = rl * t.tpr, g = (tile - tr0) / TAIL_G;
        const int g0 = tr0 + g * TAIL_G, g1 = min(g0 + TAIL_G, tr0 + t.tpr);                   // the group's tiles
        const unsigned in_group = (unsigned)(min(g1, t.t_hi) - max(g0, t.t_lo));             // ... of which launched
        if (!tail_arrive(t.counters + (size_t)rl * t.ngr + g, in_group, &s_flag)) return;
        const int lo = max(tr0, t.t_lo), hi = min(tr0 + t.tpr, t.t_hi);                        // launched tiles of the replicate
        const int gf = (lo - tr0) / TAIL_G, gl = (hi - 1 - tr0) / TAIL_G;                        // its groups with launched tiles
        double* prec = t.partials + (size_t)(t.rep_first + rl) * t.recsz;
        double* hrec = host ? host + (size_t)(t.rep_first + rl) * t.recsz : nullptr;
        // tiles of the group outside [t_lo, t_hi) hold zero records (memset by the host): summed like the others
        const bool sums = !(t.dbg & 1);
        if (gl == gf) {   // a replicate of one group (many replicates of few tiles): the group's sum IS the replicate's
            if (sums) tail_sum_records<TAIL_G>(t.tile_rec + (size_t)g0 * t.recsz, g1 - g0, t.recsz, t.recsz, t.nstat, prec, hrec);
        } else {
            if (sums) tail_sum_records<TAIL_G>(t.tile_rec + (size_t)g0 * t.recsz, g1 - g0, t.recsz, t.recsz, t.nstat,
                                               t.grec + ((size_t)rl * t.ngr + g) * t.recsz, nullptr);
            if (!tail_arrive(t.counters + (size_t)t.nrep * t.ngr + rl, (unsigned)(gl - gf + 1), &s_flag)) return;
            const double* gbase = t.grec + ((size_t)rl * t.ngr + gf) * t.recsz;
            const int ng = gl - gf + 1;
            if (sums) tail_sum_records<TAIL_G>(gbase, ng, t.recsz, t.recsz, t.nstat, prec, hrec);   // (the host only fuses when at most TAIL_G groups of a replicate have launched tiles)
        }
        if (!tail_arrive(t.counters + (size_t)t.nrep * t.ngr + t.nrep, (unsigned)t.nrep, &s_flag, hrec != nullptr && !(t.dbg & 4))) return;
        // replicates this shard does not touch
        for (size_t idx = tid; idx < count; idx += nt) {
            const int rep = (int)(idx / t.recsz);
            if (rep < t.rep_first || rep >= t.rep_first + t.nrep) {
                t.partials[idx] = 0.0;
                if (host) host[idx] = 0.0;
            }
        }
        __threadfence_block();
        __syncthreads();
    } else {
        // the tile's record went straight into `partials` (one tile per replicate): its CTA copies it out, the last arrival publishes
        if (host) {
            const size_t r0 = (size_t)(t.rep_first + tile) * t.recsz;
            for (int i = tid; i < t.recsz; i += nt) host[r0 + i] = t.partials[r0 + i];
        }
        if (!tail_arrive(t.counters + (size_t)t.nrep * t.ngr + t.nrep, (unsigned)(t.t_hi - t.t_lo), &s_flag, host != nullptr)) return;
        if (host)
            for (size_t idx = tid; idx < count; idx += nt) {
                const int rep = (int)(idx / t.recsz);
                if (rep < t.rep_first + t.t_lo || rep >= t.rep_first + t.t_hi) host[idx] = 0.0;   // (partials itself was zeroed by the host)
            }
    }
    if (t.exchange) peer_exchange_one_cta(t.peer, &s_timed_out);   // writes the world's sums to partials and host_out
    if (t.host_flag) {
        __threadfence_system();
        __syncthreads();
        if (tid == 0) st_release_sys(t.host_flag, t.flag_value);
    }
}
#endif

}  // namespace gsa
