// reduce_tiles.cuh -- stage 2: per-tile records -> per-replicate partial sums, fixed summation order.
//   tile_rec [nrep_local*tpr][recsz]  ->  partials[(rep_first + r)][recsz] (+=, so that several
//   chunks of a streamed shard can be folded into the same buffer).
#pragma once
#include "gsa_common.cuh"

namespace gsa {

constexpr int RT_X = 32, RT_Y = 32;

__global__ void __launch_bounds__(RT_X* RT_Y) reduce_tiles_kernel(const double* __restrict__ tile_rec, double* __restrict__ partials,
                                                                  int recsz, int nstat, int tpr, int rep_first) {
    __shared__ double sh_hi[RT_Y][RT_X], sh_lo[RT_Y][RT_X];
    const int i = blockIdx.x * RT_X + threadIdx.x;  // element of the record
    const int rl = blockIdx.y;                      // local replicate
    const bool valid = i < recsz;
    const int e = valid ? i % nstat : STAT_DD;
    const bool is_dd = e < STAT_DD;
    dd s{0.0, 0.0};
    if (valid && !(is_dd && (e & 1))) {
        const double* base = tile_rec + (size_t)rl * tpr * recsz + i;
        if (is_dd) {
            for (int j = threadIdx.y; j < tpr; j += RT_Y) dd_add_dd(s, dd{base[(size_t)j * recsz], base[(size_t)j * recsz + 1]});
        } else {
            // four loads in flight per thread (the loop is latency-bound: one L2 round trip per iteration otherwise)
            int j = threadIdx.y;
            for (; j + 3 * RT_Y < tpr; j += 4 * RT_Y) {
                const double v0 = base[(size_t)j * recsz], v1 = base[(size_t)(j + RT_Y) * recsz];
                const double v2 = base[(size_t)(j + 2 * RT_Y) * recsz], v3 = base[(size_t)(j + 3 * RT_Y) * recsz];
                s.hi += v0; s.hi += v1; s.hi += v2; s.hi += v3;
            }
            for (; j < tpr; j += RT_Y) s.hi += base[(size_t)j * recsz];
        }
    }
    sh_hi[threadIdx.y][threadIdx.x] = s.hi;
    sh_lo[threadIdx.y][threadIdx.x] = s.lo;
    __syncthreads();
    if (threadIdx.y == 0 && valid && !(is_dd && (e & 1))) {
        double* out = partials + (size_t)(rep_first + rl) * recsz + i;
        if (is_dd) {
            dd t{out[0], out[1]};
            for (int y = 0; y < RT_Y; ++y) dd_add_dd(t, dd{sh_hi[y][threadIdx.x], sh_lo[y][threadIdx.x]});
            out[0] = t.hi;
            out[1] = t.lo;
        } else {
            double t = 0.0;
            for (int y = 0; y < RT_Y; ++y) t += sh_hi[y][threadIdx.x];
            out[0] += t;
        }
    }
}

// Few tiles per replicate (many replicates: nboot = 1000 in BASELINE config 2): one THREAD per element of the partials
// buffer, looping over the replicate's tiles -- instead of one 1024-thread CTA per 32 elements of one replicate.
__global__ void __launch_bounds__(256) reduce_tiles_flat_kernel(const double* __restrict__ tile_rec, double* __restrict__ partials,
                                                                int recsz, int nstat, int tpr, int rep_first, int nrep) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nrep * recsz) return;
    const int rl = (int)(idx / recsz), i = (int)(idx - (int64_t)rl * recsz), e = i % nstat;
    const bool is_dd = e < STAT_DD;
    if (is_dd && (e & 1)) return;
    const double* base = tile_rec + (size_t)rl * tpr * recsz + i;
    double* out = partials + (size_t)(rep_first + rl) * recsz + i;
    if (is_dd) {
        dd t{out[0], out[1]};
        for (int j = 0; j < tpr; ++j) dd_add_dd(t, dd{base[(size_t)j * recsz], base[(size_t)j * recsz + 1]});
        out[0] = t.hi;
        out[1] = t.lo;
    } else {
        double t = 0.0;
        for (int j = 0; j < tpr; ++j) t += base[(size_t)j * recsz];
        out[0] += t;
    }
}

}  // namespace gsa
