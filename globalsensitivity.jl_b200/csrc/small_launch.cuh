// small_launch.cuh -- host launcher shared by the small_*.cu translation units.
#pragma once
#include <algorithm>

#include "fused_small.cuh"
#include "gsa_internal.h"

namespace gsa {

constexpr int SMALL_DMAX = 12;   // first order only (fused sources)
constexpr int SMALL_DMAX_Y = 24; // estimator-on-Y, first order: D+2 staged values + 2D accumulators still fit in registers
constexpr int SMALL_DMAX2 = 8;   // with second order (2D+2 values + 2D + D(D-1)/2 accumulators must fit in registers)

// `occupancy_only` != nullptr: just report the resident CTAs per SM of the kernel that would be launched
template <class Src, int D, bool SECOND, bool JANSEN>
int launch_small(gsa_handle* h, SmallArgs<Src>& a, int* occupancy_only = nullptr) {
    auto kern = small_kernel<Src, D, SECOND, JANSEN>;
    constexpr size_t smem = small_smem_bytes<Src>();
    if (smem > (48u << 10)) GSA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nb = 0;
    GSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, SMALL_THREADS, smem));
    if (occupancy_only) {
        *occupancy_only = std::max(1, nb);
        return GSA_OK;
    }
    const int units = (a.t_end - a.t_begin) * a.nout;
    if (units <= 0) return GSA_OK;
    const int grid = std::min(units, std::max(1, nb) * h->sm_count);
    GSA_CUDA(cudaEventRecord(h->evk0, h->stream));
    kern<<<grid, SMALL_THREADS, smem, h->stream>>>(a);
    GSA_CUDA(cudaEventRecord(h->evk1, h->stream));
    h->kernel_timed = true;
    h->last_launches++;
    GSA_CUDA(cudaGetLastError());
    return GSA_OK;
}

// expands CASE(D) for D = 1..12
#define GSA_SMALL_D_CASES(CASE) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
// expands CASE(D) for D = 13..24 (estimator-on-Y, first order only)
#define GSA_SMALL_D_CASES_Y(CASE) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17) CASE(18) CASE(19) CASE(20) CASE(21) CASE(22) CASE(23) CASE(24)

// declared here, defined in small_y.cu / small_ishigami.cu / small_sobolg.cu
int launch_small_y(gsa_handle* h, int d, bool second, int est, const double* Y, int64_t n, int nout, int step, const TileGeom& g,
                   int nstat, double* tile_rec, int* occupancy_only = nullptr);
int launch_small_ishigami(gsa_handle* h, int d, bool second, int est, const double* A, const double* B, double pa, double pb,
                          const TileGeom& g, int t0, int t1, int nstat, double* tile_rec, int* occupancy_only = nullptr);
int launch_small_sobolg(gsa_handle* h, int d, bool second, int est, const double* A, const double* B, const double2* coef,
                        const TileGeom& g, int t0, int t1, int nstat, double* tile_rec, int* occupancy_only = nullptr);

}  // namespace gsa
