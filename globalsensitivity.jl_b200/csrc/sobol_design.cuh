// sobol_design.cuh -- K1: unrandomised Sobol design matrices, bit-exact with the reference's generator
// (QuasiMonteCarlo.generate_design_matrices -> Sobol.jl; call sites src/sobol_sensitivity.jl:408-413,
// README.md:55-56, test/sobol_method.jl:29-30).
//
// Integer work + an HBM write stream (8 B per coordinate).  Layout = Julia's d x n column-major, i.e.
// the d coordinates of a sample are contiguous, so consecutive THREADS take consecutive coordinates of
// the flattened (point, dim) index and every warp store is one contiguous 256-byte run.  A thread owns
// one dimension j and walks points p, p+P, p+2P, ... (P a power of two): in Gray-code order the Sobol
// integer of index q+P differs from that of q by exactly two direction numbers,
//     x(q+P) = x(q) ^ V[j][lp-1] ^ V[j][lp + ctz(~(q >> lp))]      (lp = log2 P; first term absent if lp=0)
// so after one random-access start (popcount(gray(q)) XORs) each further point costs 2 XORs.
#pragma once
#include "gsa_common.cuh"

namespace gsa {

struct DesignArgs {
    const uint32_t* V;      // [dT][32] direction numbers, MSB aligned
    const double* scale;    // [d]  ub - lb   (rounded once, as Julia's `ub .- lb`)
    const double* lb;       // [d]
    double* out[8];         // up to 8 matrices per launch, each d x ncols (local columns)
    uint64_t first;         // 0-based sequence index of local column 0
    int64_t ncols;
    int d, dT, lp;          // dT = dims handled by this launch (<= 8*d), P = 1 << lp
    int voff[8];            // first dimension (row of V) of each matrix: m*d for a plain design; the p_range design of
                            // gsa(f, ::Sobol, p_range) takes A from block r and B from block nboot + r (reference :414-415)
};

constexpr int DESIGN_THREADS = 256;

__global__ void __launch_bounds__(DESIGN_THREADS) sobol_design_kernel(DesignArgs a) {
    // Direction numbers of the CTA's 256 (point, dim) slots, transposed to [bit][slot] (+1 padding) so that the
    // per-iteration lookup Vs[b][slot] is bank-conflict free (b is almost always warp-uniform).  Reading them from
    // global memory instead costs 32 L1 wavefronts per warp (each lane owns a different 128-byte table row) and
    // capped the kernel at ~2 TB/s (profiles/r1_k1_design_before.txt).
    __shared__ uint32_t Vs[32][DESIGN_THREADS + 1];
    const int64_t P = (int64_t)1 << a.lp;
    const int64_t total = P * a.dT;
    const int64_t f0 = (int64_t)blockIdx.x * DESIGN_THREADS;
    for (int idx = threadIdx.x; idx < 32 * DESIGN_THREADS; idx += DESIGN_THREADS) {
        const int slot = idx >> 5, b = idx & 31;
        const int64_t fs = f0 + slot;
        if (fs < total) {
            const int jj = (int)(fs % a.dT);
            Vs[b][slot] = __ldg(a.V + 32 * (a.voff[jj / a.d] + jj % a.d) + b);
        }
    }
    __syncthreads();
    const int64_t f = f0 + threadIdx.x;
    if (f >= total) return;
    const int64_t p0 = f / a.dT;
    const int j = (int)(f - p0 * a.dT);
    if (p0 >= a.ncols) return;
    const int m = j / a.d, i = j - m * a.d;
    const double sc = __ldg(a.scale + i), lo = __ldg(a.lb + i);
    double* out = a.out[m] + i;

    uint64_t q = a.first + (uint64_t)p0;
    uint32_t x = 0;
    for (uint64_t g = q ^ (q >> 1); g; g &= g - 1) x ^= Vs[__ffsll((long long)g) - 1][threadIdx.x];
    const uint32_t vlow = a.lp > 0 ? Vs[a.lp - 1][threadIdx.x] : 0u;

    for (int64_t p = p0; p < a.ncols; p += P) {
        const double u = (double)x * 0x1p-32;                       // exact
        out[(size_t)a.d * p] = __dadd_rn(__dmul_rn(sc, u), lo);     // (ub-lb)*u + lb, mul and add rounded separately
        const uint64_t h = q >> a.lp;
        const int c = __ffsll((long long)~h) - 1;                   // trailing ones of h
        x ^= vlow ^ Vs[min(a.lp + c, 31)][threadIdx.x];             // (clamp only matters for the unused last update)
        q += (uint64_t)P;
    }
}

// Two adjacent dimensions per thread (d even): the index arithmetic, the trailing-ones scan and the store instruction are
// shared by two coordinates (one 16-byte store), which roughly halves the instructions per point -- the one-dimension kernel
// above is issue-limited (61 % issue-active at 5.2 TB/s, profiles/r1_k1_design_after.txt).
constexpr int DESIGN_THREADS2 = 128;

__global__ void __launch_bounds__(DESIGN_THREADS2) sobol_design_kernel2(DesignArgs a) {
    constexpr int ROW = 2 * DESIGN_THREADS2 + 2;   // even row stride: the (x0, x1) pair of a thread is one aligned 8-byte word
    __shared__ __align__(8) uint32_t Vs[32 * ROW];
    const int64_t P = (int64_t)1 << a.lp;
    const int dT2 = a.dT / 2;
    const int64_t total = P * dT2;
    const int64_t f0 = (int64_t)blockIdx.x * DESIGN_THREADS2;
    for (int idx = threadIdx.x; idx < 32 * 2 * DESIGN_THREADS2; idx += DESIGN_THREADS2) {
        const int slot = idx >> 5, b = idx & 31;
        const int64_t fs = f0 + (slot >> 1);
        if (fs < total) {
            const int jj = 2 * (int)(fs % dT2) + (slot & 1);
            Vs[b * ROW + slot] = __ldg(a.V + 32 * (a.voff[jj / a.d] + jj % a.d) + b);
        }
    }
    __syncthreads();
    const int64_t f = f0 + threadIdx.x;
    if (f >= total) return;
    const int64_t p0 = f / dT2;
    const int j = 2 * (int)(f - p0 * dT2);
    if (p0 >= a.ncols) return;
    const int m = j / a.d, i = j - m * a.d;   // d is even: both dimensions belong to matrix m
    const double sc0 = __ldg(a.scale + i), lo0 = __ldg(a.lb + i), sc1 = __ldg(a.scale + i + 1), lo1 = __ldg(a.lb + i + 1);
    double* out = a.out[m] + i;
    const uint2* Vp = reinterpret_cast<const uint2*>(Vs) + threadIdx.x;   // row b at Vp[b * ROW / 2]

    uint64_t q = a.first + (uint64_t)p0;
    uint2 x = make_uint2(0u, 0u);
    for (uint64_t g = q ^ (q >> 1); g; g &= g - 1) {
        const uint2 v = Vp[(__ffsll((long long)g) - 1) * (ROW / 2)];
        x.x ^= v.x; x.y ^= v.y;
    }
    const uint2 vlow = a.lp > 0 ? Vp[(a.lp - 1) * (ROW / 2)] : make_uint2(0u, 0u);

    for (int64_t p = p0; p < a.ncols; p += P) {
        double2 r;
        r.x = __dadd_rn(__dmul_rn(sc0, (double)x.x * 0x1p-32), lo0);   // (ub-lb)*u + lb, mul and add rounded separately
        r.y = __dadd_rn(__dmul_rn(sc1, (double)x.y * 0x1p-32), lo1);
        *reinterpret_cast<double2*>(out + (size_t)a.d * p) = r;
        const uint64_t h = q >> a.lp;
        const int c = __ffsll((long long)~h) - 1;                       // trailing ones of h
        const uint2 v = Vp[min(a.lp + c, 31) * (ROW / 2)];              // (clamp only matters for the unused last update)
        x.x ^= vlow.x ^ v.x; x.y ^= vlow.y ^ v.y;
        q += (uint64_t)P;
    }
}

}  // namespace gsa
