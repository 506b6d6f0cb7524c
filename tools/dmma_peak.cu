// dmma_peak.cu -- measured FP64 ceilings of one B200: DMMA.8x8x4 alone, DFMA alone, and both interleaved (do they share a pipe?).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/dmma_peak tools/dmma_peak.cu ; run on the GPU box.
// Used to state the roofline of the kernels whose pair sums run on mma.sync.m8n8k4.f64 (linear Gram, second-order pairs).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// MODE 0: NM DMMA per iteration; 1: NF DFMA per iteration; 2: both
template <int NM, int NF>
__global__ void __launch_bounds__(256) k(double* out, int iters, double a, double b) {
    double c[NM > 0 ? NM : 1][2], f[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NM, int NF>
void run(const char* name, double* out, int sms, int ctas_per_sm, int threads) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NM, NF><<<sms * ctas_per_sm, threads>>>(out, 100, 0.999, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NM, NF><<<sms * ctas_per_sm, threads>>>(out, iters, 0.999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)sms * ctas_per_sm * threads / 32;
    const double flop = warps * iters * (NM * 512.0 + NF * 64.0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // cycles per warp instruction per SM sub-partition, at the nominal max clock (an upper bound if the part throttles)
    const double warp_instr_per_smsp = warps * iters * (NM + NF) / (sms * 4.0);
    printf("{\"test\": \"%s\", \"ctas_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"dmma_per_iter\": %d, \"dfma_per_iter\": %d, \"ns_per_warp_instr_per_smsp\": %.3f}\n",
           name, ctas_per_sm, threads, ms, flop / ms / 1e9, NM, NF, ms * 1e6 / warp_instr_per_smsp);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    for (int w : {128, 256}) {
        for (int c : {1, 2, 4}) {
            run<8, 0>("dmma_only", out, sms, c, w);
            run<0, 16>("dfma_only", out, sms, c, w);
            run<6, 16>("dmma6_dfma16", out, sms, c, w);
            run<6, 64>("dmma6_dfma64", out, sms, c, w);
        }
    }
    printf("sms %d\n", sms);
    return 0;
}
