// fp64_peak.cu -- measures the FP64 ceilings the FP64-bound kernels are quoted against:
//   DFMA (vector pipe) and DMMA (mma.sync f64 shapes) throughput on the whole GPU, CUDA events, best of 5.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/fp64_peak tools/fp64_peak.cu ; run without arguments.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

template <int CH>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, double x, double y) {
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) acc[i] = fma(acc[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// SHAPE: 0 = m8n8k4, 1 = m16n8k4, 2 = m16n8k8, 3 = m16n8k16
template <int SHAPE, int CH>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, double x) {
    constexpr int NC = SHAPE == 0 ? 2 : 4;
    double c[CH][NC];
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = x + i * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = x - i * 1e-9;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][NC - 2]), "+d"(c[i][NC - 1]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][NC - 2]), "+d"(c[i][NC - 1])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][NC - 2]), "+d"(c[i][NC - 1])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) s += c[i][j];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    double* out;
    CK(cudaMalloc(&out, 8));
    const int sms = p.multiProcessorCount;
    const int blocks = sms * 8, threads = 256;
    const double nthreads = (double)blocks * threads, nwarps = nthreads / 32;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    {
        double ms = time_ms([&] { dfma_kernel<16><<<blocks, threads>>>(out, 1.0000001, 1e-9); });
        printf(", \"dfma_tflops\": %.2f", 2.0 * nthreads * 16 * ITERS / (ms * 1e-3) / 1e12);
    }
    {
        double ms = time_ms([&] { dmma_kernel<0, 8><<<blocks, threads>>>(out, 1.0000001); });
        printf(", \"dmma_m8n8k4_tflops\": %.2f", 2.0 * nwarps * 8 * ITERS * 8 * 8 * 4 / (ms * 1e-3) / 1e12);
    }
    {
        double ms = time_ms([&] { dmma_kernel<1, 8><<<blocks, threads>>>(out, 1.0000001); });
        printf(", \"dmma_m16n8k4_tflops\": %.2f", 2.0 * nwarps * 8 * ITERS * 16 * 8 * 4 / (ms * 1e-3) / 1e12);
    }
    {
        double ms = time_ms([&] { dmma_kernel<2, 8><<<blocks, threads>>>(out, 1.0000001); });
        printf(", \"dmma_m16n8k8_tflops\": %.2f", 2.0 * nwarps * 8 * ITERS * 16 * 8 * 8 / (ms * 1e-3) / 1e12);
    }
    {
        double ms = time_ms([&] { dmma_kernel<3, 8><<<blocks, threads>>>(out, 1.0000001); });
        printf(", \"dmma_m16n8k16_tflops\": %.2f", 2.0 * nwarps * 8 * ITERS * 16 * 8 * 16 / (ms * 1e-3) / 1e12);
    }
    // DFMA and DMMA issued together by different warps of the same SM (do the two share one pipe?)
    printf("}\n");
    return 0;
}
