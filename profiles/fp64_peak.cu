// FP64 FMA peak of the part (BASELINE.md §2 asks for the measured number next to the HBM peak).
// 148 x 8 CTAs x 256 threads, 16 independent DFMA chains per thread, 4096 iterations per chain: DFMA issue-bound by
// construction (no memory traffic).  Prints one JSON line; bench.py / DESIGN.md quote "fp64_tflops".
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/fp64_peak profiles/fp64_peak.cu && profiles/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CHAINS = 16, ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) v[k] = threadIdx.x * 1e-9 + k;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < CHAINS; k++) v[k] = fma(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += v[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}

// same measurement through the tensor cores: mma.sync m8n8k4 FP64 (256 FMA per warp instruction), 8 independent
// accumulator tiles per warp
__global__ void __launch_bounds__(256) k_dmma(double *out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; k++) c[k][0] = c[k][1] = threadIdx.x * 1e-9 + k;
    for (int i = 0; i < ITERS / 4; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 8;
    double *out;
    cudaMalloc(&out, (size_t)grid * 256 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) k_dfma<<<grid, 256>>>(out, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    float best = 1e30f, sum = 0.f;
    const int reps = 20;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        k_dfma<<<grid, 256>>>(out, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
        sum += ms;
    }
    // sustained: back-to-back launches for ~2 s (the power cap pulls the SM clock down, like inside a long CG loop)
    cudaEventRecord(e0);
    int n = 0;
    float total = 0.f;
    while (total < 2000.f) {
        for (int r = 0; r < 50; r++) k_dfma<<<grid, 256>>>(out, 0.999999, 1e-7);
        n += 50;
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&total, e0, e1);
    }
    // DMMA: per thread ITERS/4 * 8 mma, each 8 FMA per lane
    for (int w = 0; w < 3; w++) k_dmma<<<grid, 256>>>(out, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    float bestm = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        k_dmma<<<grid, 256>>>(out, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        bestm = ms < bestm ? ms : bestm;
    }
    const double flopm = 2.0 * 8.0 * 8.0 * (ITERS / 4) * 256.0 * grid;
    const double flop = 2.0 * CHAINS * ITERS * 256.0 * grid;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_tflops\": %.2f, \"fp64_tflops_avg\": %.2f, \"fp64_tflops_sustained\": %.2f, \"fp64_dmma_tflops\": %.2f, "
           "\"how\": \"%d CTAs x 256 threads x %d DFMA chains x %d iterations, best / mean of %d launches (CUDA events), sustained = %d launches back to back\"}\n",
           prop.name, prop.multiProcessorCount, flop / (best * 1e-3) / 1e12, flop / (sum / reps * 1e-3) / 1e12,
           flop * n / (total * 1e-3) / 1e12, flopm / (bestm * 1e-3) / 1e12, grid, CHAINS, ITERS, reps, n);
    return 0;
}
