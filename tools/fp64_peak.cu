// FP64 peak of the B200 as seen by this repo's kernels: DFMA (CUDA cores) and DMMA
// (mma.sync.aligned.m8n8k4.f64, the only FP64 tensor instruction; tcgen05 has no FP64 kind).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, double a, double b, int iters) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters) {
    double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    double flops = 2.0 * 8 * iters * (double)blocks * threads;
    printf("DFMA: %.3f ms, %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        dmma_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    flops = 2.0 * 8 * 8 * 4 * 4 * iters * (double)blocks * (threads / 32);
    printf("DMMA m8n8k4: %.3f ms, %.2f TFLOP/s\n", ms, flops / ms * 1e-9);
    cudaError_t err = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
