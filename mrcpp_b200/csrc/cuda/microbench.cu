// Roofline denominators measured on the box itself: FP64 tensor pipe (DMMA m8n8k4), FP64 FMA pipe and
// HBM copy bandwidth. MEASURED_PEAKS.json has no FP64 entry, so bench.py measures it with these
// kernels in the same run.
#include <cuda_runtime.h>

#include <cstdio>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters) {
    // 8 independent accumulator tiles per warp to cover the DMMA latency
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) mrx::dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double4 *__restrict__ in, double4 *__restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

} // namespace

extern "C" double mrx_bench_dmma_tflops(int iters) {
    int dev = 0, sms = 0;
    MRX_CUDA(cudaGetDevice(&dev));
    MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *out;
    int blocks = sms * 8, threads = 256;
    MRX_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    MRX_CUDA(cudaEventCreate(&e0));
    MRX_CUDA(cudaEventCreate(&e1));
    dmma_peak_kernel<<<blocks, threads>>>(out, 16);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        MRX_CUDA(cudaEventRecord(e0));
        dmma_peak_kernel<<<blocks, threads>>>(out, iters);
        MRX_CUDA(cudaEventRecord(e1));
        MRX_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MRX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)blocks * (threads / 32) * (double)iters * 8 * 512.0; // 2*8*8*4 per DMMA
        best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    MRX_CUDA(cudaFree(out));
    return best;
}

extern "C" double mrx_bench_dfma_tflops(int iters) {
    int dev = 0, sms = 0;
    MRX_CUDA(cudaGetDevice(&dev));
    MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *out;
    int blocks = sms * 8, threads = 256;
    MRX_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    MRX_CUDA(cudaEventCreate(&e0));
    MRX_CUDA(cudaEventCreate(&e1));
    dfma_peak_kernel<<<blocks, threads>>>(out, 16);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        MRX_CUDA(cudaEventRecord(e0));
        dfma_peak_kernel<<<blocks, threads>>>(out, iters);
        MRX_CUDA(cudaEventRecord(e1));
        MRX_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MRX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = (double)blocks * threads * (double)iters * 16 * 2.0;
        best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    MRX_CUDA(cudaFree(out));
    return best;
}

extern "C" double mrx_bench_hbm_gbs(long long bytes, int iters) {
    double4 *a, *b;
    size_t n = (size_t)bytes / sizeof(double4);
    MRX_CUDA(cudaMalloc(&a, n * sizeof(double4)));
    MRX_CUDA(cudaMalloc(&b, n * sizeof(double4)));
    MRX_CUDA(cudaMemset(a, 0, n * sizeof(double4)));
    cudaEvent_t e0, e1;
    MRX_CUDA(cudaEventCreate(&e0));
    MRX_CUDA(cudaEventCreate(&e1));
    int dev = 0, sms = 0;
    MRX_CUDA(cudaGetDevice(&dev));
    MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    copy_kernel<<<sms * 16, 256>>>(a, b, n);
    double best = 0.0;
    for (int rep = 0; rep < iters; rep++) {
        MRX_CUDA(cudaEventRecord(e0));
        copy_kernel<<<sms * 16, 256>>>(a, b, n);
        MRX_CUDA(cudaEventRecord(e1));
        MRX_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MRX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = fmax(best, 2.0 * n * sizeof(double4) / (ms * 1e-3) / 1e9);
    }
    MRX_CUDA(cudaFree(a));
    MRX_CUDA(cudaFree(b));
    return best;
}
