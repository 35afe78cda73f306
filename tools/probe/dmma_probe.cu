// Probe of the FP64 pipes on sm_100a: DMMA.8x8x4 issue rate/latency vs warps and ILP, and whether DFMA issues
// concurrently with DMMA (separate pipes) or shares the unit. Development tool; not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ILP, int NFMA>
__global__ void probe(double *out, int iters) {
    double c[ILP][2];
    double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i][0] = c[i][1] = 0.0;
#pragma unroll
    for (int i = 0; i < NFMA; i++) f[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NFMA; i++) s += f[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP, int NFMA> void run(int warps, int ctasPerSm, int sms, double *out) {
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<ILP, NFMA><<<sms * ctasPerSm, warps * 32>>>(out, 100);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        probe<ILP, NFMA><<<sms * ctasPerSm, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double nw = (double)sms * ctasPerSm * warps;
    double dm = nw * iters * ILP * 512.0 / (best * 1e-3) / 1e12;
    double fm = nw * iters * NFMA * 64.0 / (best * 1e-3) / 1e12;
    // cycles per DMMA per SMSP at 1.965 GHz
    double cyc = best * 1e-3 * 1.965e9 / ((double)iters * ILP * (warps * ctasPerSm / 4.0));
    printf("ILP %2d NFMA %2d warps/SM %2d : DMMA %6.2f TF/s  DFMA %6.2f TF/s  total %6.2f  (%.1f clk per DMMA per SMSP-slot)\n", ILP, NFMA,
           warps * ctasPerSm, dm, fm, dm + fm, cyc);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
    // latency: 1 warp per SMSP, ILP 1,2,4,8
    run<1, 0>(4, 1, sms, out); run<2, 0>(4, 1, sms, out); run<4, 0>(4, 1, sms, out); run<8, 0>(4, 1, sms, out);
    run<1, 0>(8, 1, sms, out); run<2, 0>(8, 1, sms, out); run<4, 0>(8, 1, sms, out); run<8, 0>(8, 1, sms, out);
    run<2, 0>(16, 1, sms, out); run<4, 0>(16, 1, sms, out); run<8, 0>(8, 4, sms, out);
    // mix
    run<8, 4>(8, 2, sms, out); run<8, 8>(8, 2, sms, out); run<8, 16>(8, 2, sms, out); run<8, 32>(8, 2, sms, out); run<8, 64>(8, 2, sms, out);
    run<1, 32>(8, 2, sms, out);
    return 0;
}
