// Shared device helpers: error checking, the FP64 tensor-core MMA wrapper, bulk-copy (TMA) and
// mbarrier primitives for sm_100a.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define MRX_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t err__ = (call);                                                                      \
        if (err__ != cudaSuccess) {                                                                      \
            std::fprintf(stderr, "mrx CUDA error %s:%d: %s (%s)\n", __FILE__, __LINE__,                  \
                         cudaGetErrorString(err__), #call);                                              \
            std::abort();                                                                                \
        }                                                                                                \
    } while (0)

#ifdef __CUDACC__ // device helpers: only where the CUDA compiler reads this header (the host drivers also build without it, for the CPU mock of tests/cpp/cuda_mock)
namespace mrx {

// D(8x8) += A(8x4) * B(4x8), FP64 tensor pipe (SASS: DMMA.8x8x4).
// Fragment layout (lane = 4*r + q): a = A[r][q]; b = B[q][r]; c0/c1 = C[r][2q], C[r][2q+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// read-only global load that keeps its program order relative to the (volatile) DMMA statements: used where a batch of
// loads must be issued BEFORE the math that consumes the first of them (the compiler otherwise sinks plain loads
// next to their use to save registers, serialising the memory round trips)
__device__ __forceinline__ double ldg_ordered(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
    return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP) ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy, completion signalled on the mbarrier (bytes multiple of 16, 16B aligned)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

} // namespace mrx
#endif // __CUDACC__
