// TEST INFRASTRUCTURE: a host-memory stand-in for the few CUDA runtime calls the HOST DRIVERS of the product use
// (csrc/cabi.cpp, csrc/cuda/device_tree.cu), so that those drivers -- sizes, slot arithmetic, pair lists, residency flags,
// refinement loops -- can be executed where no GPU exists (tests/test_host_drivers_mock.py). "Device" memory is malloc'ed host
// memory, streams are synchronous, kernels are replaced by the host functions of mock_kernels.cpp. Never shipped, never on the
// product path: the product library is built by nvcc against the real CUDA runtime.
#pragma once
#include <chrono>
#include <cstddef>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMock = 1 };
struct mock_stream {};
typedef mock_stream *cudaStream_t;
struct mock_event {
    std::chrono::steady_clock::time_point t;
};
typedef mock_event *cudaEvent_t;
typedef void *cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold };
struct int4 {
    int x, y, z, w;
};
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

inline const char *cudaGetErrorString(cudaError_t) { return "mock CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = new mock_stream;
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <typename T> cudaError_t cudaMalloc(T **p, size_t bytes) {
    *p = static_cast<T *>(std::malloc(bytes ? bytes : 1));
    // poison fresh "device" memory: the drivers must not rely on zero-initialised allocations
    if (*p) std::memset(static_cast<void *>(*p), 0xFF, bytes);
    return *p ? cudaSuccess : cudaErrorMock;
}
inline cudaError_t cudaFree(void *p) {
    std::free(p);
    return cudaSuccess;
}
template <typename T> cudaError_t cudaMallocHost(T **p, size_t bytes) {
    *p = static_cast<T *>(std::malloc(bytes ? bytes : 1));
    return *p ? cudaSuccess : cudaErrorMock;
}
inline cudaError_t cudaFreeHost(void *p) {
    std::free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    std::memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
    std::memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemset2DAsync(void *d, size_t pitch, int v, size_t width, size_t height, cudaStream_t) {
    for (size_t r = 0; r < height; r++) std::memset(static_cast<char *>(d) + r * pitch, v, width);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                     cudaStream_t) {
    for (size_t r = 0; r < height; r++) std::memcpy(static_cast<char *>(d) + r * dpitch, static_cast<const char *>(s) + r * spitch, width);
    return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = new mock_event;
    return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *, int) { return cudaErrorMock; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
