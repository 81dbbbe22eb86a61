// cubicsdr_b200/csrc/backend.h -- the one place that knows whether we are nvcc (product) or a plain C++ compiler.
//
// PRODUCT: compiled by nvcc for sm_100a; kernels are real __global__ launches, memory is HBM / pinned host.
// SDR_EMU: compiled by g++ into tests/emu/libsdr_b200_emu.so ONLY.  It replays every kernel's phases serially on
//          the CPU so that index arithmetic and host logic can be checked against the oracle in the no-GPU test
//          tier.  It is test scaffolding: the Python package never loads it and it is never timed or shipped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#ifdef SDR_EMU
#define SDR_HD
#define SDR_DEV
#define SDR_FORCEINLINE inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
template <typename T> static inline T sdr_ldg(const T *p) { return *p; }
typedef void *sdr_stream_t;
#else
#include <cuda_runtime.h>
#define SDR_HD __host__ __device__
#define SDR_DEV __device__
#define SDR_FORCEINLINE __forceinline__
template <typename T> static __host__ __device__ __forceinline__ T sdr_ldg(const T *p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
typedef cudaStream_t sdr_stream_t;
#endif

namespace sdrb200 {

// ---- memory / stream shims used by engine.cu -------------------------------------------------------------------
#ifdef SDR_EMU
inline int be_malloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 1; }
inline void be_free(void *p) { free(p); }
inline int be_host_alloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 1; }
inline void be_host_free(void *p) { free(p); }
inline int be_memset(void *p, int v, size_t n, sdr_stream_t) { memset(p, v, n); return 0; }
inline int be_h2d(void *d, const void *h, size_t n, sdr_stream_t) { memcpy(d, h, n); return 0; }
inline int be_d2h(void *h, const void *d, size_t n, sdr_stream_t) { memcpy(h, d, n); return 0; }
inline int be_d2d(void *d, const void *s, size_t n, sdr_stream_t) { memmove(d, s, n); return 0; }
inline int be_sync(sdr_stream_t) { return 0; }
inline int be_malloc_async(void **p, size_t n, sdr_stream_t) { return be_malloc(p, n); }
inline void be_free_async(void *p, sdr_stream_t) { free(p); }
inline int be_d2d_2d(void *d, size_t dpitch, const void *s_, size_t spitch, size_t width, size_t height, sdr_stream_t) {
    for (size_t r = 0; r < height; ++r) memmove((char *)d + r * dpitch, (const char *)s_ + r * spitch, width);
    return 0;
}
inline const char *be_last_error() { return "emu"; }
#else
inline int be_malloc(void **p, size_t n) { return cudaMalloc(p, n ? n : 1) != cudaSuccess; }
inline void be_free(void *p) { cudaFree(p); }
inline int be_host_alloc(void **p, size_t n) { return cudaMallocHost(p, n ? n : 1) != cudaSuccess; }
inline void be_host_free(void *p) { cudaFreeHost(p); }
inline int be_memset(void *p, int v, size_t n, sdr_stream_t s) { return cudaMemsetAsync(p, v, n, s) != cudaSuccess; }
inline int be_h2d(void *d, const void *h, size_t n, sdr_stream_t s) { return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) != cudaSuccess; }
inline int be_d2h(void *h, const void *d, size_t n, sdr_stream_t s) { return cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) != cudaSuccess; }
inline int be_d2d(void *d, const void *s_, size_t n, sdr_stream_t s) { return cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s) != cudaSuccess; }
inline int be_sync(sdr_stream_t s) { return cudaStreamSynchronize(s) != cudaSuccess; }
// stream-ordered allocation: unlike cudaMalloc / cudaFree these never wait for the other streams of the device
inline int be_malloc_async(void **p, size_t n, sdr_stream_t s) { return cudaMallocAsync(p, n ? n : 1, s) != cudaSuccess; }
inline void be_free_async(void *p, sdr_stream_t s) { if (p) cudaFreeAsync(p, s); }
inline int be_d2d_2d(void *d, size_t dpitch, const void *s_, size_t spitch, size_t width, size_t height, sdr_stream_t s) {
    return cudaMemcpy2DAsync(d, dpitch, s_, spitch, width, height, cudaMemcpyDeviceToDevice, s) != cudaSuccess;
}
inline const char *be_last_error() { return cudaGetErrorString(cudaGetLastError()); }
#endif

}  // namespace sdrb200
