// Micro-benchmark (r02): cycles per step of the reference DC recurrence v' = fl(x + fl(c v)), re and im chains of one warp, for
// different ways of issuing the two singly rounded operations.  All forms give bit-identical results (checked).
//   0  FMUL + FADD, samples read as LDS.64
//   1  FMUL + FFMA(x, one, t)            (one = 1.0f held in a register the compiler cannot see through)
//   2  FFMA(c, v, negzero) + FFMA(x, one, t)
//   3  form 0 with LDS.128 (two samples per load)
//   4  form 1 with LDS.128
//   5  form 1, two independent segments per lane (four chains)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dc_chain_forms dc_chain_forms.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ffma_rn(float a, float b, float c) { float d; asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__device__ __forceinline__ float step1(float c, float one, float nz, float x, float v) {
    if (MODE == 0 || MODE == 3) return __fadd_rn(x, __fmul_rn(c, v));
    if (MODE == 1 || MODE == 4 || MODE == 5) return ffma_rn(x, one, __fmul_rn(c, v));
    return ffma_rn(x, one, ffma_rn(c, v, nz));
}
template <int MODE>
__global__ void k(const float2 *x, float2 *out, long long *cyc, float c_arg, float one, float nz, int n) {
    __shared__ __align__(16) float2 sx[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sx[i] = x[i];
    __syncwarp();
    const float c = c_arg;
    float2 v = make_float2(0.f, 0.f), w = make_float2(0.5f, -0.25f);
    long long t0 = clock64();
    for (int it = 0; it < n / 1024; ++it) {
        if (MODE >= 3 && MODE != 5) {
#pragma unroll 16
            for (int j = 0; j < 512; ++j) {
                const float4 q = reinterpret_cast<const float4 *>(sx)[j];
                v.x = step1<MODE>(c, one, nz, q.x, v.x); v.y = step1<MODE>(c, one, nz, q.y, v.y);
                v.x = step1<MODE>(c, one, nz, q.z, v.x); v.y = step1<MODE>(c, one, nz, q.w, v.y);
            }
        } else {
#pragma unroll 32
            for (int j = 0; j < 1024; ++j) {
                const float2 s = sx[j];
                v.x = step1<MODE>(c, one, nz, s.x, v.x); v.y = step1<MODE>(c, one, nz, s.y, v.y);
                if (MODE == 5) { const float2 s2 = sx[(j + 512) & 1023]; w.x = step1<MODE>(c, one, nz, s2.x, w.x); w.y = step1<MODE>(c, one, nz, s2.y, w.y); }
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = v; out[2 * blockIdx.x + 1] = w; cyc[blockIdx.x] = t1 - t0; }
}
template <int MODE> void run(const char *name, const float2 *dx, float2 *dout, long long *dc, int n) {
    for (int warps : {1, 4}) {
        k<MODE><<<1, 32 * warps>>>(dx, dout, dc, 0.9995f, 1.0f, -0.0f, n);
        cudaDeviceSynchronize();
        k<MODE><<<1, 32 * warps>>>(dx, dout, dc, 0.9995f, 1.0f, -0.0f, n);
        cudaDeviceSynchronize();
        long long h; float2 r[2];
        cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(r, dout, 16, cudaMemcpyDeviceToHost);
        printf("%-52s warps/SM %d: %6.2f cycles/step   v = (%.9g, %.9g)\n", name, warps, (double)h / n, r[0].x, r[0].y);
    }
}
int main() {
    int n = 1 << 16;
    float2 *dx, *dout; long long *dc;
    cudaMalloc(&dx, 1024 * 8); cudaMalloc(&dout, 4096 * 8); cudaMalloc(&dc, 4096 * 8);
    float2 hx[1024]; for (int i = 0; i < 1024; ++i) hx[i] = make_float2(0.01f * (i % 7) + 0.013f, -0.02f * (i % 5) - 0.007f);
    cudaMemcpy(dx, hx, sizeof hx, cudaMemcpyHostToDevice);
    run<0>("0 FMUL + FADD, LDS.64", dx, dout, dc, n);
    run<1>("1 FMUL + FFMA(x,1,t), LDS.64", dx, dout, dc, n);
    run<2>("2 FFMA(c,v,-0) + FFMA(x,1,t), LDS.64", dx, dout, dc, n);
    run<3>("3 FMUL + FADD, LDS.128", dx, dout, dc, n);
    run<4>("4 FMUL + FFMA(x,1,t), LDS.128", dx, dout, dc, n);
    run<5>("5 FMUL + FFMA, two segments per lane (per pair of steps)", dx, dout, dc, n);
    return 0;
}
