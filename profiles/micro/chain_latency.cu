// Micro-benchmark: cycles per step of the reference DC recurrence v' = fl(x + fl(c v)) for one warp (latency-bound chain).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_latency chain_latency.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(const float2 *x, float2 *out, long long *cyc, float c_arg, int n) {
    __shared__ float2 sx[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sx[i] = x[i];
    __syncwarp();
    float c = c_arg;
    if (MODE == 1) { volatile float cv = c_arg; c = cv; }        // coefficient in an ordinary register
    float2 v = make_float2(0.f, 0.f);
    long long t0 = clock64();
    for (int it = 0; it < n / 1024; ++it) {
#pragma unroll 32
        for (int j = 0; j < 1024; ++j) {
            float2 s = sx[j];
            if (MODE == 2) {                                     // single component only
                v.x = __fadd_rn(s.x, __fmul_rn(c, v.x));
            } else {
                v.x = __fadd_rn(s.x, __fmul_rn(c, v.x));
                v.y = __fadd_rn(s.y, __fmul_rn(c, v.y));
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = v; cyc[blockIdx.x] = t1 - t0; }
}
template <int MODE> void run(const char *name, const float2 *dx, float2 *dout, long long *dc, int n, int blocks) {
    k<MODE><<<blocks, 32>>>(dx, dout, dc, 0.9995f, n);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, 32>>>(dx, dout, dc, 0.9995f, n); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s blocks %4d: %.2f cycles/step (clock64), %.1f ns/step (events), kernel %.1f us\n", name, blocks, (double)h / n, ms * 1e6 / n, ms * 1e3);
}
int main() {
    int n = 1 << 16;
    float2 *dx, *dout; long long *dc;
    cudaMalloc(&dx, 1024 * 8); cudaMalloc(&dout, 4096 * 8); cudaMalloc(&dc, 4096 * 8);
    float2 hx[1024]; for (int i = 0; i < 1024; ++i) hx[i] = make_float2(0.01f * (i % 7), -0.02f * (i % 5));
    cudaMemcpy(dx, hx, sizeof hx, cudaMemcpyHostToDevice);
    for (int blocks : {1, 148, 296}) {
        run<0>("re+im, coefficient uniform", dx, dout, dc, n, blocks);
        run<1>("re+im, coefficient in register", dx, dout, dc, n, blocks);
        run<2>("re only", dx, dout, dc, n, blocks);
    }
    return 0;
}
