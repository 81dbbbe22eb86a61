// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the FP32 forms the channel front kernel
// could use: FFMA reg*reg+reg, FFMA with an immediate multiplier, FFMA with a uniform/constant multiplier, packed
// fma.rn.f32x2, FMUL, FADD, packed add/mul.  16 independent accumulators per thread, 16 warps per SM (4 per scheduler).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_throughput.bin fma_throughput.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ACC = 16;
constexpr int ITER = 4096;
template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, long long *cyc, float b_arg, float c_arg) {
    float a[ACC];
    float b = b_arg + threadIdx.x * 1e-9f, c = c_arg + threadIdx.x * 1e-9f;      // ordinary (non-uniform) registers
    unsigned long long a2[ACC / 2], b2, c2;
#pragma unroll
    for (int i = 0; i < ACC; ++i) a[i] = 0.001f * (i + threadIdx.x);
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(b), "f"(b));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(c2) : "f"(c), "f"(c));
#pragma unroll
    for (int i = 0; i < ACC / 2; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a2[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));                  // 3 registers
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, %1;" : "+f"(a[i]) : "f"(c));                 // immediate multiplier
            if (MODE == 2) asm volatile("fma.rn.f32 %0, %1, 0f3F7FBE77, %0;" : "+f"(a[i]) : "f"(c));                 // x * imm + acc
            if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b_arg), "f"(c));             // uniform / constant multiplier
            if (MODE == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 5) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
            if (MODE == 6 && i < ACC / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(b2), "l"(c2));   // packed, 8 per round
            if (MODE == 7 && i < ACC / 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(c2));
            if (MODE == 8 && i < ACC / 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(b2));
            if (MODE == 9) asm volatile("mul.rn.f32 %0, %0, 0f3F7FBE77;" : "+f"(a[i]));                               // immediate FMUL
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < ACC / 2; ++i) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a2[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, float *dout, long long *dc, int per_round) {
    k<MODE><<<148, 512>>>(dout, dc, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    k<MODE><<<148, 512>>>(dout, dc, 0.999f, 1e-3f);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
    double inst_per_smsp = 4.0 * ITER * per_round;                 // 4 warps per sub-partition
    printf("%-44s %.3f cycles per warp instruction per sub-partition\n", name, (double)h / inst_per_smsp);
}
int main() {
    float *dout; long long *dc;
    cudaMalloc(&dout, 148 * 512 * 4); cudaMalloc(&dc, 148 * 8);
    run<0>("FFMA  reg * reg + reg", dout, dc, ACC);
    run<1>("FFMA  acc * imm + reg", dout, dc, ACC);
    run<2>("FFMA  reg * imm + acc", dout, dc, ACC);
    run<3>("FFMA  acc * uniform + reg", dout, dc, ACC);
    run<4>("FMUL  reg * reg", dout, dc, ACC);
    run<9>("FMUL  reg * imm", dout, dc, ACC);
    run<5>("FADD  reg + reg", dout, dc, ACC);
    run<6>("FFMA2 packed (two FMAs per instruction)", dout, dc, ACC / 2);
    run<7>("FADD2 packed", dout, dc, ACC / 2);
    run<8>("FMUL2 packed", dout, dc, ACC / 2);
    return 0;
}
