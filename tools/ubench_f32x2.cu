// Micro-benchmark: issue throughput of packed fp32 pairs (FFMA2 / FADD2) against scalar FMUL + FADD on sm_100a.
// Each thread runs 16 independent accumulator chains of "acc = acc + v*f" (separately rounded), ITER times.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/ubench_f32x2 tools/ubench_f32x2.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(unsigned long long p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(0ULL)); return r; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
constexpr int ITER = 4096;
__global__ void scalar_k(float* out, float v0) {
    float acc[16], f[16];
    for (int i = 0; i < 16; ++i) { acc[i] = 0.f; f[i] = 1.0f + i * 1e-3f + threadIdx.x * 1e-6f; }
    float v = v0;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(v, f[i]));
        v = __fadd_rn(v, 1e-7f);
    }
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void packed_k(float* out, float v0) {
    unsigned long long acc[8], f[8];
    for (int i = 0; i < 8; ++i) { acc[i] = pk2(0.f, 0.f); f[i] = pk2(1.0f + 2 * i * 1e-3f + threadIdx.x * 1e-6f, 1.0f + (2 * i + 1) * 1e-3f + threadIdx.x * 1e-6f); }
    float v = v0;
    for (int it = 0; it < ITER; ++it) {
        const unsigned long long vv = pk2(v, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = add2(acc[i], mul2(vv, f[i]));
        v = __fadd_rn(v, 1e-7f);
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) { float a, b; upk2(acc[i], a, b); s += a; s += b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float hs[2][4], ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); scalar_k<<<148 * 8, 256>>>(out, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); float t_s = ms;
        cudaMemcpy(hs[0], out, 16, cudaMemcpyDeviceToHost);
        cudaEventRecord(e0); packed_k<<<148 * 8, 256>>>(out, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(hs[1], out, 16, cudaMemcpyDeviceToHost);
        const double flops = 148.0 * 8 * 256 * ITER * 16 * 2;
        printf("scalar FMUL+FADD: %.3f ms (%.1f TFLOP/s)   packed FFMA2+FADD2: %.3f ms (%.1f TFLOP/s)   bitwise equal: %d\n",
               t_s, flops / t_s / 1e9, ms, flops / ms / 1e9,
               hs[0][0] == hs[1][0] && hs[0][1] == hs[1][1] && hs[0][2] == hs[1][2]);
    }
    return 0;
}
