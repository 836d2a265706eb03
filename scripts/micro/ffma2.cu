// Micro-benchmark: scalar FFMA vs packed fma.rn.f32x2 throughput on sm_100a (dev tool).
#include <cstdio>
#include <cuda_runtime.h>
#define N_ITERS 4096
__global__ void k_ffma(float* out, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < N_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long x[8], av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
    for (int i = 0; i < 8; ++i) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < N_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k_ffma<<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 148.0 * 8 * 256 * N_ITERS * 8 * 2;
        printf("FFMA : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
        cudaEventRecord(e0); k_ffma2<<<148 * 8, 256>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, 2 * fl / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
