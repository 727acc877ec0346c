#include <cstdio>
#include <cuda_runtime.h>
// throughput of FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a
template <int MODE>
__global__ void k(float *out, int iters) {
    float a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    float b0 = 1.0001f, b1 = 0.9999f;
    unsigned long long p0, p1, p2, p3, q, c;
    asm("mov.b64 %0, {%1,%2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1,%2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1,%2};" : "=l"(q) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(b1), "f"(b0));
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                a0 = fmaf(a0, b0, b1); a1 = fmaf(a1, b0, b1); a2 = fmaf(a2, b0, b1); a3 = fmaf(a3, b0, b1);
                a4 = fmaf(a4, b0, b1); a5 = fmaf(a5, b0, b1); a6 = fmaf(a6, b0, b1); a7 = fmaf(a7, b0, b1);
            }
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(q), "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(q), "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(q), "l"(c));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(q), "l"(c));
            }
        }
    }
    if (MODE == 0) out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    else {
        float x, y;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(p0 ^ p1 ^ p2 ^ p3));
        out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
    }
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters); else k<1><<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            // scalar FMAs executed: threads * iters * 64
            double fma = 148.0 * 8 * 256 * iters * 64.0;
            printf("mode %d: %.3f ms  %.2f T scalar-FMA/s  (%.1f TFLOP/s)\n", mode, ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
