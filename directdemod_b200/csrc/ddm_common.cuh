// Internal helpers shared by the libddemod.so translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/ddemod.h"

namespace ddm {

// ---- error plumbing (thread local message behind ddm_last_error) -----------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define DDM_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            ::ddm::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                             cudaGetErrorString(e__));                                      \
            return DDM_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

#define DDM_REQUIRE(cond, ...)                                                              \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            ::ddm::set_error(__VA_ARGS__);                                                  \
            return DDM_ERR_INVALID;                                                         \
        }                                                                                   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int sm_count(int device);
void *scratch_get(int device, int slot, size_t bytes);   // nullptr (+ error message) on failure
void scratch_release(int device);
// recycled small device blocks for per-handle buffers (current device); pool_free also accepts
// plain cudaMalloc pointers
cudaError_t pool_alloc(void **out, size_t bytes);
void pool_free(void *p);
void pool_release(int device);
template <typename T>
inline cudaError_t pool_alloc_t(T **out, size_t bytes) {
    return pool_alloc(reinterpret_cast<void **>(out), bytes);
}

// ---- device-side primitives ------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    // make the initialised barrier visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    // order earlier generic-proxy smem accesses before later async-proxy (bulk copy) writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// src, dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// exp(-j 2 pi r g) for an integer sample index g and r = r_hi + r_lo turns/sample.
// The product r*g is reduced modulo 1 in float64 with an exact-residual fma, so the phase
// is good to ~1e-9 turns even at g ~ 1e10; the sin/cos themselves are fp32.
__device__ __forceinline__ float2 phase_rotator(double r_hi, double r_lo, long long g) {
    const double gd = static_cast<double>(g);
    const double p = r_hi * gd;
    const double e = fma(r_hi, gd, -p);
    double fr = p - rint(p);
    fr += e + r_lo * gd;
    const float fh = static_cast<float>(fr);
    const float fl = static_cast<float>(fr - static_cast<double>(fh));
    float s, c;
    sincospif(2.0f * fh, &s, &c);
    // first-order correction for the part of the phase fp32 could not hold
    const float d = 6.28318530717958647692f * fl;
    const float c2 = fmaf(-s, d, c);
    const float s2 = fmaf(c, d, s);
    return make_float2(c2, -s2);          // (cos, -sin)  == exp(-j phi)
}

// packed single precision (Blackwell FFMA2): d = a * b + c on (lo, hi) pairs, one issue slot
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b,
                                                    unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack_f32x2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

#endif  // __CUDACC__

}  // namespace ddm
