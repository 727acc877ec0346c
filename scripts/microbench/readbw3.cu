// Does the HBM read rate depend on WHERE concurrently active warps read?  Every warp owns a TMA ring
// (S stages of tile_bytes, own mbarriers, lane 0 issues) as in chain_stream_kernel; mode 0 gives every
// warp a contiguous range of the buffer (1184+ read pointers scattered over 14.7 GB), mode 1 deals the
// tiles round-robin over all warps (the tiles in flight at any time are adjacent in memory), mode 2
// deals them round-robin over the warps of a CTA inside a contiguous per-CTA range.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o readbw3 readbw3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_warp_rings(const char *x, size_t ntiles, int tile_bytes, int S, int mode, int spin, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem) + warp * S;
    unsigned char *ring = smem + 1024 + (size_t)warp * S * tile_bytes;
    if (lane == 0) {
        for (int i = 0; i < S; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nwarps = (size_t)gridDim.x * W, gw = (size_t)blockIdx.x * W + warp;
    size_t first, stride, count;
    if (mode == 0) { first = gw * ntiles / nwarps; count = (gw + 1) * ntiles / nwarps - first; stride = 1; }
    else if (mode == 1) { first = gw; stride = nwarps; count = gw < ntiles ? (ntiles - gw + nwarps - 1) / nwarps : 0; }
    else { const size_t c0 = (size_t)blockIdx.x * ntiles / gridDim.x, c1 = ((size_t)blockIdx.x + 1) * ntiles / gridDim.x;
           first = c0 + warp; stride = W; count = first < c1 ? (c1 - first + W - 1) / W : 0; }
    auto issue = [&](size_t k, int s) {
        const size_t t = first + k * stride;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(smem_u32(ring + (size_t)s * tile_bytes)), "l"(x + t * tile_bytes), "r"(tile_bytes), "r"(smem_u32(&bar[s])) : "memory");
    };
    if (lane == 0) for (int i = 0; i < S && (size_t)i < count; ++i) issue(i, i);
    float acc = 0.f;
    int s = 0; uint32_t par = 0;
    for (size_t k = 0; k < count; ++k) {
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar[s])), "r"(par) : "memory");
        acc += reinterpret_cast<float *>(ring + (size_t)s * tile_bytes)[lane];
        if (spin > 0) { const long long t0 = clock64(); while (clock64() - t0 < spin) {} }
        __syncwarp();
        if (lane == 0 && k + S < count) issue(k + S, s);
        if (++s == S) { s = 0; par ^= 1u; }
    }
    if (acc == 12345.678f) out[0] = acc;
}
int main() {
    const size_t bytes = 14745600000ull;
    char *x; float *out;
    cudaMalloc(&x, bytes); cudaMalloc(&out, 4); cudaMemset(x, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_warp_rings, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    const int tb = 8704;
    struct Cfg { int W, S; };
    const Cfg cfgs[] = {{8, 2}, {8, 3}, {12, 2}, {4, 4}, {16, 1}};
    for (const Cfg &c : cfgs)
        for (int mode = 0; mode < 3; ++mode)
            for (int spin : {0, 1500}) {
                const int smem = 1024 + c.W * c.S * tb;
                if (smem > 227 * 1024) continue;
                const size_t ntiles = bytes / tb;
                float best = 1e9;
                for (int r = 0; r < 4; ++r) {
                    cudaEventRecord(e0);
                    k_warp_rings<<<148, 32 * c.W, smem>>>(x, ntiles, tb, c.S, mode, spin, out);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
                }
                printf("warp rings W=%2d S=%d tile %d B mode %d (%s) spin %4d   %.3f ms  %.1f GB/s  (%s)\n", c.W, c.S, tb, mode,
                       mode == 0 ? "contiguous range per warp" : mode == 1 ? "tiles round-robin over all warps" : "round-robin inside a per-CTA range",
                       spin, best, ntiles * (double)tb / best / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
