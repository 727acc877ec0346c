// How many bytes must be in flight per SM to reach the read-only HBM ceiling when every tile is also
// CONSUMED for a while?  TMA bulk ring as in chain.cu (one elected thread issues, mbarrier per stage);
// after a tile has landed every thread spins `spin` cycles (the stand-in for the FIR accumulation),
// then the CTA barrier releases the stage.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o readbw2 readbw2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES>
__global__ void k_tma(const char *x, size_t bytes, int tile_bytes, int spin, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    unsigned char *st0 = smem + 128;
    const size_t ntiles = bytes / tile_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t t, int s) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"(smem_u32(st0 + (size_t)s * tile_bytes)), "l"(x + t * tile_bytes), "r"(tile_bytes), "r"(smem_u32(&bar[s])) : "memory");
    };
    size_t tile = blockIdx.x;
    if (threadIdx.x == 0) for (int i = 0; i < STAGES - 1; ++i) { size_t t = tile + (size_t)i * gridDim.x; if (t < ntiles) issue(t, i); }
    float acc = 0.f;
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const size_t nxt = tile + (size_t)(STAGES - 1) * gridDim.x;
        if (threadIdx.x == 0 && nxt < ntiles) issue(nxt, (it + STAGES - 1) % STAGES);
        uint32_t par = (it / STAGES) & 1;
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar[s])), "r"(par) : "memory");
        acc += reinterpret_cast<float *>(st0 + (size_t)s * tile_bytes)[threadIdx.x];
        if (spin > 0) {
            const long long t0 = clock64();
            while (clock64() - t0 < spin) {}
        }
        __syncthreads();
    }
    if (acc == 12345.678f) out[0] = acc;
}
int main() {
    const size_t bytes = 14745600000ull;
    char *x; float *out;
    cudaMalloc(&x, bytes); cudaMalloc(&out, 4); cudaMemset(x, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char *name, auto launch) {
        launch(); cudaDeviceSynchronize();
        float best = 1e9;
        for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        printf("%-52s %.3f ms  %.1f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    struct Cfg { int stages, ctas, tb; };
    const Cfg cfgs[] = {{2, 2, 34816}, {3, 2, 34816}, {2, 3, 34816}, {4, 1, 34816}, {6, 1, 34816}, {4, 2, 17408}, {6, 2, 17408},
                        {3, 2, 34816 - 1360}};
    for (const Cfg &c : cfgs)
        for (int spin : {0, 600, 1200, 2000}) {
            const int smem = 128 + c.stages * c.tb;
            char nm[96]; snprintf(nm, 96, "tma %d stages %d B x %d CTA/SM spin %d", c.stages, c.tb, c.ctas, spin);
            switch (c.stages) {
                case 2: cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                        time(nm, [&] { k_tma<2><<<148 * c.ctas, 128, smem>>>(x, bytes, c.tb, spin, out); }); break;
                case 3: cudaFuncSetAttribute(k_tma<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                        time(nm, [&] { k_tma<3><<<148 * c.ctas, 128, smem>>>(x, bytes, c.tb, spin, out); }); break;
                case 4: cudaFuncSetAttribute(k_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                        time(nm, [&] { k_tma<4><<<148 * c.ctas, 128, smem>>>(x, bytes, c.tb, spin, out); }); break;
                case 6: cudaFuncSetAttribute(k_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                        time(nm, [&] { k_tma<6><<<148 * c.ctas, 128, smem>>>(x, bytes, c.tb, spin, out); }); break;
            }
        }
    return 0;
}
