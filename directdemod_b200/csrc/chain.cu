// Fused chain kernel: mixer -> real-tap FIR -> integer decimation -> FM discriminator.
//
// Replaces one chunk of the reference's
//   commSignal(fs, x, chunker).offsetFreq(f).filter(fir).bwLim(bw).funcApply(demod_fm().demod)
// (comm.py:63-78, filters.py:64-70, comm.py:118-129, demod_fm.py:40-49; the loop body of
// decode_noaa.py:623).  See DESIGN.md "Fused chain kernel" for the derivation.
//
// Formulation ("input-stationary polyphase blocks").  With output positions
// n_m = off + m*D and y[m] = sum_k b[k] x'[n_m - k] (x' = mixed input), cut the input into
// blocks of D samples, block j = [B_j, B_j + D) with B_j = n_j + s - D + 1 (s in {0,1} makes
// B_j even so every block is 16-byte aligned).  Block j contributes to outputs j..j+Q-1:
//     P_q[j] = sum_a T[q][a] * x'[B_j + a],    T[q][a] = b[q*D + D-1-a-s]  (0 outside [0,K))
//     y[m]   = sum_{q<Q} P_q[m-q],             Q = ceil((K+s)/D)
// One thread owns one block: every input sample is read from shared memory exactly once, by
// exactly one thread, which also applies the mixer to it.  The taps are warp-uniform
// (broadcast loads).  The partial sums are exchanged through a small shared array and the
// discriminator is applied to consecutive y[m].  HBM traffic is the algorithmic minimum
// (8 B in per sample, 4/D B out) plus a Q/(NT-Q) halo re-read that is served by L2.
//
// Staging.  A CTA walks tiles of NT blocks (NT*D contiguous samples).  Each tile is brought
// into a shared-memory ring of kChainStages stages by the TMA engine with 1-D bulk copies
// (cp.async.bulk, SASS UBLKCP) that signal an mbarrier; the first tile's copy is split
// between the carried halo buffer and the chunk.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "ddm_common.cuh"

namespace ddm {

constexpr int kChainThreads = 128;   // blocks (threads) per tile
constexpr int kChainMaxQ = 10;
constexpr int kChainCtasPerSm = 2;
#ifndef DDM_CHAIN_STAGES
#define DDM_CHAIN_STAGES 2
#endif
#ifndef DDM_CHAIN_FFMA2
#define DDM_CHAIN_FFMA2 1
#endif
constexpr int kChainStages = DDM_CHAIN_STAGES;   // TMA ring depth per CTA
// with a 2-deep ring there is shared memory to double-buffer the partial-sum exchange (one
// barrier per tile); the 3-deep ring single-buffers it (two barriers per tile)
constexpr int kChainEBufs = kChainStages >= 3 ? 1 : 2;
constexpr bool kChainPacked = DDM_CHAIN_FFMA2 != 0;
#ifndef DDM_CHAIN_UNROLL
#define DDM_CHAIN_UNROLL 4
#endif
constexpr int kChainUnroll = DDM_CHAIN_UNROLL;   // 4-sample bodies per loop trip

struct ChainParams {
    const void *x;         // chunk, n samples (cf32 or interleaved u8 pairs)
    const void *halo;      // H samples that precede the chunk, same format
    void *halo_out;        // the other halo buffer: receives the last H samples of the chunk (or NULL)
    void *out;             // f32 (FM) or cf32 (IQ)
    const float *taps;     // [Q][DP]
    const float2 *rot;     // [DP] exp(-j 2 pi r a)
    long long n;           // chunk length
    long long n0;          // global index of x[0]
    long long M;           // output positions in this chunk
    long long b0;          // B_0 = off + s + 1 - D: first sample of block 0 (chunk coords)
    long long num_tiles;
    double r_hi, r_lo;
    int D, DP, H, s, has_prev;
    int a_lastq;           // taps of the last partial sum are zero for a < a_lastq (multiple of 4)
    // batch of independent captures (ddm_chain_apply_batch_dev): capture k reads x + k*x_stride
    // samples and writes out + k*out_stride elements; every capture starts from the same state
    long long batch, x_stride, out_stride;
    int J;                 // outputs per tile (CTA-tiled kernel)
    long long stream_warps;   // warps that share the work (warp-autonomous kernel)
    int stages;               // ring depth per warp (warp-autonomous kernel)
    double step_re, step_im;  // exp(-j 2 pi r 32 D): the block rotator's advance per warp tile
};

// The carry of the delay line, done by the kernel that consumes the chunk: the last CTA copies the last H
// raw samples into the handle's second halo buffer (the host swaps the two), so a chunk costs one launch
// and no copy node between consecutive chunks' kernels.
template <bool U8>
__device__ __forceinline__ void save_halo(const ChainParams &P, int tid, int nthreads) {
    if (P.halo_out == nullptr || blockIdx.x != gridDim.x - 1) return;
    if (U8) {
        const uchar2 *src = static_cast<const uchar2 *>(P.x) + (P.n - P.H);
        uchar2 *dst = static_cast<uchar2 *>(P.halo_out);
        for (int i = tid; i < P.H; i += nthreads) dst[i] = src[i];
    } else {
        const float2 *src = static_cast<const float2 *>(P.x) + (P.n - P.H);
        float2 *dst = static_cast<float2 *>(P.halo_out);
        for (int i = tid; i < P.H; i += nthreads) dst[i] = src[i];
    }
}

__device__ __forceinline__ double2 cmuld(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// exp(-j 2 pi r g) in float64 (block rotators of the warp-autonomous kernel, rotators of the general path)
__device__ __forceinline__ double2 phase_rotator_f64(double r_hi, double r_lo, long long g) {
    const double gd = static_cast<double>(g);
    const double p = r_hi * gd;
    const double e = fma(r_hi, gd, -p);
    double fr = p - rint(p);
    fr += e + r_lo * gd;
    double s, c;
    sincospi(2.0 * fr, &s, &c);
    return make_double2(c, -s);
}

// atan2f for the discriminator: |error| <= ~3 ulp like the library function, a third of its
// instructions (one MUFU.RCP, a degree-8 polynomial in t^2 fitted to atan(t)/t on [0, 1], no
// special-case ladder).  atan2(0, 0) = 0 like np.angle(0); NaN/Inf inputs give NaN.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = mn * __frcp_rn(mx);
    t = mx == 0.f ? 0.f : t;
    const float z = t * t;
    float p = 2.931092439e-03f;
    p = fmaf(p, z, -1.639302633e-02f);
    p = fmaf(p, z, 4.321788567e-02f);
    p = fmaf(p, z, -7.548755919e-02f);
    p = fmaf(p, z, 1.066173221e-01f);
    p = fmaf(p, z, -1.420892529e-01f);
    p = fmaf(p, z, 1.999327745e-01f);
    p = fmaf(p, z, -3.333310456e-01f);
    p = fmaf(p, z, 9.999999872e-01f);
    float r = p * t;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = x < 0.f ? 3.14159265358979324f - r : r;
    return copysignf(r, y);
}

// ------------------------------------------------------------------------------------
// The accumulation over the D samples of one block (both fused kernels): Q partial sums
//   P_q = sum_a T[q][a] * mix(x[B + a])
// from the block's raw samples at `sp` in shared memory.
// ------------------------------------------------------------------------------------
template <int Q, bool MIX, bool U8>
__device__ __forceinline__ void chain_accumulate(const unsigned char *sp, const int D, const int DP, const int a_lastq,
                                                 const float *s_taps, const float *s_rx, const float2 *s_ry,
                                                 const float2 *s_c, unsigned long long (&acc)[Q]) {
    // Packed single precision (FFMA2): accumulators are (re, im) pairs, a tap is a
    // broadcast scalar operand, and the complex rotation is
    //   m = x * cos + swap(x) * (sin, -sin)        (swap = the LO_HI operand selector)
    // so a sample costs 2 + Q issue slots instead of 4 + 2Q.
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = 0ULL;

    // raw pair -> mixed sample.  cf32: x rot.  u8: (v + 0.5 (1+j)) rot with v = b - 128.
    auto rotate = [&](unsigned long long X, float rx, float2 ry, float2 c) -> unsigned long long {
        const float2 x = unpack_f32x2(X);
        if (!MIX) return U8 ? fadd2(X, pack_f32x2(0.5f, 0.5f)) : X;
        const unsigned long long m = U8 ? ffma2(X, pack_f32x2(rx, rx), pack_f32x2(c.x, c.y))
                                        : fmul2(X, pack_f32x2(rx, rx));
        return ffma2(pack_f32x2(x.y, x.x), pack_f32x2(ry.x, ry.y), m);
    };
    // The accumulation over the D samples of this thread's block.  ODD (odd D: the blocks of
    // odd threads are only element aligned) is a compile-time tag of this lambda, chosen by
    // one warp-uniform branch per tile, so that the even-D instruction stream carries none of
    // the narrower loads (as a runtime flag inside the loads it cost D = 34 16 %).
    auto accumulate = [&](auto odd_tag) {
        constexpr bool ODD = decltype(odd_tag)::value;
        // two consecutive raw samples starting at block position a (a even) as packed pairs
        auto load2 = [&](int a, unsigned long long &X0, unsigned long long &X1) {
            if (U8) {
                unsigned int w;
                if (ODD) {
                    const unsigned short *h = reinterpret_cast<const unsigned short *>(sp + 2 * a);
                    w = static_cast<unsigned int>(h[0]) | (static_cast<unsigned int>(h[1]) << 16);
                } else {
                    w = *reinterpret_cast<const unsigned int *>(sp + 2 * a);
                }
                const unsigned long long bias = pack_f32x2(-8388736.f, -8388736.f);     // -(2^23 + 128)
                X0 = fadd2(pack_f32x2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)),
                                      __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541))), bias);
                X1 = fadd2(pack_f32x2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)),
                                      __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543))), bias);
            } else if (ODD) {
                X0 = *reinterpret_cast<const unsigned long long *>(sp + 8 * a);
                X1 = *reinterpret_cast<const unsigned long long *>(sp + 8 * a + 8);
            } else {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(sp + 8 * a);
                X0 = v.x;
                X1 = v.y;
            }
        };
        auto load1 = [&](int a) -> unsigned long long {
            if (U8) {
                const unsigned int w = *reinterpret_cast<const unsigned short *>(sp + 2 * a);
                return fadd2(pack_f32x2(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)),
                                        __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541))),
                             pack_f32x2(-8388736.f, -8388736.f));
            }
            return *reinterpret_cast<const unsigned long long *>(sp + 8 * a);
        };
        const int D4 = D & ~3;
        int a = 0;
        // four samples against the first NQ partial sums
        auto body4 = [&](auto nq_tag) {
            constexpr int NQ = decltype(nq_tag)::value;
            unsigned long long X0, X1, X2, X3;
            load2(a, X0, X1);
            load2(a + 2, X2, X3);
            float4 rx = make_float4(1.f, 1.f, 1.f, 1.f);
            float4 ry0 = make_float4(0.f, 0.f, 0.f, 0.f), ry1 = ry0, c0 = ry0, c1 = ry0;
            if (MIX) {
                rx = *reinterpret_cast<const float4 *>(s_rx + a);
                ry0 = *reinterpret_cast<const float4 *>(s_ry + a);
                ry1 = *reinterpret_cast<const float4 *>(s_ry + a + 2);
                if (U8) {
                    c0 = *reinterpret_cast<const float4 *>(s_c + a);
                    c1 = *reinterpret_cast<const float4 *>(s_c + a + 2);
                }
            }
            const unsigned long long M0 = rotate(X0, rx.x, make_float2(ry0.x, ry0.y), make_float2(c0.x, c0.y));
            const unsigned long long M1 = rotate(X1, rx.y, make_float2(ry0.z, ry0.w), make_float2(c0.z, c0.w));
            const unsigned long long M2 = rotate(X2, rx.z, make_float2(ry1.x, ry1.y), make_float2(c1.x, c1.y));
            const unsigned long long M3 = rotate(X3, rx.w, make_float2(ry1.z, ry1.w), make_float2(c1.z, c1.w));
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 t = *reinterpret_cast<const float4 *>(s_taps + q * DP + a);
                acc[q] = ffma2(pack_f32x2(t.x, t.x), M0, acc[q]);
                acc[q] = ffma2(pack_f32x2(t.y, t.y), M1, acc[q]);
                acc[q] = ffma2(pack_f32x2(t.z, t.z), M2, acc[q]);
                acc[q] = ffma2(pack_f32x2(t.w, t.w), M3, acc[q]);
            }
        };
        // the last partial sum only sees the filter's tail: its taps are zero up to a_lastq
        if (Q > 1) {
#pragma unroll(kChainUnroll)
            for (; a < a_lastq; a += 4) body4(std::integral_constant<int, (Q > 1 ? Q - 1 : 1)>());
        }
#pragma unroll(kChainUnroll)
        for (; a < D4; a += 4) body4(std::integral_constant<int, Q>());
        if (!ODD) {
            if (a < D) {                                // D % 4 == 2: one aligned pair
                unsigned long long X0, X1;
                load2(a, X0, X1);
                float2 rx = make_float2(1.f, 1.f);
                float4 ry0 = make_float4(0.f, 0.f, 0.f, 0.f), c0 = ry0;
                if (MIX) {
                    rx = *reinterpret_cast<const float2 *>(s_rx + a);
                    ry0 = *reinterpret_cast<const float4 *>(s_ry + a);
                    if (U8) c0 = *reinterpret_cast<const float4 *>(s_c + a);
                }
                const unsigned long long M0 = rotate(X0, rx.x, make_float2(ry0.x, ry0.y), make_float2(c0.x, c0.y));
                const unsigned long long M1 = rotate(X1, rx.y, make_float2(ry0.z, ry0.w), make_float2(c0.z, c0.w));
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float2 t = *reinterpret_cast<const float2 *>(s_taps + q * DP + a);
                    acc[q] = ffma2(pack_f32x2(t.x, t.x), M0, acc[q]);
                    acc[q] = ffma2(pack_f32x2(t.y, t.y), M1, acc[q]);
                }
            }
        } else {
            for (; a < D; ++a) {                        // the D % 4 samples left over
                const unsigned long long X0 = load1(a);
                float rx = 1.f;
                float2 ry = make_float2(0.f, 0.f), c0 = ry;
                if (MIX) {
                    rx = s_rx[a];
                    ry = s_ry[a];
                    if (U8) c0 = s_c[a];
                }
                const unsigned long long M0 = rotate(X0, rx, ry, c0);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float t = s_taps[q * DP + a];
                    acc[q] = ffma2(pack_f32x2(t, t), M0, acc[q]);
                }
            }
        }
    };
    if (D & 1) accumulate(std::true_type());
    else accumulate(std::false_type());
}

// ------------------------------------------------------------------------------------
// fast path: D >= 2, Q <= 8
// ------------------------------------------------------------------------------------
// IN = DDM_IN_CU8: the tile is staged as raw interleaved u8 (2 B/sample, a quarter of the HBM and
// PCIe traffic of cf32).  The bulk copies need 16-byte alignment, so a tile is copied from the
// 8-sample boundary below its first sample and every thread indexes with that shift.  The bytes
// become floats with one PRMT each (0x4B0000bb = 2^23 + b) and one packed add of -(2^23 + 128);
// the remaining +0.5 of (b - 127.5) is folded into the rotation as the addend 0.5 (1+j) rot[a].
template <int Q, bool MIX, int OUT, int IN>
__global__ void __launch_bounds__(kChainThreads, kChainCtasPerSm)
chain_fused_kernel(const ChainParams P) {
    constexpr int NT = kChainThreads;
    // outputs per tile; QH = NT - J >= Q halo blocks lead every tile.  J = NT - Q for even D; for odd
    // D it is rounded down to even so that every tile starts on an even sample (16-byte TMA source)
    const int J = P.J;
    const int QH = NT - J;
    constexpr bool U8 = IN == DDM_IN_CU8;
    constexpr int ES = U8 ? 2 : 8;                 // bytes per input sample
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int D = P.D, DP = P.DP;
    // ---- shared memory carve-up ----
    constexpr int S = kChainStages;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);                 // S barriers (<= 4)
    float *s_taps = reinterpret_cast<float *>(smem_raw + 32);                // Q*DP taps
    float *s_rx = s_taps + Q * DP;                                           // DP: cos
    float2 *s_ry = reinterpret_cast<float2 *>(s_rx + DP);                    // DP: (sin, -sin) = (-ry, ry)
    float2 *s_c = s_ry + DP;                                                 // DP: 0.5 (1+j) rot (u8 input)
    float2 *s_e = s_c + DP;                                                  // kChainEBufs*Q*NT float2
    const size_t stage_bytes = U8 ? ((static_cast<size_t>(NT) * D * 2 + 32 + 15) & ~static_cast<size_t>(15))
                                  : static_cast<size_t>(NT) * D * sizeof(float2);
    // plain offset arithmetic from the (128-byte aligned) dynamic shared base keeps the pointer in
    // the shared address space: the tile reads must be LDS, not generic loads
    const unsigned fixed_bytes =
        (32u + 4u * (Q * DP + DP) + 8u * (2 * DP + kChainEBufs * Q * NT) + 127u) & ~127u;
    unsigned char *s_stage0 = smem_raw + fixed_bytes;

    // programmatic dependent launch, as in chain_stream_kernel: set-up overlaps the previous launch's tail,
    // every read of the chunk or the halo comes after the wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        for (int i = 0; i < S; ++i) mbar_init(&mbar[i], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < Q * DP; i += NT) s_taps[i] = P.taps[i];
    for (int i = tid; i < DP; i += NT) {
        const float2 r = P.rot[i];          // (cos, -sin) = exp(-j 2 pi r a)
        s_rx[i] = r.x;
        s_ry[i] = make_float2(-r.y, r.y);
        s_c[i] = make_float2(0.5f * (r.x - r.y), 0.5f * (r.x + r.y));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    save_halo<U8>(P, tid, NT);
    __syncthreads();

    const long long n_even = P.n & ~1LL;
    const long long end_all = P.b0 + P.M * D;      // end of the last needed block

    // thread 0: start the TMA copies that fill one stage with tile `tile`
    auto issue = [&](long long gtile, int stage) {
        unsigned char *dst = s_stage0 + stage * stage_bytes;
        long long cap = 0, tile = gtile;           // (a 64-bit division per tile is not free: only for batches)
        if (P.batch > 1) {
            cap = gtile / P.num_tiles;
            tile = gtile - cap * P.num_tiles;
        }
        const long long S0 = P.b0 + (tile * J - QH) * D;           // first sample (may be < 0)
        long long E = S0 + static_cast<long long>(NT) * D;
        if (E > end_all) E = end_all;
        if (U8) {
            const unsigned char *xb = static_cast<const unsigned char *>(P.x) + cap * P.x_stride * 2;
            const unsigned char *hb = static_cast<const unsigned char *>(P.halo);
            const long long S0a = (S0 >= 0 ? S0 : S0 - 7) / 8 * 8;          // floor to 8 samples = 16 B
            const long long Ea = (E >= 0 ? E + 7 : E) / 8 * 8;              // ceil to 8 samples
            const long long n8 = P.n & ~7LL;
            const long long h_end = Ea < 0 ? Ea : 0;
            const long long c_beg = S0a > 0 ? S0a : 0;
            const long long c_end = Ea < n8 ? Ea : n8;
            uint32_t bytes = 0;
            if (S0a < 0) bytes += static_cast<uint32_t>((h_end - S0a) * 2);
            if (c_end > c_beg) bytes += static_cast<uint32_t>((c_end - c_beg) * 2);
            // tail: the last n % 8 samples of the chunk and the pad behind them (zero-tap positions)
            const long long t_beg = c_beg > n8 ? c_beg : n8;
            for (long long i = t_beg; i < E; ++i) {
                dst[(i - S0a) * 2] = i < P.n ? xb[2 * i] : 128;
                dst[(i - S0a) * 2 + 1] = i < P.n ? xb[2 * i + 1] : 128;
            }
            mbar_arrive_expect_tx(&mbar[stage], bytes);
            if (S0a < 0)
                bulk_g2s(dst, hb + (P.H + S0a) * 2, static_cast<uint32_t>((h_end - S0a) * 2), &mbar[stage]);
            if (c_end > c_beg)
                bulk_g2s(dst + (c_beg - S0a) * 2, xb + c_beg * 2, static_cast<uint32_t>((c_end - c_beg) * 2),
                         &mbar[stage]);
            return;
        }
        const float2 *xf = static_cast<const float2 *>(P.x) + cap * P.x_stride;
        const float2 *hf = static_cast<const float2 *>(P.halo);
        uint32_t bytes = 0;
        const long long Eu = E + (E & 1);          // odd D: an odd end is rounded up (16-byte copies)
        const long long h_end = Eu < 0 ? Eu : 0;
        const long long c_beg = S0 > 0 ? S0 : 0;
        const long long c_end = Eu < n_even ? Eu : n_even;
        if (S0 < 0) bytes += static_cast<uint32_t>((h_end - S0) * 8);
        if (c_end > c_beg) bytes += static_cast<uint32_t>((c_end - c_beg) * 8);
        // tail: the odd last sample of the chunk and the zero pad behind it
        long long t_beg = c_beg > n_even ? c_beg : n_even;
        for (long long i = t_beg; i < E; ++i) {
            float2 v = i < P.n ? xf[i] : make_float2(0.f, 0.f);
            reinterpret_cast<float2 *>(dst)[i - S0] = v;
        }
        mbar_arrive_expect_tx(&mbar[stage], bytes);
        if (S0 < 0)
            bulk_g2s(dst, hf + (P.H + S0), static_cast<uint32_t>((h_end - S0) * 8), &mbar[stage]);
        if (c_end > c_beg)
            bulk_g2s(dst + (c_beg - S0) * 8, xf + c_beg, static_cast<uint32_t>((c_end - c_beg) * 8),
                     &mbar[stage]);
    };

    const long long total_tiles = P.num_tiles * P.batch;
    long long gtile = blockIdx.x;
    if (tid == 0) {
        for (int i = 0; i < S - 1; ++i) {
            const long long t = gtile + static_cast<long long>(i) * gridDim.x;
            if (t < total_tiles) issue(t, i);
        }
    }

    for (int it = 0; gtile < total_tiles; gtile += gridDim.x, ++it) {
        const int stage = it % S;
        long long cap = 0, tile = gtile;
        if (P.batch > 1) {
            cap = gtile / P.num_tiles;
            tile = gtile - cap * P.num_tiles;
        }
        const long long nxt = gtile + static_cast<long long>(S - 1) * gridDim.x;
        if (tid == 0 && nxt < total_tiles) {
            // the stage being refilled was consumed in iteration it-1 (all threads are past
            // that iteration's __syncthreads)
            fence_proxy_async();
            issue(nxt, (it + S - 1) % S);
        }
        mbar_wait(&mbar[stage], (it / S) & 1);

        const long long jblk = tile * J - QH + tid;         // this thread's block index
        float2 *e_buf = s_e + (kChainEBufs == 2 ? (it & 1) * (Q * NT) : 0);
        if (jblk < P.M) {
            size_t sp_off = static_cast<size_t>(tid) * D * ES;
            if (U8) {
                const long long S0 = P.b0 + (tile * J - QH) * D;
                const long long S0a = (S0 >= 0 ? S0 : S0 - 7) / 8 * 8;
                sp_off += static_cast<size_t>(S0 - S0a) * 2;         // shift of the aligned copy (even)
            }
            const unsigned char *sp = s_stage0 + stage * stage_bytes + sp_off;
            unsigned long long acc[Q];
            chain_accumulate<Q, MIX, U8>(sp, D, DP, P.a_lastq, s_taps, s_rx, s_ry, s_c, acc);
            float2 w0 = make_float2(1.f, 0.f);
            if (MIX) {
                const long long g = P.n0 + P.b0 + jblk * D;  // global index of the block start
                w0 = phase_rotator(P.r_hi, P.r_lo, g);
            }
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float2 p = unpack_f32x2(acc[q]);
                if (MIX) p = cmul(p, w0);
                e_buf[q * NT + tid] = p;
            }
        }
        __syncthreads();

        // ---- outputs of this tile: m = tile*J + tid, tid < J ----
        const long long m = tile * J + tid;
        if (tid < J && m < P.M) {
            float2 y = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float2 p = e_buf[q * NT + (tid + QH - q)];
                y.x += p.x;
                y.y += p.y;
            }
            if (OUT == DDM_CHAIN_OUT_IQ) {
                reinterpret_cast<float2 *>(P.out)[cap * P.out_stride + m] = y;
            } else if (m > 0 || P.has_prev) {
                float2 yp = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float2 p = e_buf[q * NT + (tid + QH - 1 - q)];
                    yp.x += p.x;
                    yp.y += p.y;
                }
                const float re = fmaf(y.x, yp.x, y.y * yp.y);
                const float im = fmaf(y.y, yp.x, -y.x * yp.y);
                reinterpret_cast<float *>(P.out)[cap * P.out_stride + m - (P.has_prev ? 0 : 1)] = fast_atan2f(im, re);
            }
        }
        if (kChainEBufs == 1) __syncthreads();        // e_buf is reused by the next tile
    }
}

// ------------------------------------------------------------------------------------
// fast path, second generation: warp-autonomous streams.
//
// ncu on chain_fused_kernel (profiles/r02_chain_fused_v1_warpstates.csv) showed the CTA-tiled kernel
// bound by instruction latency, not by its copies: issue slots 43 % busy, stall samples "wait" 33 %,
// short scoreboard 16 %, CTA barrier 9 %, the mbarrier under 2 % -- two warps per scheduler cannot cover
// a 640-instruction tile of which half is per-tile work (rotator sincospif, atan2f, index arithmetic, the
// partial-sum exchange through shared memory behind a barrier, Q halo blocks recomputed per tile).  Here
// the unit of scheduling is the WARP:
//   * every warp owns a contiguous range of 32-block tiles of the stream and a private ring of
//     P.stages stages with its own mbarriers; lane 0 re-arms a stage (one bulk copy for an interior
//     tile) as soon as the warp has read it -- one __syncwarp, no CTA barrier anywhere in the loop.
//     Up to sixteen warps per SM with one stage each (the others cover a warp's reload) or 8 / 12
//     warps with two: the host picks per block length (stream_geometry) from measurements;
//   * the partial sums never touch shared memory: lane l needs P_q of lane l-q -- one shuffle --
//     and the first q lanes take it from the previous tile of the same warp, which every lane
//     still holds in registers (source lane s hands out its current sum when s < 32-q, else its
//     previous one).  Consecutive tiles of a warp are consecutive in the stream, so there is no
//     halo block to recompute (the CTA-tiled kernel re-read and re-accumulated Q of every 128
//     blocks); a range starts with one warm-up tile whose outputs are dropped.
//   * the block rotator is a float64 complex number advanced by one multiplication per tile and
//     recomputed exactly at every 64th tile of a capture (a range that starts between two anchors
//     repeats the multiplications from the anchor: a tile's value does not depend on the partition);
//     atan2f is a degree-8 polynomial with one MUFU.RCP.
// Block geometry, tap tables, accumulation (chain_accumulate) and summation order are those of the first
// kernel; results agree with it to the last bits of the rotator and the arctangent (rms 6e-8 rad).
// LBW: launch bound in warps -- the 16-warp build (Q = 5 only) has to fit 128 registers.
// Tile t of a capture holds blocks 32 (t-1) .. 32 t - 1: tile 0 is the capture's own warm-up tile
// (history from the halo buffer), tile t >= 1 produces outputs m = 32 (t-1) + lane.
// ------------------------------------------------------------------------------------
constexpr int kStreamMaxWarps = 12;                      // warps per CTA (one CTA per SM)
constexpr int kStreamMaxStages = 8;
constexpr int kStreamTile = 32;                          // blocks per warp tile
constexpr int kStreamMinTiles = 4;                       // per warp: bounds the warm-up overhead of short chunks

__host__ __device__ inline size_t stream_stage_bytes(int D, int in_format) {
    return in_format == DDM_IN_CU8 ? ((static_cast<size_t>(kStreamTile) * D * 2 + 32 + 15) & ~static_cast<size_t>(15))
                                   : static_cast<size_t>(kStreamTile) * D * sizeof(float2);
}
__host__ __device__ inline unsigned stream_fixed_bytes(int Q, int DP, int warps, int stages) {
    const unsigned bars = (8u * warps * stages + 127u) & ~127u;
    return bars + ((4u * (Q * DP + DP) + 8u * (2 * DP) + 127u) & ~127u);
}

template <int Q, bool MIX, int OUT, int IN, int LBW = kStreamMaxWarps>
__global__ void __launch_bounds__(32 * LBW, 1)
chain_stream_kernel(const ChainParams P) {
    constexpr int WT = kStreamTile;
    constexpr bool U8 = IN == DDM_IN_CU8;
    constexpr int ES = U8 ? 2 : 8;                 // bytes per input sample
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = P.stages;
    const int nw_cta = blockDim.x >> 5, nthreads = blockDim.x;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = P.D, DP = P.DP;
    // ---- shared memory carve-up: barriers | taps | rotator tables | per-warp rings ----
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw) + warp * S;      // this warp's S barriers
    float *s_taps = reinterpret_cast<float *>(smem_raw + ((8u * nw_cta * S + 127u) & ~127u));   // Q*DP taps
    float *s_rx = s_taps + Q * DP;                                           // DP: cos
    float2 *s_ry = reinterpret_cast<float2 *>(s_rx + DP);                    // DP: (sin, -sin)
    float2 *s_c = s_ry + DP;                                                 // DP: 0.5 (1+j) rot (u8 input)
    const unsigned stage_bytes = static_cast<unsigned>(stream_stage_bytes(D, IN));
    unsigned char *ring = smem_raw + stream_fixed_bytes(Q, DP, nw_cta, S) + static_cast<size_t>(warp) * S * stage_bytes;

    // Programmatic dependent launch: the chunk loops queue this kernel back to back, each launch depending
    // on the one before only through the halo.  The next launch may be scheduled as soon as every CTA of
    // this one has started, so its CTAs take over SMs as they fall free and do their set-up (barriers, tables
    // -- the handle's constants, written at creation) during this launch's tail; everything that reads the
    // chunk, the halo or anything else an earlier kernel may have produced comes after griddepcontrol.wait,
    // which returns when all earlier work in the stream is complete and visible.  Launched without the
    // attribute both instructions are no-ops.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (lane == 0) {
        for (int i = 0; i < S; ++i) mbar_init(&mbar[i], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < Q * DP; i += nthreads) s_taps[i] = P.taps[i];
    for (int i = tid; i < DP; i += nthreads) {
        const float2 r = P.rot[i];          // (cos, -sin) = exp(-j 2 pi r a)
        s_rx[i] = r.x;
        s_ry[i] = make_float2(-r.y, r.y);
        s_c[i] = make_float2(0.5f * (r.x - r.y), 0.5f * (r.x + r.y));
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    save_halo<U8>(P, tid, nthreads);
    __syncthreads();                          // the only CTA barrier: tables and barriers are set up

    // ---- this warp's range of the global tile sequence (captures back to back) ----
    const int NTC = static_cast<int>(P.num_tiles);          // tiles per capture, warm-up tile 0 included
    const long long TT = static_cast<long long>(NTC) * P.batch;
    const long long nwarps = P.stream_warps;
    const long long gw = static_cast<long long>(blockIdx.x) * nw_cta + warp;
    if (gw >= nwarps) return;
    const long long G0 = gw * TT / nwarps, G1 = (gw + 1) * TT / nwarps;
    if (G0 >= G1) return;
    long long cap = G0 / NTC;
    int tile = static_cast<int>(G0 - cap * NTC);
    int skip = 0;                                           // tiles consumed before the first emitted one
    if (tile > 0) {                                         // warm up on the tile before the range
        tile -= 1;
        skip = 1;
    }
    int todo = static_cast<int>(G1 - G0) + skip;            // tiles this warp consumes
    long long icap = cap;                                   // the next tile to issue
    int itile = tile, itodo = todo;

    const long long tile_samples = static_cast<long long>(WT) * D;
    const long long n_even = P.n & ~1LL, n8 = P.n & ~7LL;
    const long long end_all = P.b0 + P.M * D;               // end of the last needed block
    const long long H = P.H;
    const long long lim_fast = U8 ? (end_all < n8 ? end_all : n8) : (end_all < n_even ? end_all : n_even);

    // lane 0: start the bulk copies that fill `stage` with tile t of capture c
    // `first` > 0: a warm-up tile in the interior of a capture, of which only blocks first..31 are needed
    // (their partial sums feed the first emitted tile) -- the copy starts there
    auto issue = [&](long long c, int t, int stage, int first) {
        unsigned char *dst = ring + stage * stage_bytes;
        const long long S0 = P.b0 + (t - 1) * tile_samples;                  // first sample (may be < 0)
        if (U8) {
            const unsigned char *xb = static_cast<const unsigned char *>(P.x) + c * P.x_stride * 2;
            if (S0 >= 0 && ((S0 + tile_samples + 7) & ~7LL) <= lim_fast) {   // interior tile: one aligned copy
                const long long S0a = S0 & ~7LL;
                const long long S1a = (S0 + static_cast<long long>(first) * D) & ~7LL;
                const uint32_t bytes = static_cast<uint32_t>((((S0 + tile_samples + 7) & ~7LL) - S1a) * 2);
                mbar_arrive_expect_tx(&mbar[stage], bytes);
                bulk_g2s(dst + (S1a - S0a) * 2, xb + S1a * 2, bytes, &mbar[stage]);
                return;
            }
            long long E = S0 + tile_samples;
            if (E > end_all) E = end_all;
            uint32_t bytes = 0;
            const unsigned char *hb = static_cast<const unsigned char *>(P.halo);
            const long long S0a = (S0 >= 0 ? S0 : S0 - 7) / 8 * 8;          // floor to 8 samples = 16 B
            const long long Ea = (E >= 0 ? E + 7 : E) / 8 * 8;              // ceil to 8 samples
            const long long h_beg = S0a > -H ? S0a : -H;                    // nothing older than the halo exists
            const long long h_end = Ea < 0 ? Ea : 0;
            const long long c_beg = S0a > 0 ? S0a : 0;
            const long long c_end = Ea < n8 ? Ea : n8;
            if (h_end > h_beg) bytes += static_cast<uint32_t>((h_end - h_beg) * 2);
            if (c_end > c_beg) bytes += static_cast<uint32_t>((c_end - c_beg) * 2);
            // tail: the last n % 8 samples of the chunk and the pad behind them (zero-tap positions)
            const long long t_beg = c_beg > n8 ? c_beg : n8;
            for (long long i = t_beg; i < E; ++i) {
                dst[(i - S0a) * 2] = i < P.n ? xb[2 * i] : 128;
                dst[(i - S0a) * 2 + 1] = i < P.n ? xb[2 * i + 1] : 128;
            }
            mbar_arrive_expect_tx(&mbar[stage], bytes);
            if (h_end > h_beg)
                bulk_g2s(dst + (h_beg - S0a) * 2, hb + (H + h_beg) * 2, static_cast<uint32_t>((h_end - h_beg) * 2),
                         &mbar[stage]);
            if (c_end > c_beg)
                bulk_g2s(dst + (c_beg - S0a) * 2, xb + c_beg * 2, static_cast<uint32_t>((c_end - c_beg) * 2),
                         &mbar[stage]);
            return;
        }
        const float2 *xf = static_cast<const float2 *>(P.x) + c * P.x_stride;
        if (S0 >= 0 && S0 + tile_samples <= lim_fast) {                       // interior tile: one copy
            const unsigned lead = static_cast<unsigned>(first) * D * 8;       // first is even: a multiple of 16 bytes
            mbar_arrive_expect_tx(&mbar[stage], stage_bytes - lead);
            bulk_g2s(dst + lead, xf + S0 + static_cast<long long>(first) * D, stage_bytes - lead, &mbar[stage]);
            return;
        }
        long long E = S0 + tile_samples;
        if (E > end_all) E = end_all;
        uint32_t bytes = 0;
        const float2 *hf = static_cast<const float2 *>(P.halo);
        const long long Eu = E + (E & 1);          // odd D: an odd end is rounded up (16-byte copies)
        const long long h_beg = S0 > -H ? S0 : -H;
        const long long h_end = Eu < 0 ? Eu : 0;
        const long long c_beg = S0 > 0 ? S0 : 0;
        const long long c_end = Eu < n_even ? Eu : n_even;
        if (h_end > h_beg) bytes += static_cast<uint32_t>((h_end - h_beg) * 8);
        if (c_end > c_beg) bytes += static_cast<uint32_t>((c_end - c_beg) * 8);
        // tail: the odd last sample of the chunk and the zero pad behind it
        const long long t_beg = c_beg > n_even ? c_beg : n_even;
        for (long long i = t_beg; i < E; ++i)
            reinterpret_cast<float2 *>(dst)[i - S0] = i < P.n ? xf[i] : make_float2(0.f, 0.f);
        mbar_arrive_expect_tx(&mbar[stage], bytes);
        if (h_end > h_beg)
            bulk_g2s(dst + (h_beg - S0) * 8, hf + (H + h_beg), static_cast<uint32_t>((h_end - h_beg) * 8), &mbar[stage]);
        if (c_end > c_beg)
            bulk_g2s(dst + (c_beg - S0) * 8, xf + c_beg, static_cast<uint32_t>((c_end - c_beg) * 8), &mbar[stage]);
    };

    // blocks of an interior warm-up tile that matter: the last Q (rounded up to even, which keeps the
    // copy's start on a 16-byte boundary for odd D)
    const int warm_first = skip ? WT - (Q + (Q & 1)) : 0;
    for (int i = 0; i < S && itodo > 0; ++i) {
        if (lane == 0) issue(icap, itile, i, i == 0 ? warm_first : 0);
        --itodo;
        if (++itile == NTC) {
            itile = 0;
            ++icap;
        }
    }

    // ---- block rotator exp(-j 2 pi r g), g = global index of this lane's block start.  Exact (float64
    // sincospi of the double-double reduced phase) at every 64th tile of a capture; in between one
    // float64 complex multiplication by the constant per-tile step.  A range that starts between two
    // anchors repeats the multiplications from the anchor, so the value of a tile does not depend on
    // how the stream was partitioned. ----
    double2 wrot = make_double2(1.0, 0.0);
    const double2 wstep = make_double2(P.step_re, P.step_im);
    auto rot_exact = [&](int t) {
        return phase_rotator_f64(P.r_hi, P.r_lo, P.n0 + P.b0 + (static_cast<long long>(t - 1) * WT + lane) * D);
    };
    if (MIX) {
        const int anchor = tile & ~63;
        wrot = rot_exact(anchor);
        for (int t = anchor; t < tile; ++t) wrot = cmuld(wrot, wstep);
    }

    float2 prev[Q];                     // this lane's partial sums of the previous tile
#pragma unroll
    for (int q = 0; q < Q; ++q) prev[q] = make_float2(0.f, 0.f);
    float2 y_last = make_float2(0.f, 0.f);       // this lane's y of the previous tile (lane 31's feeds the discriminator)
    int stage = 0;
    uint32_t parity = 0;

    for (int k = 0; k < todo; ++k) {
        mbar_wait(&mbar[stage], parity);

        const long long jblk = static_cast<long long>(tile - 1) * WT + lane;      // tile 0: blocks -32 .. -1
        unsigned long long acc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[q] = 0ULL;
        if (jblk < P.M && (k > 0 || lane >= warm_first)) {
            unsigned sp_off = static_cast<unsigned>(lane) * D * ES;
            if (U8) {
                const long long S0 = P.b0 + (tile - 1) * tile_samples;
                const long long S0a = (S0 >= 0 ? S0 : S0 - 7) / 8 * 8;
                sp_off += static_cast<unsigned>(S0 - S0a) * 2;               // shift of the aligned copy (even)
            }
            chain_accumulate<Q, MIX, U8>(ring + stage * stage_bytes + sp_off, D, DP, P.a_lastq, s_taps, s_rx, s_ry, s_c,
                                         acc);
        }
        // the stage has been read by every lane: re-arm it with the tile S steps ahead
        __syncwarp();
        if (itodo > 0) {
            if (lane == 0) {
                fence_proxy_async();
                issue(icap, itile, stage, 0);
            }
            --itodo;
            if (++itile == NTC) {
                itile = 0;
                ++icap;
            }
        }
        if (++stage == S) {
            stage = 0;
            parity ^= 1u;
        }
        float2 cur[Q];
        const float2 w0 = make_float2(static_cast<float>(wrot.x), static_cast<float>(wrot.y));
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float2 p = unpack_f32x2(acc[q]);
            cur[q] = MIX ? cmul(p, w0) : p;
        }

        // ---- y[m] = sum_q P_q[m - q]: lane l takes P_q from lane l-q (this tile) or, for l < q, from
        // lane 32+l-q of the previous tile; source lane s therefore offers cur when s < 32-q ----
        float2 y = cur[0];
#pragma unroll
        for (int q = 1; q < Q; ++q) {
            const float2 v = lane < WT - q ? cur[q] : prev[q];
            y.x += __shfl_sync(0xffffffffu, v.x, (lane - q) & 31);
            y.y += __shfl_sync(0xffffffffu, v.y, (lane - q) & 31);
        }
#pragma unroll
        for (int q = 1; q < Q; ++q) prev[q] = cur[q];
        const long long m = jblk;
        const bool emit = tile >= 1 && k >= skip && m < P.M;
        if (OUT == DDM_CHAIN_OUT_IQ) {
            if (emit) reinterpret_cast<float2 *>(P.out)[cap * P.out_stride + m] = y;
        } else {
            // the discriminator needs y[m-1] as well: one more shuffle, lane 0 from lane 31 of the tile before
            const float2 v = lane == 31 ? y_last : y;
            const float2 yp = make_float2(__shfl_sync(0xffffffffu, v.x, (lane - 1) & 31),
                                          __shfl_sync(0xffffffffu, v.y, (lane - 1) & 31));
            y_last = y;
            if (emit && (m > 0 || P.has_prev)) {
                const float re = fmaf(y.x, yp.x, y.y * yp.y);
                const float im = fmaf(y.y, yp.x, -y.x * yp.y);
                reinterpret_cast<float *>(P.out)[cap * P.out_stride + m - (P.has_prev ? 0 : 1)] = fast_atan2f(im, re);
            }
        }
        if (++tile == NTC) {
            tile = 0;
            ++cap;
        }
        if (MIX) wrot = (tile & 63) == 0 ? rot_exact(tile) : cmuld(wrot, wstep);
    }
}

// ------------------------------------------------------------------------------------
// general path (any D, any tap count): one thread per decimated sample, direct form.
// Correct for every configuration; used when the fast path's constraints do not hold.
// ------------------------------------------------------------------------------------
struct GenericParams {
    const void *x;
    const void *halo;
    double2 *y;            // y[m+1] for m = -1..M-1
    const double *taps;    // K taps
    const double2 *ctaps;  // K taps times exp(+j 2 pi r k) (mixer folded in), nullptr without mixer
    long long n, n0, M, off;
    long long virt_before; // chunk-relative index below which history is virtual: the mixed value
                           // is 1.0 there (the all-ones history behind lfilter_zi, filters.py:45)
    double r_hi, r_lo;
    int K, D, H, mix, in_format;
};

__device__ __forceinline__ float2 chain_fetch(const void *x, const void *halo, int H, long long i, int in_format) {
    if (in_format == DDM_IN_CU8) {
        const unsigned char *p = i >= 0 ? static_cast<const unsigned char *>(x) + 2 * i
                                        : static_cast<const unsigned char *>(halo) + 2 * (H + i);
        return make_float2(static_cast<float>(p[0]) - 127.5f, static_cast<float>(p[1]) - 127.5f);
    }
    return i >= 0 ? static_cast<const float2 *>(x)[i] : static_cast<const float2 *>(halo)[H + i];
}


// The mixer commutes into the taps: x'[pos-k] = x[pos-k] rot(n0+pos) exp(+j 2 pi r k), so
//   y[m] = rot(n0 + pos) * sum_k (taps[k] exp(+j 2 pi r k)) x[pos-k]
// with complex taps precomputed in float64 on the host and ONE rotator per output (the first version
// evaluated a double-double reduced sincos per tap and sample: 151 per output at 151 taps).  Virtual
// history (mixed value 1.0) is summed separately and added after the rotation.
static __global__ void chain_generic_y_kernel(const GenericParams P) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx > P.M) return;
    const long long m = idx - 1;
    const long long pos = P.off + m * P.D;          // chunk coordinates, may be negative
    // fp64 accumulation: at D = 1 consecutive outputs differ by a tiny rotation, so the
    // discriminator needs more than fp32's ~1e-7 rad floor to stay within 1e-5 relative
    double ax = 0.0, ay = 0.0, vr = 0.0;
    for (int k = 0; k < P.K; ++k) {
        const long long i = pos - k;
        if (i < -static_cast<long long>(P.H)) break;
        if (i < P.virt_before) {
            vr += P.taps[k];
            continue;
        }
        const float2 v = chain_fetch(P.x, P.halo, P.H, i, P.in_format);
        const double vx = static_cast<double>(v.x), vy = static_cast<double>(v.y);
        if (P.mix) {
            const double2 c = P.ctaps[k];
            ax = fma(c.x, vx, fma(-c.y, vy, ax));
            ay = fma(c.x, vy, fma(c.y, vx, ay));
        } else {
            const double t = P.taps[k];
            ax = fma(t, vx, ax);
            ay = fma(t, vy, ay);
        }
    }
    if (P.mix) {
        const double2 w = phase_rotator_f64(P.r_hi, P.r_lo, P.n0 + pos);
        const double rx = fma(w.x, ax, -w.y * ay), ry = fma(w.x, ay, w.y * ax);
        ax = rx;
        ay = ry;
    }
    P.y[idx] = make_double2(ax + vr, ay);
}

// Tiled form of chain_generic_y_kernel for launches without virtual history: a CTA produces 128
// consecutive outputs from one staged tile of 127 D + K input samples, converted to float64 once
// (the per-thread form converts every sample K / D times, and F2F shares the FP64 pipe with the
// DFMAs) and stored transposed -- sample s' of the tile at [s' mod D][s' div D] -- so that the 32
// lanes of a warp, whose samples are D apart, read consecutive shared-memory words.
constexpr int kGenTile = 128;

static __global__ void __launch_bounds__(kGenTile)
chain_generic_tile_kernel(const GenericParams P, int row_len) {
    extern __shared__ double2 s_tile[];                  // [D][row_len]
    const int j = threadIdx.x;
    const long long idx0 = static_cast<long long>(blockIdx.x) * kGenTile;
    const int D = P.D, K = P.K;
    const long long lo = P.off + (idx0 - 1) * D - (K - 1);          // chunk coordinate of tile sample 0
    const int count = (kGenTile - 1) * D + K;
    for (int sp = j; sp < count; sp += kGenTile) {
        const long long i = lo + sp;
        double2 v = make_double2(0.0, 0.0);
        if (i >= -static_cast<long long>(P.H) && i < P.n) {
            const float2 f = chain_fetch(P.x, P.halo, P.H, i, P.in_format);
            v = make_double2(static_cast<double>(f.x), static_cast<double>(f.y));
        }
        s_tile[(sp % D) * row_len + sp / D] = v;
    }
    __syncthreads();
    const long long idx = idx0 + j;
    if (idx > P.M) return;
    // sample of tap k: s' = j D + e, e = K - 1 - k  ->  row e mod D, column j + e div D
    int r = (K - 1) % D, a = (K - 1) / D;
    double ax = 0.0, ay = 0.0;
    if (P.mix) {
        for (int k = 0; k < K; ++k) {
            const double2 v = s_tile[r * row_len + j + a];
            const double2 c = P.ctaps[k];
            ax = fma(c.x, v.x, fma(-c.y, v.y, ax));
            ay = fma(c.x, v.y, fma(c.y, v.x, ay));
            if (--r < 0) {
                r = D - 1;
                --a;
            }
        }
        const double2 w = phase_rotator_f64(P.r_hi, P.r_lo, P.n0 + P.off + (idx - 1) * D);
        const double rx = fma(w.x, ax, -w.y * ay), ry = fma(w.x, ay, w.y * ax);
        ax = rx;
        ay = ry;
    } else {
        for (int k = 0; k < K; ++k) {
            const double2 v = s_tile[r * row_len + j + a];
            const double t = P.taps[k];
            ax = fma(t, v.x, ax);
            ay = fma(t, v.y, ay);
            if (--r < 0) {
                r = D - 1;
                --a;
            }
        }
    }
    P.y[idx] = make_double2(ax, ay);
}

// D = 1 (filter + discriminator without decimation): consecutive outputs share all but one of their
// samples, so a thread produces four consecutive outputs from a sliding register window -- one
// shared-memory load and one tap load per tap for four outputs instead of one each per output.
// Tile layout: sample s' at [s' mod 4][s' div 4] (the lanes' windows start 4 samples apart).
constexpr int kGen1Outs = 4;

static __global__ void __launch_bounds__(kGenTile)
chain_generic_tile1_kernel(const GenericParams P, int row_len) {
    extern __shared__ double2 s_tile[];                  // [4][row_len]
    constexpr int R = kGen1Outs;
    const int j = threadIdx.x;
    const int K = P.K;
    const long long idx0 = static_cast<long long>(blockIdx.x) * (kGenTile * R);
    const long long lo = P.off + (idx0 - 1) - (K - 1);
    const int count = kGenTile * R - 1 + K;
    for (int sp = j; sp < count; sp += kGenTile) {
        const long long i = lo + sp;
        double2 v = make_double2(0.0, 0.0);
        if (i >= -static_cast<long long>(P.H) && i < P.n) {
            const float2 f = chain_fetch(P.x, P.halo, P.H, i, P.in_format);
            v = make_double2(static_cast<double>(f.x), static_cast<double>(f.y));
        }
        s_tile[(sp & 3) * row_len + (sp >> 2)] = v;
    }
    __syncthreads();
    const long long idx = idx0 + static_cast<long long>(j) * R;
    if (idx > P.M) return;
    // output c of this thread, tap k: s' = 4 j + c + e, e = K - 1 - k
    auto tile_at = [&](int sp) { return s_tile[(sp & 3) * row_len + (sp >> 2)]; };
    double2 w[R];
#pragma unroll
    for (int c = 0; c < R; ++c) w[c] = tile_at(4 * j + c + K - 1);
    double ax[R], ay[R];
#pragma unroll
    for (int c = 0; c < R; ++c) ax[c] = ay[c] = 0.0;
    for (int k = 0; k < K; ++k) {
        if (P.mix) {
            const double2 t = P.ctaps[k];
#pragma unroll
            for (int c = 0; c < R; ++c) {
                ax[c] = fma(t.x, w[c].x, fma(-t.y, w[c].y, ax[c]));
                ay[c] = fma(t.x, w[c].y, fma(t.y, w[c].x, ay[c]));
            }
        } else {
            const double t = P.taps[k];
#pragma unroll
            for (int c = 0; c < R; ++c) {
                ax[c] = fma(t, w[c].x, ax[c]);
                ay[c] = fma(t, w[c].y, ay[c]);
            }
        }
        // slide: next tap looks one sample earlier
#pragma unroll
        for (int c = R - 1; c > 0; --c) w[c] = w[c - 1];
        const int e = K - 2 - k;
        if (e >= 0) w[0] = tile_at(4 * j + e);
    }
#pragma unroll
    for (int c = 0; c < R; ++c) {
        if (idx + c > P.M) break;
        double rx = ax[c], ry = ay[c];
        if (P.mix) {
            const double2 r = phase_rotator_f64(P.r_hi, P.r_lo, P.n0 + P.off + (idx + c - 1));
            rx = fma(r.x, ax[c], -r.y * ay[c]);
            ry = fma(r.x, ay[c], r.y * ax[c]);
        }
        P.y[idx + c] = make_double2(rx, ry);
    }
}

static __global__ void chain_generic_out_kernel(const double2 *y, void *out, long long M, int has_prev,
                                         int out_mode) {
    const long long m = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (m >= M) return;
    const double2 c = y[m + 1];
    if (out_mode == DDM_CHAIN_OUT_IQ) {
        reinterpret_cast<float2 *>(out)[m] = make_float2(static_cast<float>(c.x), static_cast<float>(c.y));
        return;
    }
    if (m == 0 && !has_prev) return;
    const double2 p = y[m];
    const double re = fma(c.x, p.x, c.y * p.y);
    const double im = fma(c.y, p.x, -c.x * p.y);
    reinterpret_cast<float *>(out)[m - (has_prev ? 0 : 1)] = static_cast<float>(atan2(im, re));
}

}  // namespace ddm

// ======================================================================================
// host side
// ======================================================================================
struct ddm_chain {
    int device = 0;
    int K = 0, D = 1, out_mode = 0, in_format = 0;
    double freq = 0, fs = 1;
    bool mix = false, fast = false;
    double r_hi = 0, r_lo = 0;
    int64_t n0 = 0, dec_off = 0;
    int has_prev = 0;
    int H = 0, DP = 0;
    int Q[2] = {0, 0};
    int a_lastq[2] = {0, 0};
    int per_sm[2] = {0, 0};                  // resident CTAs per SM of the CTA-tiled kernel, per s
    bool stream = false;                     // warp-autonomous kernel (chain_stream_kernel) usable
    int st_warps = 0, st_stages = 0;         // its geometry: warps per CTA (one CTA per SM), ring depth per warp
    bool st_attr[2] = {false, false};        // dynamic shared memory attribute set, per s
    float *d_taps[2] = {nullptr, nullptr};   // [Q][DP] for s = 0, 1
    double *d_taps_lin = nullptr;            // K
    double2 *d_ctaps = nullptr;              // K: taps[k] exp(+j 2 pi r k), general path with mixer
    float2 *d_rot = nullptr;                 // DP
    void *d_halo[2] = {nullptr, nullptr};    // H samples of raw input in the handle's input format
    void *d_halo_init = nullptr;             // the reference's initial condition as raw history (cf32)
    int es = 8;                              // bytes per input sample
    long long n_real = 0;                    // real samples in the halo (u8: the rest is virtual)
    int cur = 0;
    bool pdl = true;                         // programmatic dependent launch of the fused kernels (DDM_CHAIN_NO_PDL: A/B)
    bool chained = false;                    // the handle's previous call was a chunk of the same stream (chunk loop)
    bool halo_in_kernel = true;              // the fused kernel carries the halo itself (DDM_CHAIN_HALO_MEMCPY: A/B switch)
    double step_re = 1, step_im = 0;         // block rotator advance of the warp-autonomous kernel (set once)
    bool step_set = false;
    double2 *d_ytmp = nullptr;
    size_t ytmp_cap = 0;
    void *d_in = nullptr, *d_out = nullptr;  // staging for the _host entry point
    size_t in_cap = 0, out_cap = 0;
    cudaStream_t copy_stream = nullptr;      // ... whose host->device copies run ahead of the kernels, piece by piece
    cudaEvent_t ev_h2d[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<double> taps;
    int sms = 148;
};

namespace {

using namespace ddm;

size_t chain_smem_bytes(int Q, int D, int DP, int in_format = DDM_IN_CF32) {
    size_t fixed = 32 + sizeof(float) * (Q * DP + DP) + sizeof(float2) * 2 * DP +
                   sizeof(float2) * kChainEBufs * Q * kChainThreads;
    fixed = (fixed + 127) & ~static_cast<size_t>(127);
    const size_t stage = in_format == DDM_IN_CU8
                             ? ((static_cast<size_t>(kChainThreads) * D * 2 + 32 + 15) & ~static_cast<size_t>(15))
                             : static_cast<size_t>(kChainThreads) * D * sizeof(float2);
    return fixed + kChainStages * stage;
}

// Geometry of the warp-autonomous kernel for one (D, input format): W warps (one CTA per SM), each with
// a private ring of S stages of 32 blocks.  Measured on B200 (profiles/r02_chain_ab*.jsonl,
// r02_chain_dsweep.txt, scripts/microbench/readbw3.cu): the kernel is bound by instruction latency, not
// by bytes in flight -- a read-only ring reaches 7.4 TB/s with ~140 KB of stages per SM and gets SLOWER
// with 200 KB -- so the rule is: as many warps as possible in multiples of four (equal load on the four
// schedulers), rings of at most ~150 KB per SM, a second stage per warp only if it fits under that.
// Returns false where the CTA-tiled kernel is the better tool (blocks of 19 KB and more, D >= 76 for
// cf32, which it stages up to D = 110) -- beyond that this kernel comes back with four, two or one warp.
bool stream_geometry(int Q, int D, int DP, int in_format, bool legacy_fits, int *warps, int *stages) {
    const size_t stage = stream_stage_bytes(D, in_format);
    const size_t budget = 227 * 1024;
    int forced_w = 0, forced_s = 0;
    if (const char *e = std::getenv("DDM_STREAM_WARPS")) forced_w = std::atoi(e);      // tuning knobs (bench only)
    if (const char *e = std::getenv("DDM_STREAM_STAGES")) forced_s = std::atoi(e);
    auto fits = [&](int w, int s_) { return stream_fixed_bytes(Q, DP, w, s_) + stage * w * s_ <= budget; };
    if (forced_w > 0 && forced_s > 0) {
        const int max_w = Q == 5 ? 16 : kStreamMaxWarps;
        if (forced_w > max_w || forced_s > kStreamMaxStages || forced_s < 1 || !fits(forced_w, forced_s)) return false;
        *warps = forced_w;
        *stages = forced_s;
        return true;
    }
    const size_t ring_target = 150 * 1024;
    // Q <= 5 (the NOAA configuration: 151 taps, D = 31 .. 37) has a 128-register build: sixteen warps with
    // ONE stage each -- the other fifteen warps cover a warp's reload -- measured as fast as eight warps
    // with two stages (2.17 ms) and without their slow outliers (p90 2.19 against 2.5 ms); 8-bit input, which
    // is bound by the unpacking arithmetic, gains 3.6 % from the sixteen warps (1.97 -> 1.90 ms per 1.84 G samples)
    if (Q == 5 && stage * 16 <= ring_target) {
        *warps = 16;
        *stages = (in_format == DDM_IN_CU8 && stage * 32 <= ring_target) ? 2 : 1;      // u8 stages are a quarter the size
        return true;
    }
    // everything else: the most warps (12, else 8) whose rings stay under the target, two stages each if
    // that still fits, else one.  Measured against the CTA-tiled kernel on the 1.84 G-sample pass
    // (profiles/r02_chain_ab.jsonl, r02_chain_dsweep.txt): D = 20 2.86 vs 3.28 ms, D = 24 3.08 vs 3.70,
    // D = 40 2.75 vs 3.16, D = 68 2.14 vs 2.39; a tie around D = 50 (2.19 vs 2.17), where the CTA-tiled
    // kernel's 128-block tiles amortise the per-tile work as well as eight warps hide it -- it keeps the
    // blocks of 12.5 .. 14 KB -- and the CTA-tiled kernel ahead from D = 74 on (D = 100: 2.07 vs 2.14).
    const bool tie_zone = in_format == DDM_IN_CF32 && stage > 12800 - 1 && stage < 14 * 1024 && legacy_fits;
    if (!tie_zone)
        for (int w : {12, 8}) {
            if (stage * w > ring_target) continue;
            int s_ = stage * w * 2 <= ring_target ? 2 : 1;
            if (in_format == DDM_IN_CU8)               // short stages: a deeper ring costs nothing
                while (s_ < 4 && stage * w * (s_ + 1) <= ring_target) ++s_;
            *warps = w;
            *stages = s_;
            return true;
        }
    if (legacy_fits) return false;
    for (int w : {4, 2, 1})
        if (fits(w, 2)) {
            *warps = w;
            *stages = 2;
            return true;
        }
    return false;
}

// Programmatic dependent launch is for chunk loops only: the second and later chunks of a stream, queued
// back to back by consecutive apply calls, of chunk-loop size.  A launch that follows a repositioning
// (time-sharded slabs, resets, batches) or a very long one gains nothing from it, and its CTAs, resident
// early and waiting, keep the SMs from kernels of other streams: measured on the time-sharded stream at
// N = 2, where the NCCL halo exchange of the next step then waits for a whole slab (9.91 against 8.73 ms,
// profiles/r02_timeshard_pdl_ab.jsonl).
constexpr long long kPdlMaxSamples = 1LL << 26;
inline bool use_pdl(const ddm_chain *c, const ChainParams &p) {
    return c->pdl && c->chained && p.batch == 1 && p.n <= kPdlMaxSamples;
}

template <int Q, bool MIX, int OUT, int IN>
int launch_stream_q(ddm_chain *c, const ChainParams &p0, cudaStream_t st) {
    ChainParams p = p0;
    const int W = c->st_warps, S = c->st_stages;
    const size_t smem = stream_fixed_bytes(Q, c->DP, W, S) + stream_stage_bytes(c->D, IN) * W * S;
    auto kern = chain_stream_kernel<Q, MIX, OUT, IN>;
    if constexpr (Q == 5) {
        if (W > kStreamMaxWarps) kern = chain_stream_kernel<Q, MIX, OUT, IN, 16>;      // 128-register build
    }
    if (!c->st_attr[p.s]) {   // first launch of this variant on this handle's device
        DDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        c->st_attr[p.s] = true;
    }
    p.stages = S;
    if (!c->step_set) {
        // the block rotator's advance per warp tile, exp(-j 2 pi r 32 D), from the double-double r
        const long double span = static_cast<long double>(kStreamTile) * c->D;
        long double t = static_cast<long double>(c->r_hi) * span;
        t -= std::floor(t);
        t += static_cast<long double>(c->r_lo) * span;
        const long double ang = 2.0L * 3.14159265358979323846264338327950288L * t;
        c->step_re = static_cast<double>(std::cos(ang));
        c->step_im = static_cast<double>(-std::sin(ang));
        c->step_set = true;
    }
    p.step_re = c->step_re;
    p.step_im = c->step_im;
    p.num_tiles = 1 + (p.M + kStreamTile - 1) / kStreamTile;        // per capture, warm-up tile 0 included
    if (p.num_tiles >= (1LL << 31) - 2) {
        set_error("chunk too long for one launch of the fused chain (%lld tiles)", static_cast<long long>(p.num_tiles));
        return DDM_ERR_UNSUPPORTED;
    }
    const long long total_tiles = p.num_tiles * p.batch;
    long long warps = static_cast<long long>(c->sms) * W;
    if (warps > total_tiles / kStreamMinTiles) warps = total_tiles / kStreamMinTiles;
    if (warps < 1) warps = 1;
    p.stream_warps = warps;
    const unsigned grid = static_cast<unsigned>((warps + W - 1) / W);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(32 * W);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl(c, p) ? 1 : 0;
    DDM_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    count_launch();
    return DDM_OK;
}

template <int Q, bool MIX, int OUT, int IN>
int launch_fused_q(ddm_chain *c, const ChainParams &p, cudaStream_t st) {
    if (c->stream) return launch_stream_q<Q, MIX, OUT, IN>(c, p, st);
    const size_t smem = chain_smem_bytes(Q, c->D, c->DP, IN);
    auto kern = chain_fused_kernel<Q, MIX, OUT, IN>;
    int &per_sm = c->per_sm[p.s];
    if (per_sm == 0) {   // first launch of this variant on this handle's device
        DDM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
        DDM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kChainThreads, smem));
        if (per_sm < 1) {
            per_sm = 0;
            set_error("fused chain kernel does not fit: %zu bytes of shared memory", smem);
            return DDM_ERR_UNSUPPORTED;
        }
    }
    long long grid = static_cast<long long>(c->sms) * per_sm;
    if (grid > p.num_tiles * p.batch) grid = p.num_tiles * p.batch;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kChainThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl(c, p) ? 1 : 0;
    DDM_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    count_launch();
    return DDM_OK;
}

template <bool MIX, int OUT, int IN>
int launch_fused_moi(ddm_chain *c, int Q, const ChainParams &p, cudaStream_t st) {
    switch (Q) {
        case 1: return launch_fused_q<1, MIX, OUT, IN>(c, p, st);
        case 2: return launch_fused_q<2, MIX, OUT, IN>(c, p, st);
        case 3: return launch_fused_q<3, MIX, OUT, IN>(c, p, st);
        case 4: return launch_fused_q<4, MIX, OUT, IN>(c, p, st);
        case 5: return launch_fused_q<5, MIX, OUT, IN>(c, p, st);
        case 6: return launch_fused_q<6, MIX, OUT, IN>(c, p, st);
        case 7: return launch_fused_q<7, MIX, OUT, IN>(c, p, st);
        case 8: return launch_fused_q<8, MIX, OUT, IN>(c, p, st);
        case 9: return launch_fused_q<9, MIX, OUT, IN>(c, p, st);
        case 10: return launch_fused_q<10, MIX, OUT, IN>(c, p, st);
    }
    set_error("internal: Q=%d out of range", Q);
    return DDM_ERR_UNSUPPORTED;
}

}  // namespace

// The kernel is instantiated for 10 values of Q x mixer x output mode x input format; the u8 half
// lives in its own translation unit (chain_u8.cu includes this file with DDM_CHAIN_PART = 1) so that
// the two halves compile in parallel.
int ddm_chain_launch_u8(ddm_chain *c, int mix, int out_mode, int Q, const ddm::ChainParams &p, cudaStream_t st);

#if defined(DDM_CHAIN_PART) && DDM_CHAIN_PART == 1
int ddm_chain_launch_u8(ddm_chain *c, int mix, int out_mode, int Q, const ddm::ChainParams &p, cudaStream_t st) {
    if (mix) {
        return out_mode == DDM_CHAIN_OUT_FM ? launch_fused_moi<true, DDM_CHAIN_OUT_FM, DDM_IN_CU8>(c, Q, p, st)
                                            : launch_fused_moi<true, DDM_CHAIN_OUT_IQ, DDM_IN_CU8>(c, Q, p, st);
    }
    return out_mode == DDM_CHAIN_OUT_FM ? launch_fused_moi<false, DDM_CHAIN_OUT_FM, DDM_IN_CU8>(c, Q, p, st)
                                        : launch_fused_moi<false, DDM_CHAIN_OUT_IQ, DDM_IN_CU8>(c, Q, p, st);
}
#else

namespace {

template <bool MIX, int OUT>
int launch_fused_mo(ddm_chain *c, int Q, const ChainParams &p, cudaStream_t st) {
    if (c->in_format == DDM_IN_CU8) return ddm_chain_launch_u8(c, MIX ? 1 : 0, OUT, Q, p, st);
    return launch_fused_moi<MIX, OUT, DDM_IN_CF32>(c, Q, p, st);
}

int launch_fused(ddm_chain *c, int Q, const ChainParams &p, cudaStream_t st) {
    if (c->mix) {
        return c->out_mode == DDM_CHAIN_OUT_FM ? launch_fused_mo<true, DDM_CHAIN_OUT_FM>(c, Q, p, st)
                                               : launch_fused_mo<true, DDM_CHAIN_OUT_IQ>(c, Q, p, st);
    }
    return c->out_mode == DDM_CHAIN_OUT_FM ? launch_fused_mo<false, DDM_CHAIN_OUT_FM>(c, Q, p, st)
                                           : launch_fused_mo<false, DDM_CHAIN_OUT_IQ>(c, Q, p, st);
}

int64_t positive_mod(int64_t a, int64_t b) {
    int64_t r = a % b;
    return r < 0 ? r + b : r;
}

int64_t positions_in(const ddm_chain *c, int64_t n) {
    if (n <= c->dec_off) return 0;
    return (n - c->dec_off + c->D - 1) / c->D;
}

// the reference's initial condition expressed as raw input history: zi = lfilter_zi(b,[1])
// (filters.py:45) is the state an all-ones *mixed* input leaves behind, so the raw halo is
// exp(+j 2 pi f g / fs) for g = -H..-1.
int build_initial_halo(ddm_chain *c) {
    std::vector<float2> h(c->H);
    for (int i = 0; i < c->H; ++i) {
        if (!c->mix) {
            h[i] = make_float2(1.f, 0.f);
            continue;
        }
        const double g = static_cast<double>(i - c->H);
        double t = c->r_hi * g;
        t -= std::rint(t);
        t += c->r_lo * g;
        h[i] = make_float2(static_cast<float>(std::cos(2.0 * M_PI * t)),
                           static_cast<float>(std::sin(2.0 * M_PI * t)));
    }
    DDM_CUDA(pool_alloc(&c->d_halo_init, sizeof(float2) * c->H));
    DDM_CUDA(cudaMemcpy(c->d_halo_init, h.data(), sizeof(float2) * c->H, cudaMemcpyHostToDevice));
    return DDM_OK;
}

int fill_initial_halo(ddm_chain *c, cudaStream_t st) {
    c->n_real = 0;
    c->chained = false;
    if (c->in_format == DDM_IN_CU8) {
        // u8 cannot encode the all-ones mixed history: it is kept virtual (n_real = 0) and the
        // first H samples of a fresh stream go through the general kernel, which knows about it
        DDM_CUDA(cudaMemsetAsync(c->d_halo[c->cur], 128, static_cast<size_t>(c->es) * c->H, st));
        return DDM_OK;
    }
    DDM_CUDA(cudaMemcpyAsync(c->d_halo[c->cur], c->d_halo_init, sizeof(float2) * c->H,
                             cudaMemcpyDeviceToDevice, st));
    return DDM_OK;
}

}  // namespace

extern "C" {

int ddm_chain_create(int device, const double *taps, int ntaps, int decim, double freq_offset,
                     double samp_rate, int out_mode, int in_format, ddm_chain **out) {
    DDM_REQUIRE(out != nullptr, "ddm_chain_create: out is NULL");
    *out = nullptr;
    DDM_REQUIRE(taps != nullptr && ntaps >= 1, "ddm_chain_create: need at least one tap");
    DDM_REQUIRE(decim >= 1, "ddm_chain_create: decimation must be >= 1 (got %d)", decim);
    DDM_REQUIRE(samp_rate > 0, "ddm_chain_create: sampling rate must be positive");
    DDM_REQUIRE(out_mode == DDM_CHAIN_OUT_FM || out_mode == DDM_CHAIN_OUT_IQ,
                "ddm_chain_create: bad out_mode %d", out_mode);
    DDM_REQUIRE(in_format == DDM_IN_CF32 || in_format == DDM_IN_CU8, "ddm_chain_create: bad in_format %d",
                in_format);
    int ndev = 0;
    DDM_CUDA(cudaGetDeviceCount(&ndev));
    DDM_REQUIRE(device >= 0 && device < ndev, "ddm_chain_create: no such device %d", device);
    DeviceGuard guard(device);

    ddm_chain *c = new (std::nothrow) ddm_chain();
    if (!c) {
        set_error("ddm_chain_create: out of host memory");
        return DDM_ERR_NOMEM;
    }
    c->device = device;
    c->K = ntaps;
    c->D = decim;
    c->freq = freq_offset;
    c->fs = samp_rate;
    c->out_mode = out_mode;
    c->in_format = in_format;
    c->mix = freq_offset != 0.0;
    c->taps.assign(taps, taps + ntaps);
    c->sms = sm_count(device);
    // r = f/fs turns per sample as a double-double
    c->r_hi = freq_offset / samp_rate;
    c->r_lo = std::fma(-c->r_hi, samp_rate, freq_offset) / samp_rate;

    const int D = decim, K = ntaps;
    const int qmax = (K + 1 + D - 1) / D;
    c->DP = (D + 3) & ~3;
    c->es = in_format == DDM_IN_CU8 ? 2 : 8;
    c->halo_in_kernel = std::getenv("DDM_CHAIN_HALO_MEMCPY") == nullptr;
    c->pdl = std::getenv("DDM_CHAIN_NO_PDL") == nullptr;
    const bool legacy = std::getenv("DDM_CHAIN_LEGACY") != nullptr;         // A/B against the CTA-tiled kernel
    const bool legacy_fits = D >= 2 && qmax <= kChainMaxQ && chain_smem_bytes(qmax, D, c->DP, in_format) <= 227 * 1024;
    c->stream = !legacy && D >= 2 && qmax <= kChainMaxQ &&
                stream_geometry(qmax, D, c->DP, in_format, legacy_fits, &c->st_warps, &c->st_stages);
    c->fast = c->stream || legacy_fits;
    // halo: the Q (odd D: up to Q + 1) leading blocks of a tile plus one block of slack
    c->H = (qmax + 1 + (D & 1)) * D;
    if (c->H & 1) c->H += 1;
    if (in_format == DDM_IN_CU8) c->H = (c->H + 7) / 8 * 8 + 8;   // aligned tile copies may start 7 samples early

    auto fail = [&](int code) {
        ddm_chain_destroy(c);
        return code;
    };
    cudaError_t e;
    for (int i = 0; i < 2; ++i) {
        e = pool_alloc(&c->d_halo[i], static_cast<size_t>(c->es) * c->H);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(halo) failed: %s", cudaGetErrorString(e));
            return fail(DDM_ERR_NOMEM);
        }
    }
    {
        e = pool_alloc_t(&c->d_taps_lin, sizeof(double) * K);
        if (e == cudaSuccess)
            e = cudaMemcpy(c->d_taps_lin, taps, sizeof(double) * K, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("tap upload failed: %s", cudaGetErrorString(e));
            return fail(DDM_ERR_CUDA);
        }
    }
    if (c->mix) {
        // complex taps of the general path (chain_generic_y_kernel), phases in long double
        std::vector<double2> ct(K);
        const long double r = static_cast<long double>(freq_offset) / static_cast<long double>(samp_rate);
        for (int k = 0; k < K; ++k) {
            long double ph = r * k;
            ph -= std::floor(ph);
            const long double ang = 2.0L * 3.14159265358979323846264338327950288L * ph;
            ct[k] = make_double2(static_cast<double>(taps[k] * std::cos(ang)), static_cast<double>(taps[k] * std::sin(ang)));
        }
        e = pool_alloc_t(&c->d_ctaps, sizeof(double2) * K);
        if (e == cudaSuccess) e = cudaMemcpy(c->d_ctaps, ct.data(), sizeof(double2) * K, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("tap upload failed: %s", cudaGetErrorString(e));
            return fail(DDM_ERR_CUDA);
        }
    }
    if (c->fast) {
        for (int s = 0; s < 2; ++s) {
            const int Q = (K + s + D - 1) / D;
            c->Q[s] = Q;
            std::vector<float> t(static_cast<size_t>(Q) * c->DP, 0.f);
            for (int q = 0; q < Q; ++q)
                for (int a = 0; a < D; ++a) {
                    const int k = q * D + D - 1 - a - s;
                    if (k >= 0 && k < K) t[static_cast<size_t>(q) * c->DP + a] = static_cast<float>(taps[k]);
                }
            // first tap position (rounded down to the 4-sample body) the last partial sum needs
            int first = D;
            for (int a = 0; a < D; ++a)
                if (t[static_cast<size_t>(Q - 1) * c->DP + a] != 0.f) {
                    first = a;
                    break;
                }
            c->a_lastq[s] = Q > 1 ? std::min(first & ~3, D & ~3) : 0;
            e = pool_alloc_t(&c->d_taps[s], sizeof(float) * t.size());
            if (e == cudaSuccess)
                e = cudaMemcpy(c->d_taps[s], t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) {
                set_error("tap table upload failed: %s", cudaGetErrorString(e));
                return fail(DDM_ERR_CUDA);
            }
        }
        std::vector<float2> rot(c->DP, make_float2(1.f, 0.f));
        for (int a = 0; a < D; ++a) {
            double t = c->r_hi * a;
            t -= std::rint(t);
            t += c->r_lo * a;
            rot[a] = make_float2(static_cast<float>(std::cos(2.0 * M_PI * t)),
                                 static_cast<float>(-std::sin(2.0 * M_PI * t)));
        }
        e = pool_alloc_t(&c->d_rot, sizeof(float2) * c->DP);
        if (e == cudaSuccess)
            e = cudaMemcpy(c->d_rot, rot.data(), sizeof(float2) * c->DP, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("rotator table upload failed: %s", cudaGetErrorString(e));
            return fail(DDM_ERR_CUDA);
        }
    }
    int rc = build_initial_halo(c);
    if (rc == DDM_OK) rc = fill_initial_halo(c, nullptr);
    if (rc != DDM_OK) return fail(rc);
    DDM_CUDA(cudaStreamSynchronize(nullptr));
    *out = c;
    return DDM_OK;
}

int ddm_chain_destroy(ddm_chain *c) {
    if (!c) return DDM_OK;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();          // pooled blocks are recycled at once: nothing may still read them
    for (int i = 0; i < 2; ++i) {
        pool_free(c->d_halo[i]);
        pool_free(c->d_taps[i]);
    }
    pool_free(c->d_halo_init);
    pool_free(c->d_taps_lin);
    pool_free(c->d_ctaps);
    pool_free(c->d_rot);
    cudaFree(c->d_ytmp);
    cudaFree(c->d_in);
    cudaFree(c->d_out);
    for (cudaEvent_t e : c->ev_h2d)
        if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
    return DDM_OK;
}

int ddm_chain_reset(ddm_chain *c) {
    DDM_REQUIRE(c != nullptr, "ddm_chain_reset: NULL handle");
    DeviceGuard guard(c->device);
    c->n0 = 0;
    c->dec_off = 0;
    c->has_prev = 0;
    return fill_initial_halo(c, nullptr);
}

int ddm_chain_halo_len(const ddm_chain *c, int64_t *n) {
    DDM_REQUIRE(c != nullptr && n != nullptr, "ddm_chain_halo_len: NULL argument");
    *n = c->H;
    return DDM_OK;
}

int ddm_chain_out_count(const ddm_chain *c, int64_t n, int64_t *n_out) {
    DDM_REQUIRE(c != nullptr && n_out != nullptr, "ddm_chain_out_count: NULL argument");
    DDM_REQUIRE(n >= 0, "ddm_chain_out_count: negative length");
    const int64_t M = positions_in(c, n);
    if (c->out_mode == DDM_CHAIN_OUT_IQ) *n_out = M;
    else *n_out = c->has_prev ? M : (M > 0 ? M - 1 : 0);
    return DDM_OK;
}

int ddm_chain_get_position(const ddm_chain *c, int64_t *n0, int64_t *dec_off, int *has_prev) {
    DDM_REQUIRE(c != nullptr, "ddm_chain_get_position: NULL handle");
    if (n0) *n0 = c->n0;
    if (dec_off) *dec_off = c->dec_off;
    if (has_prev) *has_prev = c->has_prev;
    return DDM_OK;
}

int ddm_chain_set_position(ddm_chain *c, int64_t n0, int64_t dec_off, int has_prev,
                           const void *halo_dev, void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_chain_set_position: NULL handle");
    DDM_REQUIRE(dec_off >= 0, "ddm_chain_set_position: negative decimation offset");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    c->n0 = n0;
    c->dec_off = dec_off;
    c->has_prev = has_prev ? 1 : 0;
    if (halo_dev == nullptr) return fill_initial_halo(c, st);
    DDM_CUDA(cudaMemcpyAsync(c->d_halo[c->cur], halo_dev, static_cast<size_t>(c->es) * c->H,
                             cudaMemcpyDeviceToDevice, st));
    c->n_real = c->H;
    c->chained = false;
    return DDM_OK;
}

int ddm_chain_get_halo(const ddm_chain *c, void *halo_dev, void *stream) {
    DDM_REQUIRE(c != nullptr && halo_dev != nullptr, "ddm_chain_get_halo: NULL argument");
    DeviceGuard guard(c->device);
    DDM_CUDA(cudaMemcpyAsync(halo_dev, c->d_halo[c->cur], static_cast<size_t>(c->es) * c->H,
                             cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return DDM_OK;
}

int ddm_chain_export_state(const ddm_chain *c, double *zi_c128_host, double *last_c128_host,
                           void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_chain_export_state: NULL handle");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<unsigned char> raw(static_cast<size_t>(c->es) * c->H);
    DDM_CUDA(cudaMemcpyAsync(raw.data(), c->d_halo[c->cur], raw.size(), cudaMemcpyDeviceToHost, st));
    DDM_CUDA(cudaStreamSynchronize(st));
    // mixed history x'[g], g = n0-H .. n0-1, rounded to complex64 like the reference's in-place mixer
    std::vector<double> xr(c->H), xi(c->H);
    const bool u8 = c->in_format == DDM_IN_CU8;
    for (int i = 0; i < c->H; ++i) {
        if (u8 && i < c->H - c->n_real) {            // virtual history: the all-ones mixed signal
            xr[i] = 1.0;
            xi[i] = 0.0;
            continue;
        }
        double re, im;
        if (u8) {
            re = static_cast<double>(raw[2 * i]) - 127.5;
            im = static_cast<double>(raw[2 * i + 1]) - 127.5;
        } else {
            re = reinterpret_cast<const float2 *>(raw.data())[i].x;
            im = reinterpret_cast<const float2 *>(raw.data())[i].y;
        }
        if (c->mix) {
            const double g = static_cast<double>(c->n0 - c->H + i);
            const double p = c->r_hi * g;
            const double e = std::fma(c->r_hi, g, -p);
            double t = p - std::rint(p);
            t += e + c->r_lo * g;
            const double cs = std::cos(2.0 * M_PI * t), sn = -std::sin(2.0 * M_PI * t);
            const double mr = re * cs - im * sn, mi = re * sn + im * cs;
            re = static_cast<double>(static_cast<float>(mr));
            im = static_cast<double>(static_cast<float>(mi));
        }
        xr[i] = re;
        xi[i] = im;
    }
    const int K = c->K;
    if (zi_c128_host) {
        // zi[i] = sum_{k>i} b[k] x'[n0 + i - k]
        for (int i = 0; i < K - 1; ++i) {
            double ar = 0, ai = 0;
            for (int k = i + 1; k < K; ++k) {
                const int j = c->H + i - k;
                if (j < 0) break;
                ar += c->taps[k] * xr[j];
                ai += c->taps[k] * xi[j];
            }
            zi_c128_host[2 * i] = ar;
            zi_c128_host[2 * i + 1] = ai;
        }
    }
    if (last_c128_host) {
        // last decimated sample y[m_last], at chunk-relative position dec_off - D
        last_c128_host[0] = last_c128_host[1] = 0.0;
        if (c->has_prev) {
            const long long pos = static_cast<long long>(c->H) + c->dec_off - c->D;
            double ar = 0, ai = 0;
            for (int k = 0; k < K; ++k) {
                const long long j = pos - k;
                if (j < 0) break;
                ar += c->taps[k] * xr[j];
                ai += c->taps[k] * xi[j];
            }
            last_c128_host[0] = ar;
            last_c128_host[1] = ai;
        }
    }
    return DDM_OK;
}

namespace {
int chain_apply_piece(ddm_chain *c, const void *x_dev, int64_t n, void *out_dev, int64_t *n_out, cudaStream_t st);
}

int ddm_chain_apply_dev(ddm_chain *c, const void *x_dev, int64_t n, void *out_dev,
                        int64_t out_capacity, int64_t *n_out, void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_chain_apply_dev: NULL handle");
    DDM_REQUIRE(n >= 0, "ddm_chain_apply_dev: negative length");
    DDM_REQUIRE(n == 0 || x_dev != nullptr, "ddm_chain_apply_dev: NULL input");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t produced = 0;
    ddm_chain_out_count(c, n, &produced);
    if (n_out) *n_out = produced;
    if (produced > out_capacity) {
        set_error("ddm_chain_apply_dev: output needs %lld samples, capacity is %lld",
                  static_cast<long long>(produced), static_cast<long long>(out_capacity));
        return DDM_ERR_CAPACITY;
    }
    DDM_REQUIRE(produced == 0 || out_dev != nullptr, "ddm_chain_apply_dev: NULL output");
    // A fresh u8 stream has virtual history the fused kernel cannot stage: its first H samples go
    // through the general kernel (as one piece), the rest through the fused kernel.
    const unsigned char *xb = static_cast<const unsigned char *>(x_dev);
    unsigned char *ob = static_cast<unsigned char *>(out_dev);
    const size_t oes = c->out_mode == DDM_CHAIN_OUT_IQ ? sizeof(float2) : sizeof(float);
    int64_t done = 0;
    while (done < n || (n == 0 && done == 0)) {
        int64_t piece = n - done;
        if (c->in_format == DDM_IN_CU8 && c->fast && c->n_real < c->H && piece > c->H - c->n_real)
            piece = c->H - c->n_real;
        int64_t got = 0;
        int rc = chain_apply_piece(c, xb + static_cast<size_t>(done) * c->es, piece, ob, &got, st);
        if (rc != DDM_OK) return rc;
        ob += static_cast<size_t>(got) * oes;
        done += piece;
        if (n == 0) break;
    }
    return DDM_OK;
}

namespace {
int chain_apply_piece(ddm_chain *c, const void *x_dev, int64_t n, void *out_dev, int64_t *n_out, cudaStream_t st) {
    const int64_t M = positions_in(c, n);
    int64_t produced = 0;
    ddm_chain_out_count(c, n, &produced);
    *n_out = produced;
    const unsigned char *x = static_cast<const unsigned char *>(x_dev);
    const int D = c->D;
    const size_t es = static_cast<size_t>(c->es);
    bool halo_saved = false;

    if (M > 0) {
        const bool aligned = (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0;
        const bool history_ok = c->in_format != DDM_IN_CU8 || c->n_real >= c->H;
        if (c->fast && aligned && history_ok) {
            const int s = static_cast<int>((c->dec_off + 1 + D) & 1);     // makes block 0 start on an even sample
            const int Q = c->Q[s];
            ChainParams p{};
            p.x = x_dev;
            p.halo = c->d_halo[c->cur];
            p.out = out_dev;
            p.taps = c->d_taps[s];
            p.rot = c->d_rot;
            p.n = n;
            p.n0 = c->n0;
            p.M = M;
            p.b0 = c->dec_off + s + 1 - D;
            const int J = (D & 1) ? ((kChainThreads - Q) & ~1) : (kChainThreads - Q);
            p.J = J;
            p.num_tiles = (M + J - 1) / J;
            p.r_hi = c->r_hi;
            p.r_lo = c->r_lo;
            p.D = D;
            p.DP = c->DP;
            p.H = c->H;
            p.s = s;
            p.has_prev = c->has_prev;
            p.a_lastq = c->a_lastq[s];
            p.batch = 1;
            p.x_stride = p.out_stride = 0;
            if (n >= c->H && c->halo_in_kernel) p.halo_out = c->d_halo[c->cur ^ 1];
            int rc = launch_fused(c, Q, p, st);
            if (rc != DDM_OK) return rc;
            if (p.halo_out != nullptr) {
                c->cur ^= 1;
                halo_saved = true;
            }
        } else {
            const size_t need = static_cast<size_t>(M + 1);
            if (need > c->ytmp_cap) {
                DDM_CUDA(cudaStreamSynchronize(st));
                cudaFree(c->d_ytmp);
                c->d_ytmp = nullptr;
                c->ytmp_cap = 0;
                DDM_CUDA(cudaMalloc(&c->d_ytmp, sizeof(double2) * need));
                c->ytmp_cap = need;
            }
            GenericParams g{};
            g.x = x_dev;
            g.halo = c->d_halo[c->cur];
            g.y = c->d_ytmp;
            g.taps = c->d_taps_lin;
            g.ctaps = c->d_ctaps;
            g.n = n;
            g.n0 = c->n0;
            g.M = M;
            g.off = c->dec_off;
            g.r_hi = c->r_hi;
            g.r_lo = c->r_lo;
            g.K = c->K;
            g.D = D;
            g.H = c->H;
            g.mix = c->mix ? 1 : 0;
            g.in_format = c->in_format;
            g.virt_before = c->in_format == DDM_IN_CU8 ? -c->n_real : -static_cast<long long>(c->H) - 1;
            const int tb = 128;
            const int row_len = kGenTile + (c->K - 1) / D + 1;
            const size_t tile_bytes = sizeof(double2) * static_cast<size_t>(D) * row_len;
            const bool has_virtual = c->in_format == DDM_IN_CU8 && c->n_real < c->H;
            if (!has_virtual && D == 1 && c->K <= 4096) {
                const int rl = (kGenTile * kGen1Outs - 1 + c->K + 3) / 4 + 1;
                const size_t bytes = sizeof(double2) * 4 * static_cast<size_t>(rl);
                DDM_CUDA(cudaFuncSetAttribute(chain_generic_tile1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(bytes)));
                const int per_cta = kGenTile * kGen1Outs;
                chain_generic_tile1_kernel<<<static_cast<unsigned>((M + 1 + per_cta - 1) / per_cta), kGenTile, bytes, st>>>(
                    g, rl);
            } else if (!has_virtual && tile_bytes <= 200 * 1024) {
                DDM_CUDA(cudaFuncSetAttribute(chain_generic_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(tile_bytes)));
                chain_generic_tile_kernel<<<static_cast<unsigned>((M + 1 + kGenTile - 1) / kGenTile), kGenTile,
                                            tile_bytes, st>>>(g, row_len);
            } else {
                chain_generic_y_kernel<<<static_cast<unsigned>((M + 1 + tb - 1) / tb), tb, 0, st>>>(g);
            }
            DDM_CUDA(cudaGetLastError());
            chain_generic_out_kernel<<<static_cast<unsigned>((M + tb - 1) / tb), tb, 0, st>>>(
                c->d_ytmp, out_dev, M, c->has_prev, c->out_mode);
            DDM_CUDA(cudaGetLastError());
            count_launch(2);
        }
    }

    // ---- carry: halo <- last H samples of (halo ++ x) ----
    if (halo_saved) {
        // the fused kernel wrote it (save_halo)
    } else if (n >= c->H) {
        DDM_CUDA(cudaMemcpyAsync(c->d_halo[c->cur], x + (n - c->H) * es, es * c->H, cudaMemcpyDeviceToDevice, st));
    } else if (n > 0) {
        const int nxt = c->cur ^ 1;
        const unsigned char *old = static_cast<const unsigned char *>(c->d_halo[c->cur]);
        unsigned char *neu = static_cast<unsigned char *>(c->d_halo[nxt]);
        DDM_CUDA(cudaMemcpyAsync(neu, old + n * es, es * (c->H - n), cudaMemcpyDeviceToDevice, st));
        DDM_CUDA(cudaMemcpyAsync(neu + (c->H - n) * es, x, es * n, cudaMemcpyDeviceToDevice, st));
        c->cur = nxt;
    }
    c->n_real = std::min<long long>(c->H, c->n_real + n);
    // comm.py:124  nextOff = (j - (len - off) % j) % j   (python modulo)
    c->dec_off = positive_mod(D - positive_mod(n - c->dec_off, D), D);
    c->n0 += n;
    if (M > 0) c->has_prev = 1;
    c->chained = true;
    return DDM_OK;
}
}  // namespace

int ddm_chain_apply_batch_dev(ddm_chain *c, const void *x_dev, int64_t n, int64_t batch, int64_t x_stride,
                              void *out_dev, int64_t out_stride, int64_t *n_out, void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_chain_apply_batch_dev: NULL handle");
    DDM_REQUIRE(n >= 0 && batch >= 0 && x_stride >= n, "ddm_chain_apply_batch_dev: bad sizes");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // every capture is a fresh stream: reset the handle (it is left in that state)
    c->n0 = 0;
    c->dec_off = 0;
    c->has_prev = 0;
    int rc = fill_initial_halo(c, st);
    if (rc != DDM_OK) return rc;
    int64_t per = 0;
    ddm_chain_out_count(c, n, &per);
    if (n_out) *n_out = per;
    DDM_REQUIRE(out_stride >= per, "ddm_chain_apply_batch_dev: out_stride %lld is below the %lld outputs per capture",
                static_cast<long long>(out_stride), static_cast<long long>(per));
    if (batch == 0 || n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_chain_apply_batch_dev: NULL buffer");
    const int64_t M = positions_in(c, n);
    const bool aligned = (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0 && (x_stride * c->es) % 16 == 0;
    if (c->fast && aligned && c->in_format == DDM_IN_CF32 && M > 0) {
        const int s = (1 + c->D) & 1;                     // dec_off = 0: block 0 starts on an even sample
        const int Q = c->Q[s];
        ChainParams p{};
        p.x = x_dev;
        p.halo = c->d_halo[c->cur];
        p.out = out_dev;
        p.taps = c->d_taps[s];
        p.rot = c->d_rot;
        p.n = n;
        p.n0 = 0;
        p.M = M;
        p.b0 = 0 + s + 1 - c->D;
        const int J = (c->D & 1) ? ((kChainThreads - Q) & ~1) : (kChainThreads - Q);
        p.J = J;
        p.num_tiles = (M + J - 1) / J;
        p.r_hi = c->r_hi;
        p.r_lo = c->r_lo;
        p.D = c->D;
        p.DP = c->DP;
        p.H = c->H;
        p.s = s;
        p.has_prev = 0;
        p.a_lastq = c->a_lastq[s];
        p.batch = batch;
        p.x_stride = x_stride;
        p.out_stride = out_stride;
        return launch_fused(c, Q, p, st);
    }
    // general configurations: capture by capture
    const size_t oes = c->out_mode == DDM_CHAIN_OUT_IQ ? sizeof(float2) : sizeof(float);
    for (int64_t k = 0; k < batch; ++k) {
        c->n0 = 0;
        c->dec_off = 0;
        c->has_prev = 0;
        rc = fill_initial_halo(c, st);
        if (rc != DDM_OK) return rc;
        int64_t got = 0;
        rc = ddm_chain_apply_dev(c, static_cast<const unsigned char *>(x_dev) + static_cast<size_t>(k) * x_stride * c->es, n,
                                 static_cast<unsigned char *>(out_dev) + static_cast<size_t>(k) * out_stride * oes,
                                 out_stride, &got, stream);
        if (rc != DDM_OK) return rc;
    }
    c->n0 = 0;
    c->dec_off = 0;
    c->has_prev = 0;
    return fill_initial_halo(c, st);
}

int ddm_chain_apply_host(ddm_chain *c, const void *x_host, int64_t n, void *out_host,
                         int64_t out_capacity, int64_t *n_out, void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_chain_apply_host: NULL handle");
    DDM_REQUIRE(n >= 0, "ddm_chain_apply_host: negative length");
    DDM_REQUIRE(n == 0 || x_host != nullptr, "ddm_chain_apply_host: NULL input");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t produced = 0;
    ddm_chain_out_count(c, n, &produced);
    if (produced > out_capacity) {
        if (n_out) *n_out = produced;
        set_error("ddm_chain_apply_host: output needs %lld samples, capacity is %lld",
                  static_cast<long long>(produced), static_cast<long long>(out_capacity));
        return DDM_ERR_CAPACITY;
    }
    const size_t in_bytes = static_cast<size_t>(c->es) * static_cast<size_t>(n);
    const size_t out_elem = c->out_mode == DDM_CHAIN_OUT_IQ ? sizeof(float2) : sizeof(float);
    const size_t out_bytes = out_elem * static_cast<size_t>(produced);
    if (in_bytes > c->in_cap) {
        DDM_CUDA(cudaStreamSynchronize(st));
        cudaFree(c->d_in);
        c->d_in = nullptr;
        c->in_cap = 0;
        DDM_CUDA(cudaMalloc(&c->d_in, in_bytes));
        c->in_cap = in_bytes;
    }
    if (out_bytes > c->out_cap) {
        DDM_CUDA(cudaStreamSynchronize(st));
        cudaFree(c->d_out);
        c->d_out = nullptr;
        c->out_cap = 0;
        DDM_CUDA(cudaMalloc(&c->d_out, out_bytes));
        c->out_cap = out_bytes;
    }
    // Long chunks go through in four pieces: all host->device copies are queued on a second stream at
    // once, the kernel of a piece waits for its copy only, and its results start back to the host while
    // the next piece's input is still arriving (the two directions use different copy engines).  The
    // serial tail behind the input transfer shrinks from a whole chunk's kernel + result copy to a
    // quarter of it: +2 % on cf32 input, +7 % on 8-bit input, whose transfer is four times shorter.
    // Piecewise application is the chunk invariance the chain guarantees anyway.
    constexpr int kPieces = 4;
    const int pieces = n >= (1 << 22) ? kPieces : 1;
    if (pieces > 1 && c->copy_stream == nullptr) {
        DDM_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < kPieces; ++i) DDM_CUDA(cudaEventCreateWithFlags(&c->ev_h2d[i], cudaEventDisableTiming));
    }
    const unsigned char *src = static_cast<const unsigned char *>(x_host);
    unsigned char *dst = static_cast<unsigned char *>(out_host);
    const int64_t piece_len = pieces > 1 ? ((n + pieces - 1) / pieces + 63) / 64 * 64 : n;
    if (pieces > 1) {
        for (int i = 0; i < pieces; ++i) {
            const int64_t a = std::min<int64_t>(n, i * piece_len), b = std::min<int64_t>(n, a + piece_len);
            if (b > a)
                DDM_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(c->d_in) + a * c->es, src + a * c->es,
                                         static_cast<size_t>(b - a) * c->es, cudaMemcpyHostToDevice, c->copy_stream));
            DDM_CUDA(cudaEventRecord(c->ev_h2d[i], c->copy_stream));
        }
    } else if (n > 0) {
        DDM_CUDA(cudaMemcpyAsync(c->d_in, x_host, in_bytes, cudaMemcpyHostToDevice, st));
    }
    int64_t total = 0;
    for (int i = 0; i < pieces; ++i) {
        const int64_t a = std::min<int64_t>(n, i * piece_len), b = std::min<int64_t>(n, a + piece_len);
        if (pieces > 1) DDM_CUDA(cudaStreamWaitEvent(st, c->ev_h2d[i], 0));
        if (b <= a && !(pieces == 1)) continue;
        int64_t got = 0;
        int rc = ddm_chain_apply_dev(c, static_cast<unsigned char *>(c->d_in) + a * c->es, b - a,
                                     static_cast<unsigned char *>(c->d_out) + total * out_elem, produced - total, &got,
                                     stream);
        if (rc != DDM_OK) {
            cudaStreamSynchronize(st);
            if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
            return rc;
        }
        if (got > 0)
            DDM_CUDA(cudaMemcpyAsync(dst + total * out_elem, static_cast<unsigned char *>(c->d_out) + total * out_elem,
                                     static_cast<size_t>(got) * out_elem, cudaMemcpyDeviceToHost, st));
        total += got;
    }
    if (n_out) *n_out = total;
    DDM_CUDA(cudaStreamSynchronize(st));
    return DDM_OK;
}

}  // extern "C"

#endif  // DDM_CHAIN_PART
