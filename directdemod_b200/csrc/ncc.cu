// Sync-pattern correlation and peak picking: decode_noaa.__correlate /
// __correlateAndFindPeaks (decode_noaa.py:659-767), and the AFSK mark/space correlator bank
// (decode_afsk1200.py:106-142).
//
//   cor[i]  = sum_k h[i - M/2 + k] needle[k]                 signal.correlate(h, needle, 'same')
//   sums[i] = sum_k h[i - M/2 + k]^2                         np.convolve(h*h, ones(M), 'same')
//   ncc[i]  = cor[i] / sqrt(sums[i] * sum(needle^2))
//
// The APT sync needles are piecewise constant (40 bits x round(fs/4160) repeats), so cor is a
// short weighted sum of range sums: two-level float64 prefix sums of h and h^2 (exclusive
// prefix inside blocks of 2048 samples + block totals, so a range sum never subtracts two large
// numbers) turn the 560- or 19 680-tap correlation into ~30 range queries per output --
// HBM-bound instead of FP64-bound.  Arbitrary needles use the direct kernel.
//
// Peak picking: the scalar threshold of decode_noaa.py:714-723 needs the K largest and K
// smallest values of 54 M correlations; a float64 radix select (8 histogram passes) finds them
// on the device, an ordered compaction extracts the candidates above the threshold, and the
// sequential group-maximum scan of :731-746 runs over that short list on the host, bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ddm_common.cuh"

namespace ddm {

constexpr int kPfxBlock = 2048;      // samples per prefix block
constexpr int kPfxThreads = 256;     // 8 samples per thread
constexpr int kMaxRuns = 128;

struct NeedleRuns {
    int count;
    int start[kMaxRuns];
    int end[kMaxRuns];
    double value[kMaxRuns];
};

template <typename T>
__device__ __forceinline__ double to_double(T v) { return static_cast<double>(v); }

// exclusive prefix inside each block of kPfxBlock samples, for h and h^2, plus block totals.
// L has n + 1 entries (entry n lets a range end at the very end of the array).
template <typename T>
__global__ void __launch_bounds__(kPfxThreads)
block_prefix_kernel(const T *__restrict__ h, long long n, double *__restrict__ L1, double *__restrict__ L2,
                    double *__restrict__ T1, double *__restrict__ T2) {
    __shared__ double s1[kPfxThreads], s2[kPfxThreads];
    const long long base = static_cast<long long>(blockIdx.x) * kPfxBlock;
    const int tid = threadIdx.x;
    constexpr int PER = kPfxBlock / kPfxThreads;
    double v[PER];
    double a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const long long j = base + tid * PER + i;
        v[i] = j < n ? to_double(h[j]) : 0.0;
        a1 += v[i];
        a2 = fma(v[i], v[i], a2);
    }
    s1[tid] = a1;
    s2[tid] = a2;
    __syncthreads();
    // Hillis-Steele over the 256 thread sums
    for (int off = 1; off < kPfxThreads; off <<= 1) {
        double b1 = 0.0, b2 = 0.0;
        if (tid >= off) {
            b1 = s1[tid - off];
            b2 = s2[tid - off];
        }
        __syncthreads();
        s1[tid] += b1;
        s2[tid] += b2;
        __syncthreads();
    }
    double p1 = tid > 0 ? s1[tid - 1] : 0.0;
    double p2 = tid > 0 ? s2[tid - 1] : 0.0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const long long j = base + tid * PER + i;
        if (j <= n) {
            L1[j] = p1;
            L2[j] = p2;
        }
        p1 += v[i];
        p2 = fma(v[i], v[i], p2);
    }
    if (tid == kPfxThreads - 1) {
        T1[blockIdx.x] = s1[tid];
        T2[blockIdx.x] = s2[tid];
    }
}

// sum over [a, b), 0 <= a <= b <= n
__device__ __forceinline__ double range_sum(const double *__restrict__ L, const double *__restrict__ T,
                                            long long a, long long b) {
    const long long ba = a / kPfxBlock, bb = b / kPfxBlock;
    if (ba == bb) return L[b] - L[a];
    double s = T[ba] - L[a];
    for (long long k = ba + 1; k < bb; ++k) s += T[k];
    return s + L[b];
}

__global__ void ncc_runs_kernel(const double *__restrict__ L1, const double *__restrict__ L2,
                                const double *__restrict__ T1, const double *__restrict__ T2, long long n,
                                int M, double needle_energy, int normalised, const NeedleRuns runs,
                                double *__restrict__ out) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long c = M / 2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const long long w0 = i - c;
        double cor = 0.0;
        for (int r = 0; r < runs.count; ++r) {
            long long a = w0 + runs.start[r], b = w0 + runs.end[r];
            a = a < 0 ? 0 : (a > n ? n : a);
            b = b < 0 ? 0 : (b > n ? n : b);
            if (b > a) cor = fma(runs.value[r], range_sum(L1, T1, a, b), cor);
        }
        if (normalised) {
            long long a = w0, b = w0 + M;
            a = a < 0 ? 0 : a;
            b = b > n ? n : b;
            const double sums = b > a ? range_sum(L2, T2, a, b) : 0.0;
            out[i] = cor / sqrt(sums * needle_energy);
        } else {
            out[i] = cor;
        }
    }
}

// Short needles (M <= kNccTileMaxM): everything in one kernel, no global prefix arrays.  A CTA
// stages the window its `tile` outputs look at, turns it into exclusive prefix sums of h and h^2
// in shared memory (window-local, so a range sum is ONE subtraction of small numbers), and
// evaluates the run-length correlation from shared memory.  HBM traffic: ~(1 + M/tile) x 4 B read
// + 8 B written per output.
constexpr int kNccTileThreads = 256;
constexpr int kNccTile = 2048;
constexpr int kNccTileMaxM = 2048;

template <typename T>
__global__ void __launch_bounds__(kNccTileThreads)
ncc_tile_kernel(const T *__restrict__ h, long long n, int M, double needle_energy, int normalised,
                const NeedleRuns runs, double *__restrict__ out) {
    extern __shared__ double ncc_sm[];
    const int win = kNccTile + M;
    double *P1 = ncc_sm;                       // win + 1
    double *P2 = P1 + (win + 1);               // win + 1
    double *S1 = P2 + (win + 1);               // kNccTileThreads
    double *S2 = S1 + kNccTileThreads;         // kNccTileThreads
    const int tid = threadIdx.x;
    const long long i0 = static_cast<long long>(blockIdx.x) * kNccTile;
    const long long base = i0 - M / 2;         // global index of window position 0
    for (int j = tid; j < win; j += kNccTileThreads) {
        const long long g = base + j;
        P1[j] = (g >= 0 && g < n) ? to_double(h[g]) : 0.0;
    }
    __syncthreads();
    const int per = (win + kNccTileThreads - 1) / kNccTileThreads;
    const int j0 = tid * per;
    const int j1 = min(j0 + per, win);
    double a1 = 0.0, a2 = 0.0;
    for (int j = j0; j < j1; ++j) {
        const double v = P1[j];
        a1 += v;
        a2 = fma(v, v, a2);
    }
    S1[tid] = a1;
    S2[tid] = a2;
    __syncthreads();
    for (int off = 1; off < kNccTileThreads; off <<= 1) {
        double b1 = 0.0, b2 = 0.0;
        if (tid >= off) {
            b1 = S1[tid - off];
            b2 = S2[tid - off];
        }
        __syncthreads();
        S1[tid] += b1;
        S2[tid] += b2;
        __syncthreads();
    }
    double p1 = tid > 0 ? S1[tid - 1] : 0.0;
    double p2 = tid > 0 ? S2[tid - 1] : 0.0;
    for (int j = j0; j < j1; ++j) {
        const double v = P1[j];
        P1[j] = p1;
        P2[j] = p2;
        p1 += v;
        p2 = fma(v, v, p2);
    }
    if (j1 == win && j0 < win) {
        P1[win] = p1;
        P2[win] = p2;
    }
    __syncthreads();
    for (int t = tid; t < kNccTile; t += kNccTileThreads) {
        const long long i = i0 + t;
        if (i >= n) break;
        double cor = 0.0;
        for (int r = 0; r < runs.count; ++r)
            cor = fma(runs.value[r], P1[t + runs.end[r]] - P1[t + runs.start[r]], cor);
        if (normalised) out[i] = cor / sqrt((P2[t + M] - P2[t]) * needle_energy);
        else out[i] = cor;
    }
}

// direct form for arbitrary needles: one output per thread, needle from global (L1-resident)
template <typename T>
__global__ void ncc_direct_kernel(const T *__restrict__ h, long long n, const double *__restrict__ needle,
                                  int M, double needle_energy, int normalised, double *__restrict__ out) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const long long w0 = i - M / 2;
    int k0 = w0 < 0 ? static_cast<int>(-w0) : 0;
    int k1 = w0 + M > n ? static_cast<int>(n - w0) : M;
    double cor = 0.0, sums = 0.0;
    for (int k = k0; k < k1; ++k) {
        const double v = to_double(h[w0 + k]);
        cor = fma(v, needle[k], cor);
        sums = fma(v, v, sums);
    }
    out[i] = normalised ? cor / sqrt(sums * needle_energy) : cor;
}

// ---- radix select on float64 ------------------------------------------------------------
// order-preserving key; NaNs sort above +inf like numpy's partition
__device__ __forceinline__ unsigned long long f64_key(double v) {
    unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}

// histograms of byte `shift/8` over the elements whose higher bytes equal the given prefixes:
// one read of the array serves the search for the k-th largest (hist[0..255], prefix_top) and
// the k-th smallest (hist[256..511], prefix_bot)
__global__ void select_hist2_kernel(const double *__restrict__ x, long long n, unsigned long long prefix_top,
                                    unsigned long long prefix_bot, int shift, unsigned int *__restrict__ hist) {
    __shared__ unsigned int sh[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const unsigned long long himask = shift >= 56 ? 0ULL : (~0ULL << (shift + 8));
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const unsigned long long k = f64_key(x[i]);
        const unsigned long long hi = k & himask;
        const unsigned int d = static_cast<unsigned int>((k >> shift) & 255);
        if (hi == prefix_top) atomicAdd(&sh[d], 1u);
        if (hi == prefix_bot) atomicAdd(&sh[256 + d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// per-block partial sums (and counts) of the values strictly above key_top and strictly below key_bot
__global__ void select_sum2_kernel(const double *__restrict__ x, long long n, unsigned long long key_top,
                                   unsigned long long key_bot, double *__restrict__ part_sum,
                                   unsigned long long *__restrict__ part_cnt) {
    __shared__ double ss[2][256];
    __shared__ unsigned long long sc[2][256];
    double s0 = 0.0, s1 = 0.0;
    unsigned long long c0 = 0, c1 = 0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const double v = x[i];
        const unsigned long long k = f64_key(v);
        if (k > key_top) {
            s0 += v;
            ++c0;
        }
        if (k < key_bot) {
            s1 += v;
            ++c1;
        }
    }
    ss[0][threadIdx.x] = s0;
    ss[1][threadIdx.x] = s1;
    sc[0][threadIdx.x] = c0;
    sc[1][threadIdx.x] = c1;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            ss[0][threadIdx.x] += ss[0][threadIdx.x + off];
            ss[1][threadIdx.x] += ss[1][threadIdx.x + off];
            sc[0][threadIdx.x] += sc[0][threadIdx.x + off];
            sc[1][threadIdx.x] += sc[1][threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part_sum[2 * blockIdx.x] = ss[0][0];
        part_sum[2 * blockIdx.x + 1] = ss[1][0];
        part_cnt[2 * blockIdx.x] = sc[0][0];
        part_cnt[2 * blockIdx.x + 1] = sc[1][0];
    }
}

// ---- ordered compaction of the samples above a threshold --------------------------------
constexpr int kCompactThreads = 256;
constexpr int kCompactPer = 16;      // samples per thread, contiguous

__global__ void compact_count_kernel(const double *__restrict__ x, long long n, double thr,
                                     unsigned int *__restrict__ block_cnt) {
    __shared__ unsigned int s[kCompactThreads];
    const long long base = (static_cast<long long>(blockIdx.x) * kCompactThreads + threadIdx.x) * kCompactPer;
    unsigned int c = 0;
    for (int i = 0; i < kCompactPer; ++i) {
        const long long j = base + i;
        if (j < n && x[j] > thr) ++c;
    }
    s[threadIdx.x] = c;
    __syncthreads();
    for (int off = kCompactThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = s[0];
}

__global__ void compact_write_kernel(const double *__restrict__ x, long long n, double thr,
                                     const unsigned long long *__restrict__ block_off, long long cap,
                                     long long *__restrict__ idx, double *__restrict__ val) {
    __shared__ unsigned int s[kCompactThreads];
    const long long base = (static_cast<long long>(blockIdx.x) * kCompactThreads + threadIdx.x) * kCompactPer;
    unsigned int c = 0;
    for (int i = 0; i < kCompactPer; ++i) {
        const long long j = base + i;
        if (j < n && x[j] > thr) ++c;
    }
    s[threadIdx.x] = c;
    __syncthreads();
    for (int off = 1; off < kCompactThreads; off <<= 1) {
        unsigned int b = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += b;
        __syncthreads();
    }
    long long pos = static_cast<long long>(block_off[blockIdx.x]) + (s[threadIdx.x] - c);
    for (int i = 0; i < kCompactPer; ++i) {
        const long long j = base + i;
        if (j < n && x[j] > thr) {
            if (pos < cap) {
                idx[pos] = j;
                val[pos] = x[j];
            }
            ++pos;
        }
    }
}

// ---- the group-maximum scan of decode_noaa.py:731-746, exactly, without the candidate list ----
// The reference walks the sorted candidates (cor > thr) keeping a running maximum; a candidate at
// distance >= minPkDist from the CURRENT maximum closes the group.  Restated: call a candidate r
// "dominant" when no later sample within ceil(minPkDist) - 1 positions is strictly greater.  The
// running maximum climbs the chain of next-strictly-greater elements and stops at the first
// dominant one, and no chain step can jump over the first dominant candidate at or after the group
// start (if it did, a greater value would sit inside that candidate's window).  Hence
//     peak(group) = first dominant candidate >= group start,
//     next group start = first candidate >= peak + ceil(minPkDist),
// which needs (1) a sliding-window maximum (van Herk / Gil-Werman: prefix and suffix maxima inside
// blocks of the window length), (2) "next flagged index" tables (suffix-min scans), and (3) a walk
// of two loads per peak.  A noisy pass has millions of candidates; none of them leaves the device.
constexpr int kWmThreads = 256;

// F[j] = max(x[block start .. j]), B[j] = max(x[j .. block end]) for blocks of `w` samples.
// One CTA per block walks it in tiles of 256 consecutive samples (coalesced), left to right for F
// and right to left for B: warp-shuffle inclusive max-scan, warp totals through shared memory, and a
// running carry from the tiles before.
__device__ __forceinline__ double shfl_up_f64(double v, int d) {
    return __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(v), d),
                            __shfl_up_sync(0xffffffffu, __double2loint(v), d));
}
__device__ __forceinline__ double shfl_down_f64(double v, int d) {
    return __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(v), d),
                            __shfl_down_sync(0xffffffffu, __double2loint(v), d));
}

__global__ void __launch_bounds__(kWmThreads)
winmax_blocks_kernel(const double *__restrict__ x, long long n, long long w, double *__restrict__ F,
                     double *__restrict__ B) {
    __shared__ double wt[kWmThreads / 32];
    const long long b0 = static_cast<long long>(blockIdx.x) * w;
    const long long b1 = b0 + w < n ? b0 + w : n;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int NW = kWmThreads / 32;
    // prefix maxima
    double carry = -INFINITY;
    for (long long t0 = b0; t0 < b1; t0 += kWmThreads) {
        const long long i = t0 + threadIdx.x;
        double v = i < b1 ? x[i] : -INFINITY;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = shfl_up_f64(v, d);
            if (lane >= d) v = fmax(v, o);
        }
        if (lane == 31) wt[wid] = v;
        __syncthreads();
        double pre = carry;
        for (int k = 0; k < wid; ++k) pre = fmax(pre, wt[k]);
        double tot = carry;
        for (int k = 0; k < NW; ++k) tot = fmax(tot, wt[k]);
        v = fmax(v, pre);
        if (i < b1) F[i] = v;
        carry = tot;
        __syncthreads();
    }
    // suffix maxima: tiles from the right end of the block
    carry = -INFINITY;
    for (long long t1 = b1; t1 > b0; t1 -= kWmThreads) {
        const long long i = t1 - kWmThreads + threadIdx.x;          // may lie before b0 in the last tile
        double v = i >= b0 ? x[i] : -INFINITY;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double o = shfl_down_f64(v, d);
            if (lane + d < 32) v = fmax(v, o);
        }
        if (lane == 0) wt[wid] = v;
        __syncthreads();
        double post = carry;
        for (int k = wid + 1; k < NW; ++k) post = fmax(post, wt[k]);
        double tot = carry;
        for (int k = 0; k < NW; ++k) tot = fmax(tot, wt[k]);
        v = fmax(v, post);
        if (i >= b0) B[i] = v;
        carry = tot;
        __syncthreads();
    }
}

constexpr int kFlagBlock = 1024;
constexpr long long kNoIndex = 0x7FFFFFFFFFFFFFFFLL;

// per sample: candidate / dominant flags, turned into "next flagged index inside this 1024-block"
// tables plus the first flagged index of every block
__global__ void __launch_bounds__(kFlagBlock)
peak_flags_kernel(const double *__restrict__ x, const double *__restrict__ F, const double *__restrict__ B,
                  long long n, long long w, double thr, long long *__restrict__ next_cand,
                  long long *__restrict__ next_dom, long long *__restrict__ first_cand,
                  long long *__restrict__ first_dom) {
    __shared__ long long sc[kFlagBlock], sd[kFlagBlock];
    const long long i = static_cast<long long>(blockIdx.x) * kFlagBlock + threadIdx.x;
    long long c = kNoIndex, d = kNoIndex;
    if (i < n) {
        const double v = x[i];
        if (v > thr) {
            c = i;
            // maximum over (i, i + w], clipped at the end of the array
            double mx = -HUGE_VAL;
            const long long a = i + 1;
            long long b = i + w;
            if (b > n - 1) b = n - 1;
            if (a <= b) {
                const long long ba = a / w, bb = b / w;
                if (ba != bb) {
                    mx = fmax(B[a], F[b]);
                } else if (b == n - 1 || (b + 1) % w == 0) {
                    mx = B[a];                   // the window runs to the end of its block (or of the array)
                } else {
                    // shorter than a block: never for windows of exactly w samples, kept for safety
                    for (long long k = a; k <= b; ++k) mx = fmax(mx, x[k]);
                }
            }
            if (!(mx > v)) d = i;
        }
    }
    sc[threadIdx.x] = c;
    sd[threadIdx.x] = d;
    __syncthreads();
    // suffix minimum inside the block
    for (int off = 1; off < kFlagBlock; off <<= 1) {
        long long oc = kNoIndex, od = kNoIndex;
        if (threadIdx.x + off < kFlagBlock) {
            oc = sc[threadIdx.x + off];
            od = sd[threadIdx.x + off];
        }
        __syncthreads();
        if (oc < sc[threadIdx.x]) sc[threadIdx.x] = oc;
        if (od < sd[threadIdx.x]) sd[threadIdx.x] = od;
        __syncthreads();
    }
    if (i < n) {
        next_cand[i] = sc[threadIdx.x];
        next_dom[i] = sd[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        first_cand[blockIdx.x] = sc[0];
        first_dom[blockIdx.x] = sd[0];
    }
}

// The walk only ever needs the dominant candidates: a dominant index is a candidate, so "first
// dominant >= first candidate >= q" is "first dominant >= q", and a candidate >= q exists exactly
// when a dominant one does (the largest of them).  Hence peak[k+1] = first dominant >= peak[k] +
// ceil(minPkDist), peak[0] = the first dominant -- and dominants are sparse (no two within a window
// unless the values tie), so they are appended to a short list here, unordered, and sorted and walked
// on the host.  The table kernels above remain the form for inputs with more than `cap` dominants
// (long plateaus above the threshold).
__global__ void __launch_bounds__(256)
dominant_append_kernel(const double *__restrict__ x, const double *__restrict__ F, const double *__restrict__ B,
                       long long n, long long w, double thr, long long *__restrict__ list, long long cap,
                       unsigned long long *__restrict__ count) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const double v = x[i];
        if (!(v > thr)) continue;
        double mx = -HUGE_VAL;
        const long long a = i + 1;
        long long b = i + w;
        if (b > n - 1) b = n - 1;
        if (a <= b) {
            const long long ba = a / w, bb = b / w;
            if (ba != bb) {
                mx = fmax(B[a], F[b]);
            } else if (b == n - 1 || (b + 1) % w == 0) {
                mx = B[a];
            } else {
                for (long long k = a; k <= b; ++k) mx = fmax(mx, x[k]);
            }
        }
        if (!(mx > v)) {
            const unsigned long long slot = atomicAdd(count, 1ULL);
            if (slot < static_cast<unsigned long long>(cap)) list[slot] = i;
        }
    }
}

// carry[b] = first flagged index in blocks >= b (suffix minimum over the per-block firsts); one CTA
__global__ void __launch_bounds__(1024)
peak_carry_kernel(long long *__restrict__ first_cand, long long *__restrict__ first_dom, long long nblk) {
    __shared__ long long sc[1024], sd[1024];
    // each thread owns a contiguous range of blocks; ranges are combined right to left
    const long long per = (nblk + 1023) / 1024;
    const long long r0 = static_cast<long long>(threadIdx.x) * per;
    const long long r1 = r0 + per < nblk ? r0 + per : nblk;
    long long mc = kNoIndex, md = kNoIndex;
    for (long long b = r0; b < r1; ++b) {
        if (first_cand[b] < mc) mc = first_cand[b];
        if (first_dom[b] < md) md = first_dom[b];
    }
    sc[threadIdx.x] = mc;
    sd[threadIdx.x] = md;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        long long oc = kNoIndex, od = kNoIndex;
        if (threadIdx.x + off < 1024) {
            oc = sc[threadIdx.x + off];
            od = sd[threadIdx.x + off];
        }
        __syncthreads();
        if (oc < sc[threadIdx.x]) sc[threadIdx.x] = oc;
        if (od < sd[threadIdx.x]) sd[threadIdx.x] = od;
        __syncthreads();
    }
    long long cc = threadIdx.x + 1 < 1024 ? sc[threadIdx.x + 1] : kNoIndex;
    long long cd = threadIdx.x + 1 < 1024 ? sd[threadIdx.x + 1] : kNoIndex;
    for (long long b = r1 - 1; b >= r0; --b) {
        if (first_cand[b] < cc) cc = first_cand[b];
        if (first_dom[b] < cd) cd = first_dom[b];
        first_cand[b] = cc;
        first_dom[b] = cd;
    }
}

__device__ __forceinline__ long long next_flagged(const long long *__restrict__ local, const long long *__restrict__ carry,
                                                  long long nblk, long long i) {
    const long long v = local[i];
    if (v != kNoIndex) return v;
    const long long b = i / kFlagBlock + 1;
    return b < nblk ? carry[b] : kNoIndex;
}

__global__ void peak_walk_kernel(const long long *__restrict__ next_cand, const long long *__restrict__ next_dom,
                                 const long long *__restrict__ carry_cand, const long long *__restrict__ carry_dom,
                                 long long n, long long nblk, long long close_dist, long long *__restrict__ peaks,
                                 long long cap, long long *__restrict__ count) {
    long long np = 0;
    long long s = next_flagged(next_cand, carry_cand, nblk, 0);
    while (s != kNoIndex) {
        const long long d = next_flagged(next_dom, carry_dom, nblk, s);
        if (d == kNoIndex) break;                    // cannot happen: the last candidate is dominant
        if (np < cap) peaks[np] = d;
        ++np;
        const long long p = d + close_dist;
        if (p >= n) break;
        s = next_flagged(next_cand, carry_cand, nblk, p);
    }
    *count = np;
}

// ---- AFSK correlator bank ---------------------------------------------------------------
// out[s] = (sum x[s+k] t0[k])^2 + (sum x[s+k] t1[k])^2 - (sum x[s+k] t2[k])^2 - (sum x[s+k] t3[k])^2
// for s < n - nbuf, 0 for the last nbuf outputs (loop bound of decode_afsk1200.py:129)
template <typename T>
__global__ void bank4_kernel(const T *__restrict__ x, long long n, const double *__restrict__ taps, int nbuf,
                             float *__restrict__ out) {
    extern __shared__ double s_taps[];
    for (int i = threadIdx.x; i < 4 * nbuf; i += blockDim.x) s_taps[i] = taps[i];
    __syncthreads();
    const long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (s >= n) return;
    if (s >= n - nbuf) {
        out[s] = 0.f;
        return;
    }
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int k = 0; k < nbuf; ++k) {
        const double v = to_double(x[s + k]);
        a0 = fma(v, s_taps[k], a0);
        a1 = fma(v, s_taps[nbuf + k], a1);
        a2 = fma(v, s_taps[2 * nbuf + k], a2);
        a3 = fma(v, s_taps[3 * nbuf + k], a3);
    }
    out[s] = static_cast<float>(a0 * a0 + a1 * a1 - a2 * a2 - a3 * a3);
}

}  // namespace ddm

using namespace ddm;

namespace {

int check_device(int device, const char *who) {
    int ndev = 0;
    DDM_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        set_error("%s: no such device %d", who, device);
        return DDM_ERR_INVALID;
    }
    return DDM_OK;
}

// a slot of the per-device scratch pool (capi.cu); nothing to free
struct DevBuf {
    void *p = nullptr;
    int alloc(int device, int slot, size_t bytes) {
        p = scratch_get(device, slot, bytes ? bytes : 1);
        return p ? DDM_OK : DDM_ERR_NOMEM;
    }
};

}  // namespace

extern "C" {

int ddm_correlate(int device, const void *hay_dev, int64_t n, int hay_is_f64, const double *needle_host,
                  int m, int normalised, void *out_f64_dev, void *stream) {
    DDM_REQUIRE(n >= 0 && m >= 1, "ddm_correlate: bad lengths");
    DDM_REQUIRE(needle_host != nullptr, "ddm_correlate: NULL needle");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(hay_dev != nullptr && out_f64_dev != nullptr, "ddm_correlate: NULL buffer");
    int rc = check_device(device, "ddm_correlate");
    if (rc != DDM_OK) return rc;
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double energy = 0.0;
    for (int k = 0; k < m; ++k) energy += needle_host[k] * needle_host[k];
    // run-length form of the needle
    NeedleRuns runs;
    runs.count = 0;
    bool compressible = true;
    for (int k = 0; k < m;) {
        int e = k + 1;
        while (e < m && needle_host[e] == needle_host[k]) ++e;
        if (needle_host[k] != 0.0) {
            if (runs.count == kMaxRuns) {
                compressible = false;
                break;
            }
            runs.start[runs.count] = k;
            runs.end[runs.count] = e;
            runs.value[runs.count] = needle_host[k];
            ++runs.count;
        }
        k = e;
    }
    double *out = static_cast<double *>(out_f64_dev);
    if (compressible && runs.count * 8 <= m && m <= kNccTileMaxM) {
        const size_t smem = sizeof(double) * (2 * (static_cast<size_t>(kNccTile) + m + 1) + 2 * kNccTileThreads);
        const unsigned grid = static_cast<unsigned>((n + kNccTile - 1) / kNccTile);
        if (hay_is_f64) {
            DDM_CUDA(cudaFuncSetAttribute(ncc_tile_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            ncc_tile_kernel<double><<<grid, kNccTileThreads, smem, st>>>(static_cast<const double *>(hay_dev), n, m, energy,
                                                                         normalised, runs, out);
        } else {
            DDM_CUDA(cudaFuncSetAttribute(ncc_tile_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            ncc_tile_kernel<float><<<grid, kNccTileThreads, smem, st>>>(static_cast<const float *>(hay_dev), n, m, energy,
                                                                        normalised, runs, out);
        }
        count_launch();
        DDM_CUDA(cudaGetLastError());
        return DDM_OK;
    }
    if (compressible && runs.count * 8 <= m) {
        const long long nblk = n / kPfxBlock + 1;
        DevBuf L1, L2, T1, T2;
        if ((rc = L1.alloc(device, 0, sizeof(double) * (n + 1))) != DDM_OK) return rc;
        if ((rc = L2.alloc(device, 1, sizeof(double) * (n + 1))) != DDM_OK) return rc;
        if ((rc = T1.alloc(device, 2, sizeof(double) * nblk)) != DDM_OK) return rc;
        if ((rc = T2.alloc(device, 3, sizeof(double) * nblk)) != DDM_OK) return rc;
        if (hay_is_f64)
            block_prefix_kernel<double><<<static_cast<unsigned>(nblk), kPfxThreads, 0, st>>>(
                static_cast<const double *>(hay_dev), n, static_cast<double *>(L1.p), static_cast<double *>(L2.p),
                static_cast<double *>(T1.p), static_cast<double *>(T2.p));
        else
            block_prefix_kernel<float><<<static_cast<unsigned>(nblk), kPfxThreads, 0, st>>>(
                static_cast<const float *>(hay_dev), n, static_cast<double *>(L1.p), static_cast<double *>(L2.p),
                static_cast<double *>(T1.p), static_cast<double *>(T2.p));
        const long long want = (n + 255) / 256;
        const unsigned grid = static_cast<unsigned>(std::min<long long>(want, static_cast<long long>(sm_count(device)) * 16));
        ncc_runs_kernel<<<grid, 256, 0, st>>>(static_cast<const double *>(L1.p), static_cast<const double *>(L2.p),
                                             static_cast<const double *>(T1.p), static_cast<const double *>(T2.p),
                                             n, m, energy, normalised, runs, out);
        count_launch(2);
        DDM_CUDA(cudaGetLastError());
        return DDM_OK;
    }
    DevBuf nd;
    if ((rc = nd.alloc(device, 0, sizeof(double) * m)) != DDM_OK) return rc;
    // (a copy from pageable host memory is staged before cudaMemcpyAsync returns: needle_host is free)
    DDM_CUDA(cudaMemcpyAsync(nd.p, needle_host, sizeof(double) * m, cudaMemcpyHostToDevice, st));
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (hay_is_f64)
        ncc_direct_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double *>(hay_dev), n,
                                                        static_cast<const double *>(nd.p), m, energy, normalised, out);
    else
        ncc_direct_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(hay_dev), n,
                                                       static_cast<const double *>(nd.p), m, energy, normalised, out);
    count_launch();
    DDM_CUDA(cudaGetLastError());
    return DDM_OK;
}

int ddm_topk_sums(int device, const void *x_f64_dev, int64_t n, int64_t k, double *sum_top, double *sum_bottom,
                  void *stream) {
    DDM_REQUIRE(n >= 1 && k >= 1 && k <= n, "ddm_topk_sums: need 1 <= k <= n (k = %lld, n = %lld)",
                static_cast<long long>(k), static_cast<long long>(n));
    DDM_REQUIRE(x_f64_dev != nullptr, "ddm_topk_sums: NULL buffer");
    int rc = check_device(device, "ddm_topk_sums");
    if (rc != DDM_OK) return rc;
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double *x = static_cast<const double *>(x_f64_dev);
    const unsigned grid = static_cast<unsigned>(std::min<long long>((n + 255) / 256, static_cast<long long>(sm_count(device)) * 8));
    DevBuf hist, psum, pcnt;
    if ((rc = hist.alloc(device, 4, sizeof(unsigned int) * 512)) != DDM_OK) return rc;
    if ((rc = psum.alloc(device, 5, sizeof(double) * 2 * grid)) != DDM_OK) return rc;
    if ((rc = pcnt.alloc(device, 6, sizeof(unsigned long long) * 2 * grid)) != DDM_OK) return rc;
    std::vector<unsigned int> h(512);
    // keys of the k-th largest and the k-th smallest element, one digit (byte) per pass, both
    // searches served by the same read of the array
    unsigned long long prefix[2] = {0, 0};
    long long remaining[2] = {static_cast<long long>(k), static_cast<long long>(k)};
    for (int shift = 56; shift >= 0; shift -= 8) {
        DDM_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(unsigned int) * 512, st));
        select_hist2_kernel<<<grid, 256, 0, st>>>(x, n, prefix[0], prefix[1], shift, static_cast<unsigned int *>(hist.p));
        count_launch();
        DDM_CUDA(cudaMemcpyAsync(h.data(), hist.p, sizeof(unsigned int) * 512, cudaMemcpyDeviceToHost, st));
        DDM_CUDA(cudaStreamSynchronize(st));
        for (int which = 0; which < 2; ++which) {
            const bool top = which == 0;
            const unsigned int *hh = h.data() + 256 * which;
            int digit = top ? 255 : 0;
            while (static_cast<long long>(hh[digit]) < remaining[which]) {
                remaining[which] -= hh[digit];
                digit += top ? -1 : 1;
                if (digit < 0 || digit > 255) {
                    set_error("ddm_topk_sums: internal selection error");
                    return DDM_ERR_INVALID;
                }
            }
            prefix[which] |= static_cast<unsigned long long>(digit) << shift;
        }
    }
    select_sum2_kernel<<<grid, 256, 0, st>>>(x, n, prefix[0], prefix[1], static_cast<double *>(psum.p),
                                             static_cast<unsigned long long *>(pcnt.p));
    count_launch();
    std::vector<double> hs(2 * grid);
    std::vector<unsigned long long> hc(2 * grid);
    DDM_CUDA(cudaMemcpyAsync(hs.data(), psum.p, sizeof(double) * 2 * grid, cudaMemcpyDeviceToHost, st));
    DDM_CUDA(cudaMemcpyAsync(hc.data(), pcnt.p, sizeof(unsigned long long) * 2 * grid, cudaMemcpyDeviceToHost, st));
    DDM_CUDA(cudaStreamSynchronize(st));
    for (int which = 0; which < 2; ++which) {
        double *result = which == 0 ? sum_top : sum_bottom;
        if (!result) continue;
        double sum = 0.0;
        unsigned long long c = 0;
        for (unsigned i = 0; i < grid; ++i) {
            sum += hs[2 * i + which];
            c += hc[2 * i + which];
        }
        // value of the k-th element from its key; (k - c) copies of it complete the selection
        const unsigned long long key = prefix[which];
        unsigned long long u = (key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFULL) : ~key;
        double kth;
        std::memcpy(&kth, &u, sizeof(kth));
        *result = sum + kth * static_cast<double>(static_cast<long long>(k) - static_cast<long long>(c));
    }
    DDM_CUDA(cudaGetLastError());
    return DDM_OK;
}

int ddm_compact_above(int device, const void *x_f64_dev, int64_t n, double threshold, void *idx_i64_dev,
                      void *val_f64_dev, int64_t capacity, int64_t *count, void *stream) {
    DDM_REQUIRE(n >= 0 && capacity >= 0 && count != nullptr, "ddm_compact_above: bad arguments");
    *count = 0;
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_f64_dev != nullptr, "ddm_compact_above: NULL buffer");
    int rc = check_device(device, "ddm_compact_above");
    if (rc != DDM_OK) return rc;
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double *x = static_cast<const double *>(x_f64_dev);
    const long long per_block = static_cast<long long>(kCompactThreads) * kCompactPer;
    const long long blocks = (n + per_block - 1) / per_block;
    DevBuf cnt, off;
    if ((rc = cnt.alloc(device, 4, sizeof(unsigned int) * blocks)) != DDM_OK) return rc;
    if ((rc = off.alloc(device, 5, sizeof(unsigned long long) * blocks)) != DDM_OK) return rc;
    compact_count_kernel<<<static_cast<unsigned>(blocks), kCompactThreads, 0, st>>>(x, n, threshold,
                                                                                    static_cast<unsigned int *>(cnt.p));
    count_launch();
    std::vector<unsigned int> hc(blocks);
    DDM_CUDA(cudaMemcpyAsync(hc.data(), cnt.p, sizeof(unsigned int) * blocks, cudaMemcpyDeviceToHost, st));
    DDM_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned long long> ho(blocks);
    unsigned long long total = 0;
    for (long long b = 0; b < blocks; ++b) {
        ho[b] = total;
        total += hc[b];
    }
    *count = static_cast<int64_t>(total);
    if (total == 0 || capacity == 0) return DDM_OK;
    DDM_REQUIRE(idx_i64_dev != nullptr && val_f64_dev != nullptr, "ddm_compact_above: NULL output");
    DDM_CUDA(cudaMemcpyAsync(off.p, ho.data(), sizeof(unsigned long long) * blocks, cudaMemcpyHostToDevice, st));
    compact_write_kernel<<<static_cast<unsigned>(blocks), kCompactThreads, 0, st>>>(
        x, n, threshold, static_cast<const unsigned long long *>(off.p), capacity,
        static_cast<long long *>(idx_i64_dev), static_cast<double *>(val_f64_dev));
    count_launch();
    DDM_CUDA(cudaGetLastError());
    DDM_CUDA(cudaStreamSynchronize(st));
    return DDM_OK;
}

int ddm_group_peaks(const int64_t *idx, const double *val, int64_t count, double min_dist, int64_t *peaks,
                    int64_t capacity, int64_t *n_peaks) {
    DDM_REQUIRE(count >= 0 && n_peaks != nullptr, "ddm_group_peaks: bad arguments");
    *n_peaks = 0;
    if (count == 0) return DDM_OK;
    DDM_REQUIRE(idx != nullptr && val != nullptr, "ddm_group_peaks: NULL input");
    // decode_noaa.py:731-746: running maximum, closed when a candidate is >= min_dist beyond the
    // CURRENT maximum; strict '<' so the first of equal maxima wins
    bool have = false;
    double cur_max = 0.0;
    int64_t cur_idx = 0;
    int64_t np = 0;
    auto emit = [&](int64_t v) {
        if (np < capacity && peaks) peaks[np] = v;
        ++np;
    };
    for (int64_t i = 0; i < count; ++i) {
        if (have && static_cast<double>(idx[i] - cur_idx) >= min_dist) {
            emit(cur_idx);
            have = false;
        }
        if (!have || cur_max < val[i]) {
            cur_max = val[i];
            cur_idx = idx[i];
            have = true;
        }
    }
    emit(cur_idx);
    *n_peaks = np;
    if (np > capacity) {
        set_error("ddm_group_peaks: %lld peaks, capacity %lld", static_cast<long long>(np),
                  static_cast<long long>(capacity));
        return DDM_ERR_CAPACITY;
    }
    return DDM_OK;
}

int ddm_pick_peaks(int device, const void *x_f64_dev, int64_t n, double threshold, double min_dist,
                   int64_t *peaks_host, int64_t capacity, int64_t *n_peaks, void *stream) {
    DDM_REQUIRE(n >= 0 && capacity >= 0 && n_peaks != nullptr, "ddm_pick_peaks: bad arguments");
    *n_peaks = 0;
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_f64_dev != nullptr, "ddm_pick_peaks: NULL buffer");
    DDM_REQUIRE(min_dist > 1.0, "ddm_pick_peaks: the minimum peak distance must exceed one sample");
    int rc = check_device(device, "ddm_pick_peaks");
    if (rc != DDM_OK) return rc;
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double *x = static_cast<const double *>(x_f64_dev);
    const long long close_dist = static_cast<long long>(std::ceil(min_dist));   // i - r >= min_dist
    long long w = close_dist - 1;                                               // records possible up to here
    if (w > n) w = n;
    const long long wblocks = (n + w - 1) / w;
    const long long nblk = (n + kFlagBlock - 1) / kFlagBlock;
    DevBuf F, B, nc, nd, fc, fd, out;
    if ((rc = F.alloc(device, 0, sizeof(double) * n)) != DDM_OK) return rc;
    if ((rc = B.alloc(device, 1, sizeof(double) * n)) != DDM_OK) return rc;
    winmax_blocks_kernel<<<static_cast<unsigned>(wblocks), kWmThreads, 0, st>>>(x, n, w, static_cast<double *>(F.p),
                                                                                static_cast<double *>(B.p));
    count_launch();
    // sparse path: the dominant candidates as a short unordered list, sorted and walked on the host
    // (DDM_PEAKS_DENSE=1 forces the table path below: used by the tests to cross-check the two)
    if (std::getenv("DDM_PEAKS_DENSE") == nullptr) {
        const long long dcap = 1 << 20;
        if ((rc = out.alloc(device, 6, sizeof(long long) * (dcap + 1))) != DDM_OK) return rc;
        long long *list = static_cast<long long *>(out.p);
        DDM_CUDA(cudaMemsetAsync(list, 0, sizeof(long long), st));
        dominant_append_kernel<<<static_cast<unsigned>(sm_count(device)) * 8, 256, 0, st>>>(
            x, static_cast<const double *>(F.p), static_cast<const double *>(B.p), n, w, threshold, list + 1, dcap,
            reinterpret_cast<unsigned long long *>(list));
        count_launch();
        DDM_CUDA(cudaGetLastError());
        long long nd_found = 0;
        DDM_CUDA(cudaMemcpyAsync(&nd_found, list, sizeof(nd_found), cudaMemcpyDeviceToHost, st));
        DDM_CUDA(cudaStreamSynchronize(st));
        if (nd_found <= dcap) {
            std::vector<long long> dom(static_cast<size_t>(nd_found));
            if (nd_found > 0) {
                DDM_CUDA(cudaMemcpyAsync(dom.data(), list + 1, sizeof(long long) * nd_found, cudaMemcpyDeviceToHost, st));
                DDM_CUDA(cudaStreamSynchronize(st));
            }
            std::sort(dom.begin(), dom.end());
            long long np = 0;
            auto it = dom.begin();
            while (it != dom.end()) {
                if (np < capacity) {
                    DDM_REQUIRE(peaks_host != nullptr, "ddm_pick_peaks: NULL output");
                    peaks_host[np] = *it;
                }
                ++np;
                const long long p = *it + close_dist;
                if (p >= n) break;
                it = std::lower_bound(it, dom.end(), p);
            }
            *n_peaks = np;
            if (np > capacity) {
                set_error("ddm_pick_peaks: %lld peaks, capacity %lld", np, static_cast<long long>(capacity));
                return DDM_ERR_CAPACITY;
            }
            return DDM_OK;
        }
    }
    // dense path (more than 2^20 dominant candidates: plateaus): next-index tables and a device walk
    if ((rc = nc.alloc(device, 2, sizeof(long long) * n)) != DDM_OK) return rc;
    if ((rc = nd.alloc(device, 3, sizeof(long long) * n)) != DDM_OK) return rc;
    if ((rc = fc.alloc(device, 4, sizeof(long long) * nblk)) != DDM_OK) return rc;
    if ((rc = fd.alloc(device, 5, sizeof(long long) * nblk)) != DDM_OK) return rc;
    if ((rc = out.alloc(device, 6, sizeof(long long) * (capacity + 1))) != DDM_OK) return rc;
    peak_flags_kernel<<<static_cast<unsigned>(nblk), kFlagBlock, 0, st>>>(
        x, static_cast<const double *>(F.p), static_cast<const double *>(B.p), n, w, threshold,
        static_cast<long long *>(nc.p), static_cast<long long *>(nd.p), static_cast<long long *>(fc.p),
        static_cast<long long *>(fd.p));
    peak_carry_kernel<<<1, 1024, 0, st>>>(static_cast<long long *>(fc.p), static_cast<long long *>(fd.p), nblk);
    long long *peaks_dev = static_cast<long long *>(out.p);
    peak_walk_kernel<<<1, 1, 0, st>>>(static_cast<const long long *>(nc.p), static_cast<const long long *>(nd.p),
                                     static_cast<const long long *>(fc.p), static_cast<const long long *>(fd.p), n, nblk,
                                     close_dist, peaks_dev + 1, capacity, peaks_dev);
    count_launch(3);
    DDM_CUDA(cudaGetLastError());
    long long cnt = 0;
    DDM_CUDA(cudaMemcpyAsync(&cnt, peaks_dev, sizeof(cnt), cudaMemcpyDeviceToHost, st));
    DDM_CUDA(cudaStreamSynchronize(st));
    *n_peaks = cnt;
    const long long take = cnt < capacity ? cnt : capacity;
    if (take > 0) {
        DDM_REQUIRE(peaks_host != nullptr, "ddm_pick_peaks: NULL output");
        DDM_CUDA(cudaMemcpyAsync(peaks_host, peaks_dev + 1, sizeof(long long) * take, cudaMemcpyDeviceToHost, st));
        DDM_CUDA(cudaStreamSynchronize(st));
    }
    if (cnt > capacity) {
        set_error("ddm_pick_peaks: %lld peaks, capacity %lld", cnt, static_cast<long long>(capacity));
        return DDM_ERR_CAPACITY;
    }
    return DDM_OK;
}

int ddm_bank4(int device, const void *x_dev, int64_t n, int x_is_f64, const double *taps4_host, int nbuf,
              void *out_f32_dev, void *stream) {
    DDM_REQUIRE(n >= 0 && nbuf >= 1, "ddm_bank4: bad lengths");
    DDM_REQUIRE(taps4_host != nullptr, "ddm_bank4: NULL taps");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_f32_dev != nullptr, "ddm_bank4: NULL buffer");
    DDM_REQUIRE(nbuf <= 4096, "ddm_bank4: at most 4096 taps per correlator");
    int rc = check_device(device, "ddm_bank4");
    if (rc != DDM_OK) return rc;
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevBuf t;
    if ((rc = t.alloc(device, 0, sizeof(double) * 4 * nbuf)) != DDM_OK) return rc;
    DDM_CUDA(cudaMemcpyAsync(t.p, taps4_host, sizeof(double) * 4 * nbuf, cudaMemcpyHostToDevice, st));
    const size_t smem = sizeof(double) * 4 * nbuf;
    if (smem > 48 * 1024) {
        DDM_CUDA(cudaFuncSetAttribute(bank4_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        DDM_CUDA(cudaFuncSetAttribute(bank4_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (x_is_f64)
        bank4_kernel<double><<<grid, 256, smem, st>>>(static_cast<const double *>(x_dev), n,
                                                      static_cast<const double *>(t.p), nbuf, static_cast<float *>(out_f32_dev));
    else
        bank4_kernel<float><<<grid, 256, smem, st>>>(static_cast<const float *>(x_dev), n,
                                                     static_cast<const double *>(t.p), nbuf, static_cast<float *>(out_f32_dev));
    count_launch();
    DDM_CUDA(cudaGetLastError());
    DDM_CUDA(cudaStreamSynchronize(st));
    return DDM_OK;
}

}  // extern "C"
