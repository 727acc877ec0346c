// Stand-alone element-wise operators of the hot path (each also exists fused inside
// chain.cu; these serve arbitrary user chains built from the reference's fluent API).
//
//   ddm_mix_cf32      commSignal.offsetFreq         comm.py:63-78
//   ddm_mix_var_cf32  offsetFreq with a per-sample frequency array (decode_funcube.py:228)
//   ddm_fm_demod      demod_fm.demod                demod_fm.py:29-51
//   ddm_fm_angle_diff demod_fmAD.demod              demod_fm.py:74-96
//   ddm_abs           np.abs (demod_am.py:29,62)
//   ddm_stride_copy   x[off::j] of bwLim            comm.py:127
//   ddm_cu8_to_cf32   source.read                   source.py:117-118, :209-210
//
// All of them are HBM-bound streaming kernels: 16-byte vector accesses where alignment
// allows, grid = SMs x 8 CTAs grid-striding over the array.
#include "ddm_common.cuh"

namespace ddm {

constexpr int kOpsThreads = 256;

static inline unsigned ops_grid(int device, long long work_items) {
    long long blocks = (work_items + kOpsThreads - 1) / kOpsThreads;
    const long long cap = static_cast<long long>(sm_count(device)) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<unsigned>(blocks);
}

// ---- mixer -----------------------------------------------------------------------------
// Each thread handles 2 consecutive samples (one 16-byte access); the rotator of the first
// is computed from the exactly reduced double-double phase, the second is first * step.
__global__ void mix_kernel(float2 *x, long long n, long long n0, double r_hi, double r_lo,
                           float2 step) {
    const long long pairs = n >> 1;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < pairs;
         p += stride) {
        float4 v = reinterpret_cast<float4 *>(x)[p];
        const float2 w0 = phase_rotator(r_hi, r_lo, n0 + 2 * p);
        const float2 w1 = cmul(w0, step);
        const float2 a = cmul(make_float2(v.x, v.y), w0);
        const float2 b = cmul(make_float2(v.z, v.w), w1);
        reinterpret_cast<float4 *>(x)[p] = make_float4(a.x, a.y, b.x, b.y);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        x[n - 1] = cmul(x[n - 1], phase_rotator(r_hi, r_lo, n0 + n - 1));
    }
}

// generic (unaligned base pointer): one sample per thread
__global__ void mix_kernel_scalar(float2 *x, long long n, long long n0, double r_hi, double r_lo) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
        x[i] = cmul(x[i], phase_rotator(r_hi, r_lo, n0 + i));
}

// per-sample frequency: phase = f[i] * (n0 + i) / fs turns, reduced in float64
__global__ void mix_var_kernel(float2 *x, const double *f, long long n, long long n0, double fs) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const double r = f[i] / fs;
        const double e = fma(-r, fs, f[i]) / fs;                 // residual of the division
        x[i] = cmul(x[i], phase_rotator(r, e, n0 + i));
    }
}

// ---- FM discriminator ------------------------------------------------------------------
// out[i - first] = arg(x[i] * conj(x[i-1])), i = first..n-1, where x[-1] = *prev when given.
__global__ void fm_kernel(const float2 *x, const float2 *prev, float *out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long first = prev ? 0 : 1;
    for (long long i = first + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += stride) {
        const float2 c = x[i];
        const float2 p = i > 0 ? x[i - 1] : *prev;
        const float re = fmaf(c.x, p.x, c.y * p.y);
        const float im = fmaf(c.y, p.x, -c.x * p.y);
        out[i - first] = atan2f(im, re);
    }
}

// demod_fmAD: diff(unwrap(angle(x))).  The cumulative 2*pi corrections of np.unwrap cancel
// in the difference, so every output only needs its own two angles:
//   d = a[i] - a[i-1];  dd = mod(d + pi, 2 pi) - pi;  if (dd == -pi && d > 0) dd = pi;
//   out = |d| < pi ? d : dd                       (numpy/lib/function_base.py unwrap)
// The carried state is the last complex SAMPLE, not its angle rounded to float32: a[i-1] is then
// evaluated in float64 from the same bits whether i-1 lies in this chunk or in the previous one, so
// the result does not depend on where the stream was cut (the reference carries a float64 angle).
__global__ void fm_ad_kernel(const float2 *x, const float2 *prev_sample, float *out, float2 *last_sample,
                             long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long first = prev_sample ? 0 : 1;
    const double pi = 3.14159265358979323846;
    for (long long i = first + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += stride) {
        const double a1 = atan2(static_cast<double>(x[i].y), static_cast<double>(x[i].x));
        const double a0 = i > 0 ? atan2(static_cast<double>(x[i - 1].y), static_cast<double>(x[i - 1].x))
                                : atan2(static_cast<double>(prev_sample->y), static_cast<double>(prev_sample->x));
        const double d = a1 - a0;
        double dd = d + pi;
        dd = dd - floor(dd / (2 * pi)) * (2 * pi) - pi;
        if (dd == -pi && d > 0) dd = pi;
        out[i - first] = static_cast<float>(fabs(d) < pi ? d : dd);
        if (i == n - 1 && last_sample) *last_sample = x[i];
    }
}

// ---- misc ------------------------------------------------------------------------------
template <typename T>
__global__ void abs_kernel(const T *x, float *out, long long n);

template <>
__global__ void abs_kernel<float2>(const float2 *x, float *out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
        out[i] = hypotf(x[i].x, x[i].y);
}

template <>
__global__ void abs_kernel<float>(const float *x, float *out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
        out[i] = fabsf(x[i]);
}

__global__ void sign_kernel(const float *x, float *out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
        const float v = x[i];
        out[i] = v > 0.f ? 1.f : (v < 0.f ? -1.f : v);       // np.sign: 0 -> 0, nan -> nan
    }
}

template <typename T>
__global__ void stride_copy_kernel(const T *x, T *out, long long m, long long off, long long step) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < m; i += stride)
        out[i] = x[off + i * step];
}

// 8 samples (16 input bytes) per thread
__global__ void cu8_kernel(const unsigned char *in, float2 *out, long long n) {
    const long long groups = n >> 3;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < groups;
         g += stride) {
        const uint4 v = reinterpret_cast<const uint4 *>(in)[g];
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float4 *o = reinterpret_cast<float4 *>(out + g * 8);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k] = make_float4(static_cast<float>(w[k] & 255u) - 127.5f,
                               static_cast<float>((w[k] >> 8) & 255u) - 127.5f,
                               static_cast<float>((w[k] >> 16) & 255u) - 127.5f,
                               static_cast<float>(w[k] >> 24) - 127.5f);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long i = groups << 3; i < n; ++i)
            out[i] = make_float2(static_cast<float>(in[2 * i]) - 127.5f,
                                 static_cast<float>(in[2 * i + 1]) - 127.5f);
    }
}

__global__ void cu8_kernel_scalar(const unsigned char *in, float2 *out, long long n) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride)
        out[i] = make_float2(static_cast<float>(in[2 * i]) - 127.5f,
                             static_cast<float>(in[2 * i + 1]) - 127.5f);
}

// ---- operators over many equal-length rows (batched accurate sync, decode_noaa.py:844-877) ----
// mixer restarting its sample index at every row (a commSignal without a chunker starts at 0)
__global__ void mix_rows_kernel(float2 *x, long long row_len, long long total, double r_hi, double r_lo) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += stride)
        x[i] = cmul(x[i], phase_rotator(r_hi, r_lo, i % row_len));
}

// out[r][i] = arg(x[r][i+1] conj x[r][i]), i < row_len - 1 (a fresh demod_fm per row)
__global__ void fm_rows_kernel(const float2 *__restrict__ x, long long row_len, long long rows, float *__restrict__ out) {
    const long long per = row_len - 1;
    const long long total = per * rows;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total; t += stride) {
        const long long r = t / per, i = t - r * per;
        const float2 p = x[r * row_len + i], c = x[r * row_len + i + 1];
        out[t] = atan2f(fmaf(c.y, p.x, -c.x * p.y), fmaf(c.x, p.x, c.y * p.y));
    }
}

// first index of the maximum of every row of a float64 matrix (strict '<' scan order: the first
// of equal maxima wins, like decode_noaa.py:742), and the maximum itself
__global__ void __launch_bounds__(256)
rows_argmax_kernel(const double *__restrict__ x, long long row_stride, long long row_len, long long *__restrict__ idx,
                   double *__restrict__ val) {
    __shared__ double sv[256];
    __shared__ long long si[256];
    const double *row = x + static_cast<size_t>(blockIdx.x) * row_stride;
    double best = -INFINITY;
    long long bi = -1;
    for (long long i = threadIdx.x; i < row_len; i += blockDim.x) {
        const double v = row[i];
        if (bi < 0 || v > best) {          // ascending i per thread: '>' keeps the first maximum
            best = v;
            bi = i;
        }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            const double ov = sv[threadIdx.x + off];
            const long long oi = si[threadIdx.x + off];
            const double mv = sv[threadIdx.x];
            const long long mi = si[threadIdx.x];
            if (oi >= 0 && (mi < 0 || ov > mv || (ov == mv && oi < mi))) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        idx[blockIdx.x] = si[0];
        val[blockIdx.x] = sv[0];
    }
}

// mean of x[start[r] : start[r] + len] per row (np.average of a float32 slice, accumulated in f64)
__global__ void __launch_bounds__(256)
rows_mean_kernel(const float *__restrict__ x, const long long *__restrict__ start, long long len, double *__restrict__ out) {
    __shared__ double ss[256];
    const float *p = x + start[blockIdx.x];
    double s = 0.0;
    for (long long i = threadIdx.x; i < len; i += blockDim.x) s += static_cast<double>(p[i]);
    ss[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) ss[threadIdx.x] += ss[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = ss[0] / static_cast<double>(len);
}

// scipy.signal.medfilt(x, k): median of the k-sample window centred on each sample, zeros
// outside the array (filters.py:322-326).  One thread per output; the window is kept sorted by
// insertion in local memory (k is small: the reference's default is 5).
constexpr int kMedMaxK = 255;
__global__ void medfilt_kernel(const float *__restrict__ x, long long n, int k, float *__restrict__ out) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float w[kMedMaxK];
    const int half = k / 2;
    for (int j = 0; j < k; ++j) {
        const long long g = i - half + j;
        const float v = (g >= 0 && g < n) ? x[g] : 0.f;
        int p = j;
        while (p > 0 && w[p - 1] > v) {
            w[p] = w[p - 1];
            --p;
        }
        w[p] = v;
    }
    out[i] = w[half];
}

// np.median of many rows of one array: out[r] = median(x[start[r] : start[r] + len]).
// Short rows (<= 64): one thread per row, insertion sort in local memory.
constexpr int kRowMedSmall = 64;
__global__ void row_median_small_kernel(const float *__restrict__ x, const long long *__restrict__ start,
                                        long long rows, long long stride, int len, double *__restrict__ out) {
    const long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (r >= rows) return;
    const float *p = x + (start ? start[r] : r * stride);
    float w[kRowMedSmall];
    for (int j = 0; j < len; ++j) {
        const float v = p[j];
        int q = j;
        while (q > 0 && w[q - 1] > v) {
            w[q] = w[q - 1];
            --q;
        }
        w[q] = v;
    }
    out[r] = (len & 1) ? static_cast<double>(w[len / 2])
                       : 0.5 * (static_cast<double>(w[len / 2 - 1]) + static_cast<double>(w[len / 2]));
}

// Long rows: one CTA per row, 4-pass radix select on order-preserving float keys (shared-memory
// histograms); for an even length both middle elements are found and averaged like numpy.
__device__ __forceinline__ unsigned int f32_key(float v) {
    const unsigned int u = __float_as_uint(v);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_from_key(unsigned int k) {
    return __uint_as_float((k >> 31) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(256)
row_median_select_kernel(const float *__restrict__ x, const long long *__restrict__ start, long long stride,
                         long long len, double *__restrict__ out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_rank;
    const float *p = x + (start ? start[blockIdx.x] : blockIdx.x * stride);
    double acc = 0.0;
    const int picks = (len & 1) ? 1 : 2;
    for (int pick = 0; pick < picks; ++pick) {
        // rank (0-based, ascending) of the element wanted
        long long want = (len & 1) ? len / 2 : len / 2 - 1 + pick;
        unsigned int prefix = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const unsigned int himask = shift >= 24 ? 0u : (~0u << (shift + 8));
            for (long long i = threadIdx.x; i < len; i += blockDim.x) {
                const unsigned int k = f32_key(p[i]);
                if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                long long rem = want;
                int d = 0;
                while (static_cast<long long>(hist[d]) <= rem) {
                    rem -= hist[d];
                    ++d;
                }
                s_prefix = prefix | (static_cast<unsigned int>(d) << shift);
                s_rank = static_cast<unsigned int>(rem);
            }
            __syncthreads();
            prefix = s_prefix;
            want = s_rank;
            __syncthreads();
        }
        acc += static_cast<double>(f32_from_key(prefix));
    }
    if (threadIdx.x == 0) out[blockIdx.x] = acc / picks;
}

}  // namespace ddm

using namespace ddm;

#define DDM_CHECK_DEVICE(device, who)                                                   \
    do {                                                                                \
        int ndev__ = 0;                                                                 \
        DDM_CUDA(cudaGetDeviceCount(&ndev__));                                          \
        DDM_REQUIRE((device) >= 0 && (device) < ndev__, who ": no such device %d", (device)); \
    } while (0)

extern "C" {

int ddm_mix_cf32(int device, void *x_dev, int64_t n, double freq_offset, double samp_rate,
                 int64_t n0, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_mix_cf32: negative length");
    DDM_REQUIRE(samp_rate > 0, "ddm_mix_cf32: sampling rate must be positive");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr, "ddm_mix_cf32: NULL signal");
    DDM_CHECK_DEVICE(device, "ddm_mix_cf32");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double r_hi = freq_offset / samp_rate;
    const double r_lo = std::fma(-r_hi, samp_rate, freq_offset) / samp_rate;
    if ((reinterpret_cast<uintptr_t>(x_dev) & 15) == 0) {
        const double t = r_hi - std::rint(r_hi);
        const float2 step = make_float2(static_cast<float>(std::cos(2.0 * M_PI * t)),
                                        static_cast<float>(-std::sin(2.0 * M_PI * t)));
        mix_kernel<<<ops_grid(device, (n + 1) / 2), kOpsThreads, 0, st>>>(
            static_cast<float2 *>(x_dev), n, n0, r_hi, r_lo, step);
    } else {
        mix_kernel_scalar<<<ops_grid(device, n), kOpsThreads, 0, st>>>(static_cast<float2 *>(x_dev), n,
                                                                        n0, r_hi, r_lo);
    }
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_mix_var_cf32(int device, void *x_dev, const double *freq_dev, int64_t n, double samp_rate,
                     int64_t n0, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_mix_var_cf32: negative length");
    DDM_REQUIRE(samp_rate > 0, "ddm_mix_var_cf32: sampling rate must be positive");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && freq_dev != nullptr, "ddm_mix_var_cf32: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_mix_var_cf32");
    DeviceGuard guard(device);
    mix_var_kernel<<<ops_grid(device, n), kOpsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float2 *>(x_dev), freq_dev, n, n0, samp_rate);
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_fm_demod(int device, const void *x_dev, int64_t n, const void *prev_dev, void *out_dev,
                 int64_t *n_out, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_fm_demod: negative length");
    const int64_t m = prev_dev ? n : (n > 0 ? n - 1 : 0);
    if (n_out) *n_out = m;
    if (m == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_fm_demod: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_fm_demod");
    DeviceGuard guard(device);
    // (a 4-samples-per-thread variant measured slower: the kernel is balanced between the
    // atan2f issue rate and HBM, not limited by load width)
    fm_kernel<<<ops_grid(device, m), kOpsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2 *>(x_dev), static_cast<const float2 *>(prev_dev),
        static_cast<float *>(out_dev), n);
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_fm_angle_diff(int device, const void *x_dev, int64_t n, const void *prev_sample_dev,
                      void *out_dev, void *last_sample_dev, int64_t *n_out, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_fm_angle_diff: negative length");
    const int64_t m = prev_sample_dev ? n : (n > 0 ? n - 1 : 0);
    if (n_out) *n_out = m;
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr, "ddm_fm_angle_diff: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_fm_angle_diff");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (m > 0) {
        DDM_REQUIRE(out_dev != nullptr, "ddm_fm_angle_diff: NULL output");
        fm_ad_kernel<<<ops_grid(device, m), kOpsThreads, 0, st>>>(
            static_cast<const float2 *>(x_dev), static_cast<const float2 *>(prev_sample_dev),
            static_cast<float *>(out_dev), static_cast<float2 *>(last_sample_dev), n);
        DDM_CUDA(cudaGetLastError());
        count_launch();
    } else if (last_sample_dev) {
        // a single sample and no predecessor: only the carried sample is produced
        DDM_CUDA(cudaMemcpyAsync(last_sample_dev, x_dev, sizeof(float2), cudaMemcpyDeviceToDevice, st));
    }
    return DDM_OK;
}

int ddm_abs(int device, const void *x_dev, int64_t n, int is_complex, void *out_dev, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_abs: negative length");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_abs: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_abs");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (is_complex)
        abs_kernel<float2><<<ops_grid(device, n), kOpsThreads, 0, st>>>(
            static_cast<const float2 *>(x_dev), static_cast<float *>(out_dev), n);
    else
        abs_kernel<float><<<ops_grid(device, n), kOpsThreads, 0, st>>>(
            static_cast<const float *>(x_dev), static_cast<float *>(out_dev), n);
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_sign(int device, const void *x_dev, int64_t n, void *out_dev, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_sign: negative length");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_sign: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_sign");
    DeviceGuard guard(device);
    sign_kernel<<<ops_grid(device, n), kOpsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(x_dev), static_cast<float *>(out_dev), n);
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_stride_copy(int device, const void *x_dev, int64_t n, int elem_bytes, int64_t offset,
                    int64_t step, void *out_dev, int64_t *n_out, void *stream) {
    DDM_REQUIRE(n >= 0 && offset >= 0 && step >= 1, "ddm_stride_copy: bad length/offset/step");
    DDM_REQUIRE(elem_bytes == 4 || elem_bytes == 8 || elem_bytes == 16,
                "ddm_stride_copy: element size must be 4, 8 or 16 bytes");
    const int64_t m = n > offset ? (n - offset + step - 1) / step : 0;
    if (n_out) *n_out = m;
    if (m == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_stride_copy: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_stride_copy");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = ops_grid(device, m);
    if (elem_bytes == 4)
        stride_copy_kernel<float><<<grid, kOpsThreads, 0, st>>>(
            static_cast<const float *>(x_dev), static_cast<float *>(out_dev), m, offset, step);
    else if (elem_bytes == 8)
        stride_copy_kernel<float2><<<grid, kOpsThreads, 0, st>>>(
            static_cast<const float2 *>(x_dev), static_cast<float2 *>(out_dev), m, offset, step);
    else
        stride_copy_kernel<double2><<<grid, kOpsThreads, 0, st>>>(
            static_cast<const double2 *>(x_dev), static_cast<double2 *>(out_dev), m, offset, step);
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_mix_rows_cf32(int device, void *x_dev, int64_t rows, int64_t row_len, double freq_offset, double samp_rate,
                      void *stream) {
    DDM_REQUIRE(rows >= 0 && row_len >= 0, "ddm_mix_rows_cf32: bad sizes");
    DDM_REQUIRE(samp_rate > 0, "ddm_mix_rows_cf32: sampling rate must be positive");
    if (rows == 0 || row_len == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr, "ddm_mix_rows_cf32: NULL signal");
    DDM_CHECK_DEVICE(device, "ddm_mix_rows_cf32");
    DeviceGuard guard(device);
    const double r_hi = freq_offset / samp_rate;
    const double r_lo = std::fma(-r_hi, samp_rate, freq_offset) / samp_rate;
    mix_rows_kernel<<<ops_grid(device, rows * row_len), kOpsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<float2 *>(x_dev), row_len, rows * row_len, r_hi, r_lo);
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_fm_demod_rows(int device, const void *x_dev, int64_t rows, int64_t row_len, void *out_dev, void *stream) {
    DDM_REQUIRE(rows >= 0 && row_len >= 0, "ddm_fm_demod_rows: bad sizes");
    if (rows == 0 || row_len < 2) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_fm_demod_rows: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_fm_demod_rows");
    DeviceGuard guard(device);
    fm_rows_kernel<<<ops_grid(device, rows * (row_len - 1)), kOpsThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2 *>(x_dev), row_len, rows, static_cast<float *>(out_dev));
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_rows_argmax(int device, const void *x_f64_dev, int64_t rows, int64_t row_stride, int64_t row_len,
                    void *idx_i64_dev, void *val_f64_dev, void *stream) {
    DDM_REQUIRE(rows >= 0 && row_len >= 1 && row_stride >= row_len, "ddm_rows_argmax: bad sizes");
    if (rows == 0) return DDM_OK;
    DDM_REQUIRE(x_f64_dev && idx_i64_dev && val_f64_dev, "ddm_rows_argmax: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_rows_argmax");
    DeviceGuard guard(device);
    rows_argmax_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const double *>(x_f64_dev), row_stride, row_len, static_cast<long long *>(idx_i64_dev),
        static_cast<double *>(val_f64_dev));
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_rows_mean(int device, const void *x_dev, const int64_t *row_start_dev, int64_t rows, int64_t len,
                  void *out_f64_dev, void *stream) {
    DDM_REQUIRE(rows >= 0 && len >= 1, "ddm_rows_mean: bad sizes");
    if (rows == 0) return DDM_OK;
    DDM_REQUIRE(x_dev && row_start_dev && out_f64_dev, "ddm_rows_mean: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_rows_mean");
    DeviceGuard guard(device);
    rows_mean_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(x_dev), reinterpret_cast<const long long *>(row_start_dev), len,
        static_cast<double *>(out_f64_dev));
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_medfilt(int device, const void *x_dev, int64_t n, int kernel_size, void *out_dev, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_medfilt: negative length");
    DDM_REQUIRE(kernel_size >= 1 && (kernel_size & 1) == 1, "ddm_medfilt: Each element of kernel_size should be odd.");
    DDM_REQUIRE(kernel_size <= kMedMaxK, "ddm_medfilt: kernel sizes above %d are not supported", kMedMaxK);
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_medfilt: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_medfilt");
    DeviceGuard guard(device);
    medfilt_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float *>(x_dev), n, kernel_size, static_cast<float *>(out_dev));
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_row_medians(int device, const void *x_dev, const int64_t *row_start_dev, int64_t rows, int64_t row_stride,
                    int64_t row_len, void *out_f64_dev, void *stream) {
    DDM_REQUIRE(rows >= 0 && row_len >= 1, "ddm_row_medians: bad sizes");
    if (rows == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_f64_dev != nullptr, "ddm_row_medians: NULL buffer");
    DDM_CHECK_DEVICE(device, "ddm_row_medians");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long *start = reinterpret_cast<const long long *>(row_start_dev);
    if (row_len <= kRowMedSmall) {
        row_median_small_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, 0, st>>>(
            static_cast<const float *>(x_dev), start, rows, row_stride, static_cast<int>(row_len),
            static_cast<double *>(out_f64_dev));
    } else {
        DDM_REQUIRE(rows <= 2147483647LL, "ddm_row_medians: too many rows");
        row_median_select_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(
            static_cast<const float *>(x_dev), start, row_stride, row_len, static_cast<double *>(out_f64_dev));
    }
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

int ddm_cu8_to_cf32(int device, const void *iq_u8_dev, int64_t n, void *out_dev, void *stream) {
    DDM_REQUIRE(n >= 0, "ddm_cu8_to_cf32: negative length");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(iq_u8_dev != nullptr && out_dev != nullptr, "ddm_cu8_to_cf32: NULL argument");
    DDM_CHECK_DEVICE(device, "ddm_cu8_to_cf32");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(iq_u8_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15) == 0;
    if (aligned)
        cu8_kernel<<<ops_grid(device, (n + 7) / 8), kOpsThreads, 0, st>>>(
            static_cast<const unsigned char *>(iq_u8_dev), static_cast<float2 *>(out_dev), n);
    else
        cu8_kernel_scalar<<<ops_grid(device, n), kOpsThreads, 0, st>>>(
            static_cast<const unsigned char *>(iq_u8_dev), static_cast<float2 *>(out_dev), n);
    DDM_CUDA(cudaGetLastError());
    count_launch();
    return DDM_OK;
}

}  // extern "C"
