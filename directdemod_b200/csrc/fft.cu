// FFT-defined operators of the hot path, float64 on the device:
//
//   ddm_am_hilbert   abs(scipy.signal.hilbert(x)) per chunk    demod_am.py:18-29,
//                    chunked like decode_noaa.__getAM          decode_noaa.py:631-657
//   ddm_resample     scipy.signal.resample(x, num)             comm.py:110-116 (strict bwLim),
//                                                              decode_noaa.py:350-351
//
// Both are *circular*, length-dependent operations on lengths that are not powers of two.
// 2-3-5 smooth lengths (240 000 = 2^7 3 5^4, the AM chunk of decode_noaa.py:647) are transformed
// directly by batched mixed-radix Stockham passes (radix 16/8/4/2, 25/5, 3 butterflies in
// registers, float64, twiddles from a float64 table); any other length (resample: 588 235 =
// 5 71 1657) is a Bluestein chirp-z over a power-of-two engine of the same passes (chirp phases
// k^2 mod 2n in exact integer arithmetic).  Independent chunks are the batch dimension: one
// launch per pass covers every chunk of a pass over a whole capture.
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

#include "ddm_common.cuh"

namespace ddm {

constexpr int kFftThreads = 256;

__device__ __forceinline__ double2 cmuld(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 conjd(double2 a) { return make_double2(a.x, -a.y); }

// tw[t] = exp(-2 pi i t / m)
__global__ void fft_twiddle_kernel(double2 *tw, int m) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    double s, c;
    sincospi(-2.0 * static_cast<double>(t) / static_cast<double>(m), &s, &c);
    tw[t] = make_double2(c, s);
}

// chirp w[k] = exp(-i pi k^2 / n), k < n;  b (length m) = conj chirp wrapped around
__global__ void fft_chirp_kernel(double2 *w, double2 *b, long long n, int m) {
    const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (k >= m) return;
    if (k < n) {
        const unsigned long long r = (static_cast<unsigned long long>(k) * static_cast<unsigned long long>(k)) %
                                     static_cast<unsigned long long>(2 * n);
        double s, c;
        sincospi(-static_cast<double>(r) / static_cast<double>(n), &s, &c);
        w[k] = make_double2(c, s);
        b[k] = make_double2(c, -s);
        if (k > 0) b[m - k] = make_double2(c, -s);
    } else if (k <= m - n) {
        b[k] = make_double2(0.0, 0.0);
    }
}

// ---- small DFTs in registers -------------------------------------------------------------
// DIR = +1: forward (exp(-i...)), DIR = -1: inverse (exp(+i...)), unnormalised.
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <int DIR>
__device__ __forceinline__ double2 mul_mi(double2 a) {
    return DIR > 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}
// w = (c, s) with s the FORWARD sine; conjugated for the inverse
template <int DIR>
__device__ __forceinline__ double2 cmul_w(double2 a, double c, double s) {
    const double ss = DIR > 0 ? s : -s;
    return make_double2(fma(a.x, c, -a.y * ss), fma(a.x, ss, a.y * c));
}

template <int R, int DIR>
struct SmallDft;

template <int DIR>
struct SmallDft<2, DIR> {
    __device__ static __forceinline__ void run(double2 (&v)[2]) {
        const double2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int DIR>
struct SmallDft<3, DIR> {
    __device__ static __forceinline__ void run(double2 (&v)[3]) {
        const double s3 = 0.86602540378443864676;    // sin(2 pi / 3)
        const double2 t = cadd(v[1], v[2]);
        const double2 d = mul_mi<DIR>(csub(v[1], v[2]));       // -i (v1 - v2) forward
        const double2 m = make_double2(v[0].x - 0.5 * t.x, v[0].y - 0.5 * t.y);
        v[0] = cadd(v[0], t);
        v[1] = make_double2(m.x + s3 * d.x, m.y + s3 * d.y);
        v[2] = make_double2(m.x - s3 * d.x, m.y - s3 * d.y);
    }
};

template <int DIR>
struct SmallDft<4, DIR> {
    __device__ static __forceinline__ void run(double2 (&v)[4]) {
        const double2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
        const double2 a2 = cadd(v[1], v[3]), a3 = mul_mi<DIR>(csub(v[1], v[3]));
        v[0] = cadd(a0, a2);
        v[1] = cadd(a1, a3);
        v[2] = csub(a0, a2);
        v[3] = csub(a1, a3);
    }
};

template <int DIR>
struct SmallDft<5, DIR> {
    __device__ static __forceinline__ void run(double2 (&v)[5]) {
        const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;   // cos(2pi/5), cos(4pi/5)
        const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;    // sin(2pi/5), sin(4pi/5)
        const double2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]);
        const double2 d1 = mul_mi<DIR>(csub(v[1], v[4])), d2 = mul_mi<DIR>(csub(v[2], v[3]));
        const double2 x0 = v[0];
        v[0] = make_double2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
        const double2 m1 = make_double2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
        const double2 m2 = make_double2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
        const double2 n1 = make_double2(s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y);
        const double2 n2 = make_double2(s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y);
        v[1] = cadd(m1, n1);
        v[4] = csub(m1, n1);
        v[2] = cadd(m2, n2);
        v[3] = csub(m2, n2);
    }
};

// composite sizes by one Cooley-Tukey step in registers: R = R1 * R2, input n = R2 n1 + n2,
// output k = k1 + R1 k2; inner twiddles exp(-2 pi i n2 k1 / R) from a constant table
__constant__ double2 c_tw8[8], c_tw16[16], c_tw25[25];

template <int R>
__device__ __forceinline__ double2 small_tw(int idx);
template <>
__device__ __forceinline__ double2 small_tw<8>(int idx) { return c_tw8[idx]; }
template <>
__device__ __forceinline__ double2 small_tw<16>(int idx) { return c_tw16[idx]; }
template <>
__device__ __forceinline__ double2 small_tw<25>(int idx) { return c_tw25[idx]; }

template <int R, int R1, int R2, int DIR>
__device__ __forceinline__ void composite_dft(double2 (&v)[R]) {
    double2 a[R1][R2];
#pragma unroll
    for (int n2 = 0; n2 < R2; ++n2) {
        double2 col[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) col[n1] = v[R2 * n1 + n2];
        SmallDft<R1, DIR>::run(col);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            double2 x = col[k1];
            if (n2 * k1 != 0) {
                const double2 w = small_tw<R>((n2 * k1) % R);
                x = cmul_w<DIR>(x, w.x, w.y);
            }
            a[k1][n2] = x;
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
        double2 row[R2];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) row[n2] = a[k1][n2];
        SmallDft<R2, DIR>::run(row);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = row[k2];
    }
}

template <int DIR>
struct SmallDft<8, DIR> {
    __device__ static __forceinline__ void run(double2 (&v)[8]) { composite_dft<8, 2, 4, DIR>(v); }
};
template <int DIR>
struct SmallDft<16, DIR> {
    __device__ static __forceinline__ void run(double2 (&v)[16]) { composite_dft<16, 4, 4, DIR>(v); }
};
template <int DIR>
struct SmallDft<25, DIR> {
    __device__ static __forceinline__ void run(double2 (&v)[25]) { composite_dft<25, 5, 5, DIR>(v); }
};

// one Stockham radix-R pass over `batch` rows of length m (grid.y = row):
//   v[r] = in[j + r m/R] * w^(r k),  k = j mod Ns,  w = exp(-+2 pi i / (Ns R));  DFT_R;
//   out[(j - k) R + k + r Ns] = v[r]
//
// IO selects what the first / last pass of a transform fuses (the Hilbert envelope of real chunks,
// two chunks per complex transform -- see ddm_am_hilbert):
//   kIoPlain     double2 in, double2 out
//   kIoLoadPair  first pass: row p reads the real chunks 2p (-> re) and 2p+1 (-> im) as f32
//   kIoHilbert   last forward pass: bin k is multiplied by -i sgn(k) / n (0 for k = 0 and the Nyquist bin)
//   kIoAbsPair   last inverse pass: out_a = hypot(x_a, re), out_b = hypot(x_b, im) as f32
constexpr int kIoPlain = 0, kIoLoadPair = 1, kIoHilbert = 2, kIoAbsPair = 3;
struct PassIo {
    const float *x;         // the real chunks, chunk c at x + c * len
    float *env;             // the envelopes, same layout
    long long len;          // chunk length (= m)
    long long chunks;       // number of chunks (the last pair may lack its second one)
    double scale;
};

template <int R, int DIR, int IO = kIoPlain>
__global__ void __launch_bounds__(R >= 16 ? 128 : 256)
fft_pass(const double2 *__restrict__ in, double2 *__restrict__ out, const double2 *__restrict__ tw, int m,
         int Ns, const PassIo io = PassIo{}) {
    const int q = m / R;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= q) return;
    const size_t row = static_cast<size_t>(blockIdx.y) * m;
    in += row;
    out += row;
    const int k = j % Ns;
    const int step = q / Ns;                 // m / (R Ns)
    double2 v[R];
    const long long ca = 2LL * blockIdx.y, cb = ca + 1;          // the pair's chunks (IO modes)
    if (IO == kIoLoadPair) {
        const float *xa = io.x + ca * io.len;
        const float *xb = io.x + cb * io.len;
        const bool has_b = cb < io.chunks;
#pragma unroll
        for (int r = 0; r < R; ++r)
            v[r] = make_double2(static_cast<double>(xa[j + r * q]), has_b ? static_cast<double>(xb[j + r * q]) : 0.0);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[j + r * q];
    }
    if (k != 0) {
#pragma unroll
        for (int r = 1; r < R; ++r) {
            const double2 t = tw[static_cast<size_t>(k) * step * r];
            v[r] = cmul_w<DIR>(v[r], t.x, t.y);
        }
    }
    SmallDft<R, DIR>::run(v);
    const int j0 = (j - k) * R + k;
    if (IO == kIoHilbert) {
        // -i sgn(k) X[k] / n: the spectrum of the Hilbert transforms of both packed chunks at once
        // (H is a real operator: ifft gives H{x_a} + i H{x_b}); scipy.signal.hilbert's mask minus one
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int o = j0 + r * Ns;
            double2 w = make_double2(0.0, 0.0);
            if (o != 0 && 2 * o != m) {
                const double sc = io.scale;
                w = 2 * o < m ? make_double2(v[r].y * sc, -v[r].x * sc) : make_double2(-v[r].y * sc, v[r].x * sc);
            }
            out[o] = w;
        }
    } else if (IO == kIoAbsPair) {
        const float *xa = io.x + ca * io.len;
        const float *xb = io.x + cb * io.len;
        float *ea = io.env + ca * io.len;
        float *eb = io.env + cb * io.len;
        const bool has_b = cb < io.chunks;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int o = j0 + r * Ns;
            ea[o] = static_cast<float>(hypot(static_cast<double>(xa[o]), v[r].x));
            if (has_b) eb[o] = static_cast<float>(hypot(static_cast<double>(xb[o]), v[r].y));
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) out[j0 + r * Ns] = v[r];
    }
}

// buf[row][k] *= bfft[k]
__global__ void fft_pointwise_kernel(double2 *buf, const double2 *__restrict__ bfft, int m) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    double2 *p = buf + static_cast<size_t>(blockIdx.y) * m + k;
    *p = cmuld(*p, bfft[k]);
}

// ---- Bluestein stages ----------------------------------------------------------------
// load:  a[row][k] = conj?(src[row][k]) * w[k]  (k < n), 0 (n <= k < m)
// src kinds: 0 = f32 real rows, 1 = cf32 rows, 2 = double2 rows (row stride src_stride elements)
template <int KIND, bool CONJ>
__global__ void blu_load_kernel(const void *__restrict__ src, long long src_stride, long long n,
                                const double2 *__restrict__ w, double2 *__restrict__ a, int m,
                                const long long *__restrict__ row_off) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const size_t row = blockIdx.y;
    const size_t r0 = row_off ? static_cast<size_t>(row_off[row]) : row * src_stride;
    double2 v = make_double2(0.0, 0.0);
    if (k < n) {
        if (KIND == 0) v.x = static_cast<double>(static_cast<const float *>(src)[r0 + k]);
        else if (KIND == 1) {
            const float2 s = static_cast<const float2 *>(src)[r0 + k];
            v = make_double2(static_cast<double>(s.x), static_cast<double>(s.y));
        } else v = static_cast<const double2 *>(src)[r0 + k];
        if (CONJ) v.y = -v.y;
        v = cmuld(v, w[k]);
    }
    a[row * m + k] = v;
}

// finish: X[row][k] = w[k] * c[row][k] / m   (k < n) -> dst (double2 rows of stride dst_stride),
// optionally conjugated and scaled (the inverse transform is conj(DFT(conj(.))) / n)
template <bool CONJ>
__global__ void blu_finish_kernel(const double2 *__restrict__ c, int m, long long n,
                                  const double2 *__restrict__ w, double scale, double2 *__restrict__ dst,
                                  long long dst_stride) {
    const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const size_t row = blockIdx.y;
    double2 v = c[row * m + k];
    if (w != nullptr) v = cmuld(v, w[k]);
    v.x *= scale;
    v.y *= scale;
    if (CONJ) v.y = -v.y;
    dst[row * dst_stride + k] = v;
}

// power-of-two direct path helpers
template <int KIND>
__global__ void fft_load_plain_kernel(const void *__restrict__ src, long long src_stride, long long n,
                                      double2 *__restrict__ a, const long long *__restrict__ row_off) {
    const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const size_t row = blockIdx.y;
    const size_t r0 = row_off ? static_cast<size_t>(row_off[row]) : row * src_stride;
    double2 v = make_double2(0.0, 0.0);
    if (KIND == 0) v.x = static_cast<double>(static_cast<const float *>(src)[r0 + k]);
    else if (KIND == 1) {
        const float2 s = static_cast<const float2 *>(src)[r0 + k];
        v = make_double2(static_cast<double>(s.x), static_cast<double>(s.y));
    } else v = static_cast<const double2 *>(src)[r0 + k];
    a[row * n + k] = v;
}

// ---- operator-specific spectrum edits --------------------------------------------------
// hilbert mask (scipy.signal.hilbert): h[0] = 1, h[1..(n-1)/2] = 2, h[n/2] = 1 (n even), else 0
__global__ void hilbert_mask_kernel(double2 *spec, long long n, long long stride) {
    const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    double2 *p = spec + static_cast<size_t>(blockIdx.y) * stride + k;
    double h;
    if (k == 0) h = 1.0;
    else if ((n & 1) == 0 && k == n / 2) h = 1.0;
    else if (k < (n + 1) / 2) h = 2.0;
    else h = 0.0;
    p->x *= h;
    p->y *= h;
}

__global__ void abs_out_kernel(const double2 *__restrict__ z, long long n, long long stride,
                               float *__restrict__ out, long long out_stride) {
    const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (k >= n) return;
    const double2 v = z[static_cast<size_t>(blockIdx.y) * stride + k];
    out[static_cast<size_t>(blockIdx.y) * out_stride + k] = static_cast<float>(hypot(v.x, v.y));
}

// scipy.signal.resample spectrum placement, X (length n) -> Y (length num)
//   real input:    Hermitian spectrum of irfft(rfft(x)[:N//2+1] ...) semantics
//   complex input: the fft/ifft branch
__global__ void resample_spec_kernel(const double2 *__restrict__ X, long long n, double2 *__restrict__ Y,
                                     long long num, int is_real) {
    const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (k >= num) return;
    X += static_cast<size_t>(blockIdx.y) * n;          // one spectrum per row
    Y += static_cast<size_t>(blockIdx.y) * num;
    const long long N = n < num ? n : num;
    const long long half = N / 2;
    double2 v = make_double2(0.0, 0.0);
    if (is_real) {
        // half spectrum Yh[0..num/2]; Yh[j] = X[j] for j < N/2+1; even N: Nyquist-of-N bin doubled
        // (down) or halved (up); irfft ignores the imaginary part of DC and of the num/2 bin
        const long long j = k <= num / 2 ? k : num - k;            // mirrored index
        if (j <= half) {
            v = X[j];
            if ((N & 1) == 0 && j == half) {
                if (num < n) { v.x *= 2.0; v.y *= 2.0; }
                else if (num > n) { v.x *= 0.5; v.y *= 0.5; }
            }
            if (j == 0 || ((num & 1) == 0 && j == num / 2)) v.y = 0.0;
            if (k > num / 2) v.y = -v.y;                              // Hermitian mirror
        }
    } else {
        const long long nyq = half + 1;
        if (k < nyq) {
            v = X[k];
        } else if (N > 2 && k >= num - (N - nyq)) {
            v = X[n - (num - k)];
        }
        if ((N & 1) == 0) {
            if (num < n) {
                // downsampling: Y[-N/2] += X[-N/2]
                if (k == num - half) {
                    const double2 e = X[n - half];
                    // when num - half == half (N == num) the slot already holds X[half]
                    v = make_double2(v.x + e.x, v.y + e.y);
                }
            } else if (num > n) {
                // upsampling: Y[N/2] *= 0.5 and mirrored into Y[num - N/2]
                if (k == half) { v.x *= 0.5; v.y *= 0.5; }
                else if (k == num - half) { const double2 e = X[half]; v = make_double2(0.5 * e.x, 0.5 * e.y); }
            }
        }
    }
    Y[k] = v;
}

__global__ void resample_out_kernel(const double2 *__restrict__ z, long long num, double scale, void *out,
                                    int is_real) {
    const long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (k >= num) return;
    const size_t r0 = static_cast<size_t>(blockIdx.y) * num;
    const double2 v = z[r0 + k];
    if (is_real) static_cast<float *>(out)[r0 + k] = static_cast<float>(v.x * scale);
    else static_cast<float2 *>(out)[r0 + k] = make_float2(static_cast<float>(v.x * scale), static_cast<float>(v.y * scale));
}

}  // namespace ddm

using namespace ddm;

// =====================================================================================
// host side: plans and context
// =====================================================================================
struct FftPlan {
    long long n = 0;
    int m = 0;              // transform size of the power-of-two engine
    bool pow2 = false;      // n itself is 2-3-5 smooth: transformed directly, no chirp
    std::vector<int> radices;   // Stockham pass radices of the size-m engine
    double2 *d_tw = nullptr;
    double2 *d_w = nullptr;
    double2 *d_bfft = nullptr;
};

struct ddm_fft {
    int device = 0;
    std::map<long long, FftPlan> plans;
    double2 *d_ws[2] = {nullptr, nullptr};
    size_t ws_elems = 0;
    double2 *d_spec[2] = {nullptr, nullptr};   // length-n spectra (rows)
    size_t spec_elems = 0;
};

namespace {

int ilog2(long long v) {
    int l = 0;
    while ((1LL << l) < v) ++l;
    return l;
}

// radices of a 2-3-5 smooth length (largest butterflies first); empty if n is not smooth
std::vector<int> smooth_radices(long long n) {
    std::vector<int> r;
    int e2 = 0, e3 = 0, e5 = 0;
    while (n % 2 == 0) { n /= 2; ++e2; }
    while (n % 3 == 0) { n /= 3; ++e3; }
    while (n % 5 == 0) { n /= 5; ++e5; }
    if (n != 1) return r;
    while (e2 >= 4) { r.push_back(16); e2 -= 4; }
    if (e2 == 3) r.push_back(8);
    if (e2 == 2) r.push_back(4);
    if (e2 == 1) r.push_back(2);
    while (e5 >= 2) { r.push_back(25); e5 -= 2; }
    if (e5 == 1) r.push_back(5);
    while (e3 >= 1) { r.push_back(3); --e3; }
    return r;
}

template <int R, int DIR, int IO = kIoPlain>
void launch_pass(double2 *in, double2 *out, const double2 *tw, int m, int Ns, int batch, cudaStream_t st,
                 const PassIo &io = PassIo{}) {
    const int threads = R >= 16 ? 128 : 256;
    const dim3 grid((m / R + threads - 1) / threads, batch);
    fft_pass<R, DIR, IO><<<grid, threads, 0, st>>>(in, out, tw, m, Ns, io);
    count_launch();
}

template <int DIR, int IO>
void launch_pass_r(int R, double2 *in, double2 *out, const double2 *tw, int m, int Ns, int batch, cudaStream_t st,
                   const PassIo &io) {
    switch (R) {
        case 2: launch_pass<2, DIR, IO>(in, out, tw, m, Ns, batch, st, io); break;
        case 3: launch_pass<3, DIR, IO>(in, out, tw, m, Ns, batch, st, io); break;
        case 4: launch_pass<4, DIR, IO>(in, out, tw, m, Ns, batch, st, io); break;
        case 5: launch_pass<5, DIR, IO>(in, out, tw, m, Ns, batch, st, io); break;
        case 8: launch_pass<8, DIR, IO>(in, out, tw, m, Ns, batch, st, io); break;
        case 16: launch_pass<16, DIR, IO>(in, out, tw, m, Ns, batch, st, io); break;
        case 25: launch_pass<25, DIR, IO>(in, out, tw, m, Ns, batch, st, io); break;
    }
}

// batched FFT of a 2-3-5 smooth size m, rows in bufs[cur]; returns the index of the buffer holding
// the result
template <int DIR>
int pow2_fft(double2 *bufs[2], int cur, const double2 *tw, int m, int batch, cudaStream_t st,
             const std::vector<int> &radices) {
    int Ns = 1;
    for (int R : radices) {
        double2 *in = bufs[cur], *out = bufs[cur ^ 1];
        switch (R) {
            case 2: launch_pass<2, DIR>(in, out, tw, m, Ns, batch, st); break;
            case 3: launch_pass<3, DIR>(in, out, tw, m, Ns, batch, st); break;
            case 4: launch_pass<4, DIR>(in, out, tw, m, Ns, batch, st); break;
            case 5: launch_pass<5, DIR>(in, out, tw, m, Ns, batch, st); break;
            case 8: launch_pass<8, DIR>(in, out, tw, m, Ns, batch, st); break;
            case 16: launch_pass<16, DIR>(in, out, tw, m, Ns, batch, st); break;
            case 25: launch_pass<25, DIR>(in, out, tw, m, Ns, batch, st); break;
        }
        cur ^= 1;
        Ns *= R;
    }
    return cur;
}

int upload_small_twiddles() {
    static bool done[64] = {false};
    int dev = 0;
    DDM_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return DDM_OK;
    double2 t8[8], t16[16], t25[25];
    for (int i = 0; i < 8; ++i) t8[i] = make_double2(std::cos(2.0 * M_PI * i / 8), -std::sin(2.0 * M_PI * i / 8));
    for (int i = 0; i < 16; ++i) t16[i] = make_double2(std::cos(2.0 * M_PI * i / 16), -std::sin(2.0 * M_PI * i / 16));
    for (int i = 0; i < 25; ++i) t25[i] = make_double2(std::cos(2.0 * M_PI * i / 25), -std::sin(2.0 * M_PI * i / 25));
    DDM_CUDA(cudaMemcpyToSymbol(c_tw8, t8, sizeof(t8)));
    DDM_CUDA(cudaMemcpyToSymbol(c_tw16, t16, sizeof(t16)));
    DDM_CUDA(cudaMemcpyToSymbol(c_tw25, t25, sizeof(t25)));
    if (dev < 64) done[dev] = true;
    return DDM_OK;
}

int ensure_ws(ddm_fft *c, size_t elems, cudaStream_t st) {
    if (elems <= c->ws_elems) return DDM_OK;
    DDM_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_ws[i]);
        c->d_ws[i] = nullptr;
    }
    c->ws_elems = 0;
    for (int i = 0; i < 2; ++i) DDM_CUDA(cudaMalloc(&c->d_ws[i], sizeof(double2) * elems));
    c->ws_elems = elems;
    return DDM_OK;
}

int ensure_spec(ddm_fft *c, size_t elems, cudaStream_t st) {
    if (elems <= c->spec_elems) return DDM_OK;
    DDM_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_spec[i]);
        c->d_spec[i] = nullptr;
    }
    c->spec_elems = 0;
    for (int i = 0; i < 2; ++i) DDM_CUDA(cudaMalloc(&c->d_spec[i], sizeof(double2) * elems));
    c->spec_elems = elems;
    return DDM_OK;
}

int get_plan(ddm_fft *c, long long n, cudaStream_t st, FftPlan **out) {
    auto it = c->plans.find(n);
    if (it != c->plans.end()) {
        *out = &it->second;
        return DDM_OK;
    }
    if (n > (1LL << 29)) {
        set_error("FFT length %lld is above the supported maximum 2^29", n);
        return DDM_ERR_UNSUPPORTED;
    }
    FftPlan p;
    p.n = n;
    int rc0 = upload_small_twiddles();
    if (rc0 != DDM_OK) return rc0;
    p.radices = smooth_radices(n);
    p.pow2 = !p.radices.empty() || n == 1;
    p.m = p.pow2 ? static_cast<int>(n) : (1 << ilog2(2 * n - 1));
    if (!p.pow2) p.radices = smooth_radices(p.m);
    DDM_CUDA(cudaMalloc(&p.d_tw, sizeof(double2) * p.m));
    fft_twiddle_kernel<<<(p.m + 255) / 256, 256, 0, st>>>(p.d_tw, p.m);
    count_launch();
    if (!p.pow2) {
        DDM_CUDA(cudaMalloc(&p.d_w, sizeof(double2) * n));
        DDM_CUDA(cudaMalloc(&p.d_bfft, sizeof(double2) * p.m));
        int rc = ensure_ws(c, static_cast<size_t>(p.m), st);
        if (rc != DDM_OK) return rc;
        fft_chirp_kernel<<<(p.m + 255) / 256, 256, 0, st>>>(p.d_w, c->d_ws[0], n, p.m);
        count_launch();
        const int r = pow2_fft<1>(c->d_ws, 0, p.d_tw, p.m, 1, st, p.radices);
        DDM_CUDA(cudaMemcpyAsync(p.d_bfft, c->d_ws[r], sizeof(double2) * p.m, cudaMemcpyDeviceToDevice, st));
    }
    DDM_CUDA(cudaGetLastError());
    auto ins = c->plans.emplace(n, p);
    *out = &ins.first->second;
    return DDM_OK;
}

// DFT of `batch` rows.  src: rows of kind KIND (0 f32, 1 cf32, 2 double2) with row stride
// src_stride; dst: double2 rows of stride dst_stride.  INVERSE computes the unnormalised-by-caller
// inverse as conj(DFT(conj(x))) * scale.
template <int KIND, bool INVERSE>
int dft_rows(ddm_fft *c, FftPlan *p, const void *src, long long src_stride, int batch, double scale,
             double2 *dst, long long dst_stride, cudaStream_t st, const long long *row_off = nullptr) {
    const long long n = p->n;
    const int m = p->m;
    int rc = ensure_ws(c, static_cast<size_t>(m) * batch, st);
    if (rc != DDM_OK) return rc;
    if (p->pow2) {
        const dim3 g((n + kFftThreads - 1) / kFftThreads, batch);
        fft_load_plain_kernel<KIND><<<g, kFftThreads, 0, st>>>(src, src_stride, n, c->d_ws[0], row_off);
        count_launch();
        const int r = INVERSE ? pow2_fft<-1>(c->d_ws, 0, p->d_tw, m, batch, st, p->radices)
                              : pow2_fft<1>(c->d_ws, 0, p->d_tw, m, batch, st, p->radices);
        // copy out with scale (no chirp on the power-of-two path)
        blu_finish_kernel<false><<<g, kFftThreads, 0, st>>>(c->d_ws[r], m, n, nullptr, scale, dst, dst_stride);
        count_launch();
        DDM_CUDA(cudaGetLastError());
        return DDM_OK;
    }
    const dim3 gm((m + kFftThreads - 1) / kFftThreads, batch);
    const dim3 gn((n + kFftThreads - 1) / kFftThreads, batch);
    blu_load_kernel<KIND, INVERSE><<<gm, kFftThreads, 0, st>>>(src, src_stride, n, p->d_w, c->d_ws[0], m, row_off);
    count_launch();
    int r = pow2_fft<1>(c->d_ws, 0, p->d_tw, m, batch, st, p->radices);
    fft_pointwise_kernel<<<gm, kFftThreads, 0, st>>>(c->d_ws[r], p->d_bfft, m);
    count_launch();
    r = pow2_fft<-1>(c->d_ws, r, p->d_tw, m, batch, st, p->radices);
    blu_finish_kernel<INVERSE><<<gn, kFftThreads, 0, st>>>(c->d_ws[r], m, n, p->d_w,
                                                           scale / static_cast<double>(m), dst, dst_stride);
    count_launch();
    DDM_CUDA(cudaGetLastError());
    return DDM_OK;
}

// Hilbert envelope of `chunks` real chunks of a 2-3-5 smooth length, two chunks per complex transform,
// load / spectrum edit / envelope fused into the first and last passes: 2 x passes launches, each one
// read + one write of the packed rows.  (A length with a single radix would need both fusions in one
// pass; such chunks take the general route.)
int hilbert_pairs(ddm_fft *c, FftPlan *p, const float *x, float *env, long long len, long long chunks,
                  cudaStream_t st) {
    const int m = p->m;
    const long long pairs_total = (chunks + 1) / 2;
    long long rows_per = std::max<long long>(1, (64LL << 20) / (2 * static_cast<long long>(m)));
    rows_per = std::min<long long>(rows_per, 65535);
    const int np = static_cast<int>(p->radices.size());
    for (long long p0 = 0; p0 < pairs_total; p0 += rows_per) {
        const int batch = static_cast<int>(std::min<long long>(rows_per, pairs_total - p0));
        int rc = ensure_ws(c, static_cast<size_t>(m) * batch, st);
        if (rc != DDM_OK) return rc;
        PassIo io;
        io.x = x + 2 * p0 * len;
        io.env = env + 2 * p0 * len;
        io.len = len;
        io.chunks = chunks - 2 * p0;
        io.scale = 1.0 / static_cast<double>(len);
        int cur = 0, Ns = 1;
        for (int i = 0; i < np; ++i) {                       // forward
            const int R = p->radices[i];
            double2 *in = c->d_ws[cur], *out = c->d_ws[cur ^ 1];
            if (i == 0) launch_pass_r<1, kIoLoadPair>(R, in, out, p->d_tw, m, Ns, batch, st, io);
            else if (i == np - 1) launch_pass_r<1, kIoHilbert>(R, in, out, p->d_tw, m, Ns, batch, st, io);
            else launch_pass_r<1, kIoPlain>(R, in, out, p->d_tw, m, Ns, batch, st, io);
            cur ^= 1;
            Ns *= R;
        }
        Ns = 1;
        for (int i = 0; i < np; ++i) {                       // inverse
            const int R = p->radices[i];
            double2 *in = c->d_ws[cur], *out = c->d_ws[cur ^ 1];
            if (i == np - 1) launch_pass_r<-1, kIoAbsPair>(R, in, out, p->d_tw, m, Ns, batch, st, io);
            else launch_pass_r<-1, kIoPlain>(R, in, out, p->d_tw, m, Ns, batch, st, io);
            cur ^= 1;
            Ns *= R;
        }
    }
    DDM_CUDA(cudaGetLastError());
    return DDM_OK;
}

}  // namespace

extern "C" {

int ddm_fft_create(int device, ddm_fft **out) {
    DDM_REQUIRE(out != nullptr, "ddm_fft_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    DDM_CUDA(cudaGetDeviceCount(&ndev));
    DDM_REQUIRE(device >= 0 && device < ndev, "ddm_fft_create: no such device %d", device);
    ddm_fft *c = new (std::nothrow) ddm_fft();
    if (!c) {
        set_error("ddm_fft_create: out of host memory");
        return DDM_ERR_NOMEM;
    }
    c->device = device;
    *out = c;
    return DDM_OK;
}

int ddm_fft_destroy(ddm_fft *c) {
    if (!c) return DDM_OK;
    DeviceGuard guard(c->device);
    for (auto &kv : c->plans) {
        cudaFree(kv.second.d_tw);
        cudaFree(kv.second.d_w);
        cudaFree(kv.second.d_bfft);
    }
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->d_ws[i]);
        cudaFree(c->d_spec[i]);
    }
    delete c;
    return DDM_OK;
}

int ddm_am_hilbert(ddm_fft *c, const void *x_dev, int64_t n, int64_t chunk, void *out_dev, void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_am_hilbert: NULL context");
    DDM_REQUIRE(n >= 0 && chunk >= 1, "ddm_am_hilbert: bad length/chunk");
    if (n == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_am_hilbert: NULL buffer");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float *x = static_cast<const float *>(x_dev);
    float *out = static_cast<float *>(out_dev);
    // chunk bounds like chunker.py:32-45: full chunks while start + chunk < n, then the rest
    int64_t full = 0;
    while ((full + 1) * chunk < n) ++full;
    const int64_t tail = n - full * chunk;          // 1..chunk samples
    struct Part { int64_t start, len, rows; };
    Part parts[2] = {{0, chunk, full}, {full * chunk, tail, 1}};
    if (tail == chunk) {                            // the last chunk is a full one too
        parts[0].rows = full + 1;
        parts[1].rows = 0;
    }
    for (const Part &pt : parts) {
        if (pt.rows == 0) continue;
        FftPlan *p = nullptr;
        int rc = get_plan(c, pt.len, st, &p);
        if (rc != DDM_OK) return rc;
        if (p->pow2 && p->radices.size() >= 2) {
            rc = hilbert_pairs(c, p, x + pt.start, out + pt.start, pt.len, pt.rows, st);
            if (rc != DDM_OK) return rc;
            continue;
        }
        // sub-batches keep the workspace below ~1 GiB
        int64_t rows_per = std::max<int64_t>(1, (64LL << 20) / (2 * static_cast<int64_t>(p->m)));
        rows_per = std::min<int64_t>(rows_per, 65535);
        for (int64_t r0 = 0; r0 < pt.rows; r0 += rows_per) {
            const int batch = static_cast<int>(std::min<int64_t>(rows_per, pt.rows - r0));
            rc = ensure_spec(c, static_cast<size_t>(pt.len) * batch, st);
            if (rc != DDM_OK) return rc;
            const float *src = x + pt.start + r0 * pt.len;
            rc = dft_rows<0, false>(c, p, src, pt.len, batch, 1.0, c->d_spec[0], pt.len, st);
            if (rc != DDM_OK) return rc;
            const dim3 gn((pt.len + kFftThreads - 1) / kFftThreads, batch);
            hilbert_mask_kernel<<<gn, kFftThreads, 0, st>>>(c->d_spec[0], pt.len, pt.len);
            count_launch();
            rc = dft_rows<2, true>(c, p, c->d_spec[0], pt.len, batch, 1.0 / static_cast<double>(pt.len),
                                   c->d_spec[1], pt.len, st);
            if (rc != DDM_OK) return rc;
            abs_out_kernel<<<gn, kFftThreads, 0, st>>>(c->d_spec[1], pt.len, pt.len,
                                                       out + pt.start + r0 * pt.len, pt.len);
            count_launch();
        }
    }
    DDM_CUDA(cudaGetLastError());
    return DDM_OK;
}

int ddm_resample(ddm_fft *c, const void *x_dev, int64_t n, int is_complex, int64_t num, void *out_dev,
                 void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_resample: NULL context");
    DDM_REQUIRE(n >= 1 && num >= 0, "ddm_resample: bad lengths");
    if (num == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && out_dev != nullptr, "ddm_resample: NULL buffer");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FftPlan *pn = nullptr, *pm = nullptr;
    int rc = get_plan(c, n, st, &pn);
    if (rc != DDM_OK) return rc;
    rc = get_plan(c, num, st, &pm);
    if (rc != DDM_OK) return rc;
    // std::map nodes are stable, pn stays valid after the second insertion
    rc = ensure_spec(c, static_cast<size_t>(std::max<int64_t>(n, num)), st);
    if (rc != DDM_OK) return rc;
    if (is_complex) rc = dft_rows<1, false>(c, pn, x_dev, n, 1, 1.0, c->d_spec[0], n, st);
    else rc = dft_rows<0, false>(c, pn, x_dev, n, 1, 1.0, c->d_spec[0], n, st);
    if (rc != DDM_OK) return rc;
    const unsigned g = static_cast<unsigned>((num + kFftThreads - 1) / kFftThreads);
    resample_spec_kernel<<<g, kFftThreads, 0, st>>>(c->d_spec[0], n, c->d_spec[1], num, is_complex ? 0 : 1);
    count_launch();
    // y = ifft(Y) * num / n = conj(DFT(conj Y)) / n
    rc = dft_rows<2, true>(c, pm, c->d_spec[1], num, 1, 1.0 / static_cast<double>(n), c->d_spec[0], num, st);
    if (rc != DDM_OK) return rc;
    resample_out_kernel<<<g, kFftThreads, 0, st>>>(c->d_spec[0], num, 1.0, out_dev, is_complex ? 0 : 1);
    count_launch();
    DDM_CUDA(cudaGetLastError());
    return DDM_OK;
}

int ddm_resample_rows(ddm_fft *c, const void *x_dev, const int64_t *row_start_dev, int64_t rows, int64_t n,
                      int64_t num, void *out_dev, void *stream) {
    DDM_REQUIRE(c != nullptr, "ddm_resample_rows: NULL context");
    DDM_REQUIRE(rows >= 0 && n >= 1 && num >= 0, "ddm_resample_rows: bad sizes");
    if (rows == 0 || num == 0) return DDM_OK;
    DDM_REQUIRE(x_dev != nullptr && row_start_dev != nullptr && out_dev != nullptr, "ddm_resample_rows: NULL buffer");
    DeviceGuard guard(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FftPlan *pn = nullptr, *pm = nullptr;
    int rc = get_plan(c, n, st, &pn);
    if (rc != DDM_OK) return rc;
    rc = get_plan(c, num, st, &pm);
    if (rc != DDM_OK) return rc;
    const int64_t mmax = std::max<int64_t>(pn->m, pm->m);
    int64_t rows_per = std::max<int64_t>(1, (32LL << 20) / mmax);       // workspace <= ~1 GiB
    rows_per = std::min<int64_t>(rows_per, 65535);
    const int64_t big = std::max<int64_t>(n, num);
    float *out = static_cast<float *>(out_dev);
    for (int64_t r0 = 0; r0 < rows; r0 += rows_per) {
        const int batch = static_cast<int>(std::min<int64_t>(rows_per, rows - r0));
        rc = ensure_spec(c, static_cast<size_t>(big) * batch, st);
        if (rc != DDM_OK) return rc;
        rc = dft_rows<0, false>(c, pn, x_dev, 0, batch, 1.0, c->d_spec[0], n, st,
                                reinterpret_cast<const long long *>(row_start_dev + r0));
        if (rc != DDM_OK) return rc;
        const dim3 g(static_cast<unsigned>((num + kFftThreads - 1) / kFftThreads), batch);
        resample_spec_kernel<<<g, kFftThreads, 0, st>>>(c->d_spec[0], n, c->d_spec[1], num, 1);
        count_launch();
        rc = dft_rows<2, true>(c, pm, c->d_spec[1], num, batch, 1.0 / static_cast<double>(n), c->d_spec[0], num, st);
        if (rc != DDM_OK) return rc;
        resample_out_kernel<<<g, kFftThreads, 0, st>>>(c->d_spec[0], num, 1.0, out + r0 * num, 1);
        count_launch();
    }
    DDM_CUDA(cudaGetLastError());
    return DDM_OK;
}

}  // extern "C"
