/*
 * ddemod.h -- C ABI of libddemod.so, the B200 (sm_100a) implementation of DirectDemod's
 * IQ-stream demodulation hot path.
 *
 * The reference (aerospaceresearch/DirectDemod) is pure Python and has NO FFI of its own;
 * its "plugin interface" for this path is the duck-typed operator protocol of
 * directdemod/comm.py (commSignal.offsetFreq/.filter/.bwLim/.funcApply),
 * directdemod/filters.py (filter.applyOn) and directdemod/demod_fm.py / demod_am.py
 * (.demod).  Each entry point below names the reference interface (file:line, relative to
 * the reference checkout) whose arithmetic it replaces.  The Python host layer
 * (directdemod_b200/*.py) mirrors the reference classes one-to-one and binds these symbols
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative ddm_status on failure; the message is
 *     available from ddm_last_error() (thread local).  No exception crosses the ABI.
 *   - Plain pointers and sizes only.  "_dev" pointers are CUDA device pointers on the
 *     handle's device; "_host" pointers are ordinary host memory (pinned or pageable).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  All
 *     work is enqueued on it; "_host" entry points synchronise the stream before returning,
 *     "_dev" entry points do not.
 *   - A handle is bound to one device.  Calls on one handle must be serialised by the
 *     caller; different handles are independent.
 *   - cf32 = interleaved (re, im) float32 pairs; f32 = float32; f64 = float64.
 */
#ifndef DDEMOD_H
#define DDEMOD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ddm_status {
    DDM_OK = 0,
    DDM_ERR_INVALID = -1,      /* bad argument                                           */
    DDM_ERR_CUDA = -2,         /* a CUDA runtime call failed                             */
    DDM_ERR_NOMEM = -3,        /* allocation failed                                      */
    DDM_ERR_CAPACITY = -4,     /* output buffer too small                                */
    DDM_ERR_UNSUPPORTED = -5   /* configuration not implemented                          */
} ddm_status;

/* ---- library ------------------------------------------------------------------------ */
int ddm_version(void);
const char *ddm_last_error(void);
/* number of kernels this library has launched in the calling process (all handles) */
int64_t ddm_launch_count(void);
int ddm_device_count(int *count);
/* free the per-device scratch pool kept by ddm_correlate / ddm_topk_sums / ddm_compact_above / ddm_bank4 */
int ddm_release_scratch(int device);

/* ---- fused chain ---------------------------------------------------------------------
 * offsetFreq -> real-tap FIR with carried state -> integer decimation -> FM discriminator
 * in ONE kernel launch per chunk.  Replaces, for one chunk of a chunked stream,
 *     comm.py:63-78   commSignal.offsetFreq   (mixer, global sample index carried)
 *     filters.py:64-70 filter.applyOn stateful (lfilter(b,[1],x,zi), zi0 = lfilter_zi)
 *     comm.py:118-129 commSignal.bwLim non-strict (x[off::j], carried offset)
 *     demod_fm.py:40-49 demod_fm.demod        (angle(x[n] conj x[n-1]), carried sample)
 * i.e. the body of decode_noaa.py:623 / decode_fm.py:64-68 / decode_afsk1200.py:79-91.
 *
 * State carried between calls (what the reference keeps in chunker vars, filter.__zi and
 * demod_fm.__last): the global sample index n0, the decimation offset, whether a previous
 * decimated sample exists, and a raw-input halo of ddm_chain_halo_len() samples from which
 * the FIR state and the previous decimated sample are recomputed.
 */
typedef struct ddm_chain ddm_chain;

#define DDM_CHAIN_OUT_FM 0   /* f32 output: FM discriminator                             */
#define DDM_CHAIN_OUT_IQ 1   /* cf32 output: filtered + decimated IQ (no discriminator)  */

/* input sample formats */
#define DDM_IN_CF32 0        /* interleaved float32 (re, im)                             */
#define DDM_IN_CU8 1         /* interleaved uint8 (I, Q), value - 127.5 (source.py:117)  */

int ddm_chain_create(int device,
                     const double *taps, int ntaps,     /* FIR b[] (a = [1])             */
                     int decim,                         /* jumpIndex of comm.py:119, >=1 */
                     double freq_offset, double samp_rate, /* comm.py:77; 0 = no mixer   */
                     int out_mode,                      /* DDM_CHAIN_OUT_*               */
                     int in_format,                     /* DDM_IN_*                      */
                     ddm_chain **out);
int ddm_chain_destroy(ddm_chain *c);
/* back to the reference's initial state (n0 = 0, offset 0, zi = lfilter_zi, no last) */
int ddm_chain_reset(ddm_chain *c);
/* halo length in input samples */
int ddm_chain_halo_len(const ddm_chain *c, int64_t *n);
/* number of output samples the NEXT apply call will produce for n input samples */
int ddm_chain_out_count(const ddm_chain *c, int64_t n, int64_t *n_out);
/* carried scalars: global index, decimation offset (comm.py:124), has-previous flag */
int ddm_chain_get_position(const ddm_chain *c, int64_t *n0, int64_t *dec_off, int *has_prev);
/* time-sharding seam: place a handle at an arbitrary point of a stream.  halo_dev holds the
 * halo_len raw input samples that precede global index n0 (NULL = the reference's initial
 * condition, only meaningful at n0 = 0). */
int ddm_chain_set_position(ddm_chain *c, int64_t n0, int64_t dec_off, int has_prev,
                           const void *halo_dev, void *stream);
/* copy the current halo (the last halo_len raw input samples seen) to halo_dev */
int ddm_chain_get_halo(const ddm_chain *c, void *halo_dev, void *stream);

/* One chunk of the stream: x_dev (device, n samples in the handle's input format) -> out_dev (device,
 * f32 or cf32), *n_out results; the carried state advances by n samples.  One kernel launch for the
 * configurations of the fused fast path (decimation >= 2, at most 10 partial sums per output: 151
 * taps from D = 16 on), which is either the warp-autonomous kernel (per-warp TMA rings, partial sums
 * exchanged by shuffles) or the CTA-tiled one, chosen per block length (chain.cu, stream_geometry);
 * any other configuration runs the general float64-accumulating kernels.  In a chunk loop the launch of
 * chunk k+1 overlaps the tail of chunk k (programmatic dependent launch; everything that reads the chunk
 * or the carried state waits for all earlier work of the stream), and the kernel itself saves the last
 * halo_len samples for the next call -- no other node is queued per chunk.  The chunk must stay
 * valid and unmodified until the launch has completed in stream order, as for any asynchronous call. */
int ddm_chain_apply_dev(ddm_chain *c, const void *x_dev, int64_t n,
                        void *out_dev, int64_t out_capacity, int64_t *n_out, void *stream);
/* `batch` independent captures of n samples each (capture k at x_dev + k*x_stride samples), every
 * one demodulated as a fresh stream from the reference's initial state, in ONE launch (BASELINE
 * config 5: 256 separate captures).  Output k at out_dev + k*out_stride elements, *n_out outputs
 * per capture.  The handle is reset and left in its reset state. */
int ddm_chain_apply_batch_dev(ddm_chain *c, const void *x_dev, int64_t n, int64_t batch, int64_t x_stride,
                              void *out_dev, int64_t out_stride, int64_t *n_out, void *stream);
/* The same for HOST buffers (what source.read(a, b) returned; the audio array): copies in, runs the
 * chain, copies the results back, returns when they are there.  Chunks of >= 2^22 samples go through in
 * four pieces whose host->device copies are queued ahead of the kernels on a second stream.  Pinned
 * host memory makes the copies asynchronous; pageable memory works, slower. */
int ddm_chain_apply_host(ddm_chain *c, const void *x_host, int64_t n,
                         void *out_host, int64_t out_capacity, int64_t *n_out, void *stream);

/* the chain's carried state in the reference's own terms, for handing a stream over to the
 * un-fused operators: zi_c128_host receives ntaps-1 complex128 values = filter.__zi
 * (filters.py:69) and last_c128_host one complex128 = demod_fm.__last (demod_fm.py:44,48;
 * only meaningful when has_prev).  Synchronises the stream. */
int ddm_chain_export_state(const ddm_chain *c, double *zi_c128_host, double *last_c128_host,
                           void *stream);

/* ---- stand-alone element-wise operators ----------------------------------------------- */
/* comm.py:63-78  x[i] *= exp(-j 2 pi f (n0+i) / fs), in place on cf32 */
int ddm_mix_cf32(int device, void *x_dev, int64_t n, double freq_offset, double samp_rate,
                 int64_t n0, void *stream);
/* same with a per-sample frequency array (f64, device) -- decode_funcube.py:228 */
int ddm_mix_var_cf32(int device, void *x_dev, const double *freq_dev, int64_t n, double samp_rate,
                     int64_t n0, void *stream);
/* demod_fm.py:29-51  out = angle(x[i] conj x[i-1]); prev_dev = carried last sample (one cf32
 * on the device) or NULL for the first chunk (then n-1 outputs).  *n_out = outputs written. */
int ddm_fm_demod(int device, const void *x_dev, int64_t n, const void *prev_dev, void *out_dev,
                 int64_t *n_out, void *stream);
/* demod_fm.py:74-96  diff(unwrap(angle(x))); prev_sample_dev / last_sample_dev: one cf32 each
 * (the reference carries the last ANGLE in float64; carrying the sample and re-evaluating its angle
 * in float64 gives the same value whatever the chunking) */
int ddm_fm_angle_diff(int device, const void *x_dev, int64_t n, const void *prev_sample_dev,
                      void *out_dev, void *last_sample_dev, int64_t *n_out, void *stream);
/* np.abs of a cf32 (is_complex) or f32 array -> f32   (demod_am.py:29,62) */
int ddm_abs(int device, const void *x_dev, int64_t n, int is_complex, void *out_dev, void *stream);
/* np.sign of an f32 array (decode_afsk1200.py:157) */
int ddm_sign(int device, const void *x_dev, int64_t n, void *out_dev, void *stream);
/* comm.py:127  out = x[offset::step]; elem_bytes 4 (f32), 8 (cf32) or 16 (c128) */
int ddm_stride_copy(int device, const void *x_dev, int64_t n, int elem_bytes, int64_t offset,
                    int64_t step, void *out_dev, int64_t *n_out, void *stream);
/* filters.py:322-326  scipy.signal.medfilt(x, kernel_size) on f32 (odd kernel_size <= 255, zero padded) */
int ddm_medfilt(int device, const void *x_dev, int64_t n, int kernel_size, void *out_dev, void *stream);
/* source.py:117-118 / :209-210  interleaved u8 IQ -> cf32 minus (127.5 + 127.5j) */
int ddm_cu8_to_cf32(int device, const void *iq_u8_dev, int64_t n, void *out_dev, void *stream);

/* ---- stateful linear filters (filters.py:21-75) ----------------------------------------
 * One handle = one filters.filter object: coefficients b, a (normalised by a[0] like scipy)
 * and the carried delay line zi (scipy's direct-form-II-transposed state, max(na,nb)-1
 * complex128 values).  FIR (a == [1]) runs on the FP32 pipes, IIR as a float64 blocked scan.
 * Signals are f32 (is_complex = 0) or cf32 (is_complex = 1) on the handle's device. */
typedef struct ddm_filter ddm_filter;
int ddm_filter_create(int device, const double *b, int nb, const double *a, int na, ddm_filter **out);
int ddm_filter_destroy(ddm_filter *f);
int ddm_filter_state_len(const ddm_filter *f, int *n);
/* host <-> handle copy of zi (interleaved re, im float64); both synchronise the stream */
int ddm_filter_set_state(ddm_filter *f, const double *zi_c128_host, void *stream);
int ddm_filter_get_state(const ddm_filter *f, double *zi_c128_host, void *stream);
/* zi = lfilter_zi(b, a), used unscaled like filters.py:45 */
int ddm_filter_reset(ddm_filter *f, void *stream);
/* use_state = 1: y, zi = lfilter(b, a, x, zi=zi) (filters.py:69); 0: lfilter(b, a, x) (:75) */
int ddm_filter_apply_dev(ddm_filter *f, const void *x_dev, int64_t n, int is_complex, void *y_dev,
                         int use_state, void *stream);
/* scipy.signal.filtfilt(b, a, x) with its defaults (filters.py:73); does not touch zi */
int ddm_filter_filtfilt_dev(ddm_filter *f, const void *x_dev, int64_t n, int is_complex, void *y_dev,
                            void *stream);
/* replace the handle's lfilter_zi vector (used by reset and to seed filtfilt) -- a caller that
 * already holds scipy's own lfilter_zi passes it here so that even ill-conditioned filters start
 * from the reference's exact bits */
int ddm_filter_set_zi_base(ddm_filter *f, const double *zi_host);
/* IIR execution mode.  scipy's float64 transfer-function recursion has a filter-dependent
 * roundoff noise floor (3e-4 relative for the reference's 12th-order NOAA band-pass,
 * decode_noaa.py:274); a run that is not the very same sequential loop cannot agree with it
 * better than that floor.  AUTO measures the floor at create time and replays the loop
 * sequentially (one thread, bit-exact float64) when it exceeds 1e-7, otherwise runs
 * segment-parallel.  The segment-parallel mode contracts every multiply-add pair of the recursion
 * into one DFMA (half the FP64 issues: the kernel becomes HBM-bound); PARALLEL_EXACT keeps scipy's
 * separately rounded operations in the segments as well. */
#define DDM_IIR_AUTO 0
#define DDM_IIR_PARALLEL 1
#define DDM_IIR_SEQUENTIAL 2
#define DDM_IIR_PARALLEL_EXACT 3
int ddm_filter_set_iir_mode(ddm_filter *f, int mode);
/* The switch of DDM_IIR_AUTO: filters whose measured float64 roundoff floor (ddm_filter_info) is above
 * `floor` replay the loop sequentially, the others run segment-parallel.  Default 1e-7: AUTO never
 * gives up agreement with the reference that the 1e-5 parity tolerance could see.  A caller whose
 * tolerance is stated against scipy's float64 result (which itself is only good to the filter's
 * floor) may raise it -- e.g. 1e-3 lets the 12th-order band-passes of the decoders run parallel.
 * floor <= 0 restores the default. */
int ddm_filter_set_iir_auto_floor(ddm_filter *f, double floor);
/* FIR execution path: direct register-tiled convolution (FP32-pipe bound, cost grows with the tap
 * count) or overlap-save through 4096-point FFTs in shared memory (HBM bound, up to 2049 taps).
 * AUTO takes the FFT path from 96 taps on. */
#define DDM_FIR_AUTO 0
#define DDM_FIR_DIRECT 1
#define DDM_FIR_FFT 2
int ddm_filter_set_fir_mode(ddm_filter *f, int mode);
/* is_fir, warm-up length of the segment-parallel IIR (-1: never decays), measured noise floor */
int ddm_filter_info(const ddm_filter *f, int *is_fir, int64_t *warmup, double *noise_floor);
/* host only (no device needed): warm-up length of the segment-parallel IIR (-1: the zero-input
 * response never decays) and the measured float64 roundoff floor; 0 / 0.0 for a FIR */
int ddm_iir_analyse(const double *b, int nb, const double *a, int na, int64_t *warmup, double *noise_floor);
/* scipy.signal.lfilter_zi restated; zi_out has max(na,nb)-1 entries */
int ddm_lfilter_zi(const double *b, int nb, const double *a, int na, double *zi_out);

/* ---- FFT-defined operators (float64 on the device) -------------------------------------
 * A context owns the Bluestein/power-of-two plans (cached per length) and the workspace. */
typedef struct ddm_fft ddm_fft;
int ddm_fft_create(int device, ddm_fft **out);
int ddm_fft_destroy(ddm_fft *c);
/* demod_am.py:29 abs(hilbert(x)) applied per chunk exactly like decode_noaa.__getAM
 * (decode_noaa.py:631-657; chunk bounds per chunker.py:32-45).  chunk >= n: one transform over
 * the whole array (demod_am().demod).  x, out: f32 on the device; all chunks of equal length
 * run as one batch. */
int ddm_am_hilbert(ddm_fft *c, const void *x_dev, int64_t n, int64_t chunk, void *out_dev, void *stream);
/* scipy.signal.resample(x, num) (comm.py:114 strict bwLim, decode_noaa.py:350-351): f32 -> f32
 * (real branch: rfft/irfft semantics) or cf32 -> cf32 (complex branch) */
int ddm_resample(ddm_fft *c, const void *x_dev, int64_t n, int is_complex, int64_t num, void *out_dev,
                 void *stream);

/* ---- sync correlation and peak picking (decode_noaa.py:659-767) -------------------------
 * out[i] = sum_k hay[i - m/2 + k] needle[k]  (signal.correlate 'same'), divided by
 * sqrt(sum_k hay[i - m/2 + k]^2 * sum(needle^2)) when normalised (decode_noaa.__correlate).
 * hay: f32 or f64 on the device; needle: host f64; out: f64 on the device.  Piecewise-constant
 * needles (the APT sync words) run on two-level prefix sums, others on the direct kernel. */
int ddm_correlate(int device, const void *hay_dev, int64_t n, int hay_is_f64, const double *needle_host,
                  int m, int normalised, void *out_f64_dev, void *stream);
/* sum of the k largest and of the k smallest values of a float64 device array -- what
 * np.sum(cor[np.argpartition(cor, -k)[-k:]]) and its mirror compute (decode_noaa.py:714-720) */
int ddm_topk_sums(int device, const void *x_f64_dev, int64_t n, int64_t k, double *sum_top,
                  double *sum_bottom, void *stream);
/* np.argwhere(x > threshold), ascending, with the values (decode_noaa.py:723); *count is the
 * number found even when it exceeds capacity (call again with larger buffers) */
int ddm_compact_above(int device, const void *x_f64_dev, int64_t n, double threshold, void *idx_i64_dev,
                      void *val_f64_dev, int64_t capacity, int64_t *count, void *stream);
/* host: the sequential group-maximum scan of decode_noaa.py:731-746 over (idx, val) candidates */
int ddm_group_peaks(const int64_t *idx, const double *val, int64_t count, double min_dist, int64_t *peaks,
                    int64_t capacity, int64_t *n_peaks);
/* decode_noaa.py:723-746 in one call, on the device: candidates x > threshold, group-maximum scan
 * with groups closed at distance >= min_dist from the running maximum (strict '<': the first of
 * equal maxima wins).  Equals ddm_compact_above + ddm_group_peaks, but no candidate leaves the
 * device (a noisy pass has millions): only the "dominant" candidates -- those no later sample within
 * the window exceeds, a few thousand -- are listed, and the walk over them runs on the host.
 * peaks_host receives the ascending peak indices. */
int ddm_pick_peaks(int device, const void *x_f64_dev, int64_t n, double threshold, double min_dist,
                   int64_t *peaks_host, int64_t capacity, int64_t *n_peaks, void *stream);
/* decode_afsk1200.py:106-142: four nbuf-tap correlators over a real signal (taps4_host =
 * [mark cos | mark sin | space cos | space sin], host f64); out[s] = mi^2 + mq^2 - si^2 - sq^2 for
 * s < n - nbuf, 0 for the last nbuf samples; out f32 on the device */
int ddm_bank4(int device, const void *x_dev, int64_t n, int x_is_f64, const double *taps4_host, int nbuf,
              void *out_f32_dev, void *stream);

/* ---- many equal-length rows at once (the accurate-sync windows of decode_noaa.py:844-877) ----
 * Each row is an independent short signal; batching them removes ~60 launches per window. */
/* offsetFreq with the sample index restarting at 0 on every row (cf32 [rows][row_len], in place) */
int ddm_mix_rows_cf32(int device, void *x_dev, int64_t rows, int64_t row_len, double freq_offset,
                      double samp_rate, void *stream);
/* filtfilt(b, a, row) for every row; x, y: [rows][n] f32 or cf32 */
int ddm_filter_filtfilt_rows_dev(ddm_filter *f, const void *x_dev, int64_t rows, int64_t n, int is_complex,
                                 void *y_dev, void *stream);
/* a fresh demod_fm per row: out [rows][row_len - 1] f32 */
int ddm_fm_demod_rows(int device, const void *x_dev, int64_t rows, int64_t row_len, void *out_dev, void *stream);
/* first index and value of the maximum of every row of a float64 matrix */
int ddm_rows_argmax(int device, const void *x_f64_dev, int64_t rows, int64_t row_stride, int64_t row_len,
                    void *idx_i64_dev, void *val_f64_dev, void *stream);
/* mean of x[row_start[r] : row_start[r] + len] (f32 in, f64 out) */
int ddm_rows_mean(int device, const void *x_dev, const int64_t *row_start_dev, int64_t rows, int64_t len,
                  void *out_f64_dev, void *stream);

/* ---- image-line assembly helpers (decode_noaa.getImage, decode_noaa.py:255-465) ---------
 * signal.resample of many equal-length rows of one f32 array in one batch: row r is
 * x[row_start[r] : row_start[r] + n] -> out[r][0:num]   (decode_noaa.py:350-351) */
int ddm_resample_rows(ddm_fft *c, const void *x_dev, const int64_t *row_start_dev, int64_t rows, int64_t n,
                      int64_t num, void *out_dev, void *stream);
/* np.median of rows of one f32 array: out[r] = median(x[s : s + row_len]), s = row_start[r], or
 * r * row_stride when row_start_dev is NULL; out f64 on the device (decode_noaa.py:317,355,372,452) */
int ddm_row_medians(int device, const void *x_dev, const int64_t *row_start_dev, int64_t rows, int64_t row_stride,
                    int64_t row_len, void *out_f64_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DDEMOD_H */
