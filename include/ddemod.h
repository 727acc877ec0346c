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

int ddm_chain_apply_dev(ddm_chain *c, const void *x_dev, int64_t n,
                        void *out_dev, int64_t out_capacity, int64_t *n_out, void *stream);
int ddm_chain_apply_host(ddm_chain *c, const void *x_host, int64_t n,
                         void *out_host, int64_t out_capacity, int64_t *n_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DDEMOD_H */
