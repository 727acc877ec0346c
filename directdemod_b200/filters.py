"""Filters (directdemod/filters.py): same classes, arguments and state semantics as the
reference; the arithmetic of ``applyOn`` runs on the GPU through libddemod.so
(ddm_filter_* in include/ddemod.h).  Coefficient *design* stays on the host in scipy, so
``b, a`` are identical to the reference's by construction.

``applyOn(x)`` accepts a numpy array (returns a numpy array in the reference's dtype,
float64 / complex128) or a cuda torch tensor (returns a cuda tensor, float32 / complex64 --
the device-resident handoff commSignal uses to avoid PCIe round trips between operators).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.signal as signal

from . import _dev, _lib, constants


# Default of filter.setIIRTolerance for filters created afterwards (None = the library's 1e-7): a
# script that accepts scipy-float64-level agreement everywhere sets it once, e.g.
# ``filters.IIR_AUTO_FLOOR = 1e-3``, and its 12th-order band-passes stream segment-parallel.
IIR_AUTO_FLOOR = None


class filter:
    """Parent of all filters (filters.py:15-89).

    storeState: carry the lfilter delay line between calls (zi starts as the UNSCALED
    ``lfilter_zi(b, a)``, filters.py:45).  zeroPhase: ``filtfilt`` instead, which disables
    storeState and initOut (filters.py:38-42).  initOut: first call seeds zi with
    ``lfiltic(b, a, x, initOut)`` (filters.py:66-67).
    """

    _ddm_native = True      # commSignal may hand this object device tensors

    def __init__(self, b, a, storeState=True, zeroPhase=False, initOut=None):
        self._storeState = bool(storeState)
        self._zeroPhase = bool(zeroPhase)
        self._initOut = initOut
        if self._storeState and self._zeroPhase:
            self._storeState = False
        if self._initOut is not None and self._zeroPhase:
            self._initOut = None
        self._b = b
        self._a = a
        self._bd = np.atleast_1d(np.asarray(b, dtype=np.float64)).copy()
        self._ad = np.atleast_1d(np.asarray(a, dtype=np.float64)).copy()
        # the reference computes lfilter_zi in its constructor (filters.py:45) and so raises there for a
        # filter without a steady state (pole at z = 1); an FIR always has one, and for 1023 taps scipy's
        # companion-matrix solve takes 0.1 s, so only recursive filters are checked eagerly
        self._zi0 = None
        if self._storeState and np.any(self._ad[1:] != 0.0):
            self._zi0 = np.ascontiguousarray(signal.lfilter_zi(self._b, self._a), dtype=np.float64)
        self._h = None              # ddm_filter handle, created on first use (needs the GPU)
        self._needs_lfiltic = self._storeState and self._initOut is not None
        self._chain = None          # fused chain currently holding this filter's state
        self._used = False

    # -- handle management ------------------------------------------------------------
    # Filters without carried state (zeroPhase / storeState=False) are pure functions of their
    # coefficients: their native handles are shared per (device, b, a), so code that builds a new
    # filter object per window (decode_noaa.getAccurateSync, decode_noaa.py:852) does not pay for
    # device allocations and coefficient uploads every time.
    _shared_handles = {}

    def _handle(self, dev=None):
        """The native handle on device ``dev`` (the device of the tensor being filtered; torch's
        current device when no tensor is involved yet).  Filters without carried state follow their
        input from device to device (shared handles are per device); a filter that carries a delay
        line lives on the device it first ran on and refuses tensors from another one."""
        if dev is None:
            dev = self._dev_index if self._h is not None else _dev.device_index()
        dev = int(dev)
        if self._h is not None and dev != self._dev_index:
            if self._storeState and self._used:
                raise ValueError("this filter carries state on cuda:%d and was handed a tensor on cuda:%d"
                                 % (self._dev_index, dev))
            if getattr(self, "_owns_handle", True):
                _lib.lib().ddm_filter_destroy(self._h)
            self._h = None
        shared = not self._storeState and not getattr(self, "_private", False)
        key = (dev, self._zeroPhase, IIR_AUTO_FLOOR, self._bd.tobytes(), self._ad.tobytes())
        if self._h is None and shared:
            _dev.require_cuda()
            cached = filter._shared_handles.get(key)
            if cached is not None:
                self._h = cached
                self._owns_handle = False
                self._dev_index = dev
                return self._h
        if self._h is None:
            _dev.require_cuda()
            l = _lib.lib()
            h = C.c_void_p()
            _lib.check(l.ddm_filter_create(
                dev, self._bd.ctypes.data_as(C.POINTER(C.c_double)), self._bd.size,
                self._ad.ctypes.data_as(C.POINTER(C.c_double)), self._ad.size, C.byref(h)),
                "ddm_filter_create")
            self._h = h
            self._dev_index = dev
            self._owns_handle = True
            if self._state_len() > 0 and (self._storeState or self._zeroPhase):
                # scipy's own lfilter_zi, so even ill-conditioned filters start from the very
                # bits the reference starts from (filters.py:45)
                zi = self._zi0 if self._zi0 is not None else \
                    np.ascontiguousarray(signal.lfilter_zi(self._b, self._a), dtype=np.float64)
                _lib.check(l.ddm_filter_set_zi_base(h, zi.ctypes.data_as(C.POINTER(C.c_double))),
                           "ddm_filter_set_zi_base")
            if IIR_AUTO_FLOOR is not None and not hasattr(self, "_shard_floor"):
                self._shard_floor = float(IIR_AUTO_FLOOR)
            if getattr(self, "_shard_floor", None) is not None:
                _lib.check(l.ddm_filter_set_iir_auto_floor(h, self._shard_floor), "ddm_filter_set_iir_auto_floor")
            if self._storeState and not self._needs_lfiltic:
                _lib.check(l.ddm_filter_reset(h, _dev.stream_ptr(dev)), "ddm_filter_reset")
            if shared and len(filter._shared_handles) < 256:
                # one stream at a time: the handle's seed slot and the filtfilt scratch are shared too
                filter._shared_handles[key] = h
                self._owns_handle = False
        return self._h

    def _sync_pending(self):
        """commSignal.filter() only queues a stateful filter; anything that reads or advances the
        delay line directly must first let that queue run, so state is consumed in call order like
        in the reference (which executes eagerly).  A filter whose state currently lives in a fused
        cascade (commSignal ran ``.filter(f1).filter(f2)`` as one equivalent filter) takes it back."""
        self._flush_owner()
        self._leave_cascade()

    def _flush_owner(self):
        owner = getattr(self, "_pending_owner", None)
        if owner is not None:
            owner._flush()

    def _leave_cascade(self):
        cas = self.__dict__.get("_cascade")
        if cas is not None:
            cas.release()

    def _unshare(self):
        """Execution modes are per-handle settings: a filter that changes one gets its own handle."""
        if getattr(self, "_h", None) is not None and not getattr(self, "_owns_handle", True):
            self._h = None
        self._private = True
        self._owns_handle = True

    def setIIRMode(self, mode):
        """0 auto (default), 1 segment-parallel (DFMA-contracted), 2 sequential bit-exact replay,
        3 segment-parallel with scipy's separately rounded operations (see ddemod.h)."""
        self._unshare()
        _lib.check(_lib.lib().ddm_filter_set_iir_mode(self._handle(), int(mode)), "ddm_filter_set_iir_mode")
        return self

    def setIIRTolerance(self, floor):
        """Where AUTO switches from segment-parallel to the sequential bit-exact replay: filters whose
        measured float64 roundoff floor (``analysis()[1]``) exceeds ``floor`` are replayed.  The default
        (1e-7) keeps every result within the 1e-5 parity tolerance of the reference's bits; a decoder
        whose own tolerance is stated against scipy's float64 output -- which for the 12th-order
        band-passes is itself only good to 4e-5 .. 3e-4 -- can raise it (decode_noaa.getImage and
        afsk.front_end do) and stream such filters at HBM speed."""
        self._unshare()
        self.__dict__.pop("_cascadable", None)        # (commSignal's cached verdict depends on this switch)
        self.__dict__.pop("_no_fir_form", None)
        self._shard_floor = float(floor) if floor > 0 else 1e-7
        _lib.check(_lib.lib().ddm_filter_set_iir_auto_floor(self._handle(), float(floor)),
                   "ddm_filter_set_iir_auto_floor")
        return self

    def analysis(self):
        """(warm-up length of the segment-parallel IIR, measured roundoff floor) -- host only,
        needs no device; (0, 0.0) for a FIR."""
        w, nf = C.c_int64(), C.c_double()
        _lib.check(_lib.lib().ddm_iir_analyse(
            self._bd.ctypes.data_as(C.POINTER(C.c_double)), self._bd.size,
            self._ad.ctypes.data_as(C.POINTER(C.c_double)), self._ad.size, C.byref(w), C.byref(nf)),
            "ddm_iir_analyse")
        return int(w.value), float(nf.value)

    def lookback(self):
        """Input history (samples) a time slab needs from its predecessor for this filter:
        ntaps-1 for a FIR, the warm-up length for a segment-parallel IIR.  Raises for filters
        that only run as a sequential replay (their state is a serial dependency)."""
        if self.isFIR:
            return max(self._bd.size, self._ad.size) - 1
        w, nf = self.analysis()
        if w < 0 or nf > getattr(self, "_shard_floor", 1e-7):
            raise ValueError("this IIR runs as a sequential bit-exact replay (roundoff floor %.1e) and cannot "
                             "be time-sharded; shard it by independent units" % nf)
        return w

    def setFIRMode(self, mode):
        """0 auto (default), 1 direct convolution, 2 overlap-save FFT (see ddemod.h)."""
        self._unshare()
        _lib.check(_lib.lib().ddm_filter_set_fir_mode(self._handle(), int(mode)), "ddm_filter_set_fir_mode")
        return self

    def info(self):
        """(is_fir, parallel warm-up length, measured float64 roundoff floor of the recursion)."""
        fir, w, nf = C.c_int(), C.c_int64(), C.c_double()
        _lib.check(_lib.lib().ddm_filter_info(self._handle(), C.byref(fir), C.byref(w), C.byref(nf)),
                   "ddm_filter_info")
        return bool(fir.value), int(w.value), float(nf.value)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and getattr(self, "_owns_handle", True):
                _lib.lib().ddm_filter_destroy(self._h)
            self._h = None
        except Exception:
            pass

    @property
    def isFIR(self):
        v = self.__dict__.get("_is_fir")
        if v is None:
            v = self.__dict__["_is_fir"] = bool(np.all(self._ad[1:] == 0.0))
        return v

    @property
    def _fresh(self):
        """True until the filter has processed its first sample."""
        return not self._used

    def _state_len(self):
        return max(self._bd.size, self._ad.size) - 1

    def getState(self):
        """The carried delay line as complex128 (what the reference keeps in filter.__zi)."""
        self._sync_pending()
        self._release_chain()
        z = np.zeros(2 * max(self._state_len(), 1))
        _lib.check(_lib.lib().ddm_filter_get_state(self._handle(), z.ctypes.data_as(C.POINTER(C.c_double)),
                                                   _dev.stream_ptr()), "ddm_filter_get_state")
        return (z[0::2] + 1j * z[1::2])[:self._state_len()]

    def setState(self, zi):
        zi = np.asarray(zi, dtype=np.complex128).ravel()
        if zi.size != self._state_len():
            raise ValueError("state must have %d entries" % self._state_len())
        z = np.zeros(2 * max(zi.size, 1))
        z[0:2 * zi.size:2] = zi.real
        z[1:2 * zi.size:2] = zi.imag
        self._sync_pending()
        self._chain = None
        _lib.check(_lib.lib().ddm_filter_set_state(self._handle(), z.ctypes.data_as(C.POINTER(C.c_double)),
                                                   _dev.stream_ptr()), "ddm_filter_set_state")
        self._needs_lfiltic = False
        self._used = True           # no longer "fresh": a fused kernel must start from THIS state, not lfilter_zi

    def _release_chain(self):
        """If a fused chain owns this filter's state, take it back (commSignal fused this
        filter into a mixer/decimator/FM kernel on earlier chunks)."""
        ch = self._chain
        if ch is not None:
            self._chain = None
            zi, _ = ch.export_state()
            self.setState(zi)

    # -- the operator -----------------------------------------------------------------
    def applyOn(self, x):
        """Apply the filter to a signal array (filters.py:53-75)."""
        if _dev.is_tensor(x) and x.is_cuda:
            return self._apply_dev(_dev.to_device(x))
        xd = _dev.to_device(x)
        return _dev.to_host(self._apply_dev(xd, host_x=x))

    # A stable recursive filter forgets: to 1e-9 it is an FIR of a few hundred taps (see ``cascade``).
    # For COMPLEX chunk-sized signals that form is the faster kernel: the segment-parallel recursion
    # repeats a ~1000-sample warm-up per segment and a 20 M-sample chunk does not leave segments long
    # enough to hide it (107 Gsps for the 8th-order low-pass), the overlap-save FFT kernel does not care
    # (140 Gsps for its 418-tap equivalent); slab-sized signals (290 Gsps) and real signals (half the recursion's work) stay with
    # the recursion.  The filter's own state moves in and out like a cascade stage's.
    _FIR_FORM_MIN, _FIR_FORM_MAX, _FIR_FORM_TAPS = 1 << 20, 50000000, 1025       # measured crossover: ~60 M samples

    def _fir_form(self, xd):
        if (self.isFIR or not self._storeState or self._needs_lfiltic or not xd.is_complex()
                or not (self._FIR_FORM_MIN <= xd.numel() <= self._FIR_FORM_MAX)
                or self.__dict__.get("_no_fir_form") or type(self) is cascade):
            return None
        cas = self.__dict__.get("_cascade")
        if cas is not None and cas.__dict__.get("_key") == (id(self),):
            return cas
        try:
            self.lookback()                        # recursive filters the library replays bit-exactly stay as they are
            self._leave_cascade()
            cas = cascade([self], states=[None if self._fresh else self.getState()], max_taps=self._FIR_FORM_TAPS)
        except ValueError:
            self._no_fir_form = True
            return None
        cas._key = (id(self),)
        self._cascade = cas
        return cas

    def _apply_dev(self, xd, host_x=None, _queued=False):
        if not _queued:
            self._flush_owner()
        cas = self._fir_form(xd)
        if cas is not None:
            y = cas._apply_dev(xd, _queued=True)
            self._used = True
            return y
        self._leave_cascade()
        l = _lib.lib()
        h = self._handle(xd.device.index)
        self._release_chain()
        n = xd.numel()
        cplx = xd.is_complex()
        y = _dev.empty_like_kind(n, cplx, xd.device.index)
        st = _dev.stream_ptr(xd.device.index)
        if self._storeState:
            if self._needs_lfiltic:
                # filters.py:66-67: zi = lfiltic(b, a, x, initOut) -- only the first few
                # samples of x enter, so this tiny design-time step stays on the host
                m = max(self._bd.size, self._ad.size)
                head = host_x[:m] if host_x is not None else _dev.to_host(xd[:m])
                self.setState(signal.lfiltic(self._b, self._a, np.asarray(head), self._initOut))
            _lib.check(l.ddm_filter_apply_dev(h, _dev.ptr(xd), n, int(cplx), _dev.ptr(y), 1, st),
                       "ddm_filter_apply_dev")
        elif self._zeroPhase:
            padlen = 3 * max(self._bd.size, self._ad.size)
            if n <= padlen:
                raise ValueError("The length of the input vector x must be greater than padlen, "
                                 "which is %d." % padlen)
            _lib.check(l.ddm_filter_filtfilt_dev(h, _dev.ptr(xd), n, int(cplx), _dev.ptr(y), st),
                       "ddm_filter_filtfilt_dev")
        else:
            _lib.check(l.ddm_filter_apply_dev(h, _dev.ptr(xd), n, int(cplx), _dev.ptr(y), 0, st),
                       "ddm_filter_apply_dev")
        if n > 0:
            self._used = True
        return y

    @property
    def getA(self):
        return self._a

    @property
    def getB(self):
        return self._b


class rollingAverage(filter):
    """Rolling average over n samples (filters.py:95-114)."""

    def __init__(self, n=3, storeState=True, zeroPhase=False, initOut=None):
        self._n = n
        super().__init__([1.0 / n] * n, [1], storeState, zeroPhase, initOut)


class blackmanHarris(filter):
    """Blackman-Harris window as FIR taps, unnormalised (filters.py:120-139)."""

    def __init__(self, n, storeState=True, zeroPhase=False, initOut=None):
        self._n = n
        super().__init__(signal.windows.blackmanharris(n), [1], storeState, zeroPhase, initOut)


class hamming(filter):
    """Hamming window as FIR taps, unnormalised (filters.py:180-199)."""

    def __init__(self, n, storeState=True, zeroPhase=False, initOut=None):
        self._n = n
        super().__init__(signal.windows.hamming(n), [1], storeState, zeroPhase, initOut)


class gaussian(filter):
    """Gaussian window as FIR taps, unnormalised (filters.py:205-226)."""

    def __init__(self, n, sigma, storeState=True, zeroPhase=False, initOut=None):
        self._n = n
        self._sigma = sigma
        super().__init__(signal.windows.gaussian(n, sigma), [1], storeState, zeroPhase, initOut)


class butter(filter):
    """Butterworth filter in transfer-function form (filters.py:232-273)."""

    def __init__(self, Fs, cutoffA, cutoffB=None, n=6, typeFlt=constants.FLT_LP, storeState=True,
                 zeroPhase=False, initOut=None):
        if typeFlt in (constants.FLT_BP, constants.FLT_BS) and cutoffB is None:
            raise ValueError("CutoffB must be given")
        nyq = 0.5 * Fs
        if typeFlt == constants.FLT_LP:
            b, a = signal.butter(n, cutoffA / nyq, btype="lowpass")
        elif typeFlt == constants.FLT_HP:
            b, a = signal.butter(n, cutoffA / nyq, btype="highpass")
        elif typeFlt == constants.FLT_BP:
            b, a = signal.butter(n, [cutoffA / nyq, cutoffB / nyq], btype="bandpass")
        elif typeFlt == constants.FLT_BS:
            b, a = signal.butter(n, [cutoffA / nyq, cutoffB / nyq], btype="bandstop")
        else:
            raise ValueError("Invalid filter type")
        super().__init__(b, a, storeState, zeroPhase, initOut)


class remez(filter):
    """Parks-McClellan band filter (filters.py:279-314)."""

    def __init__(self, Fs, bands, gains, ntaps=128, storeState=True, zeroPhase=False, initOut=None):
        if len(bands) == 0:
            raise ValueError("Atleast one band must be given")
        if bands[-1][1] >= (Fs / 2):
            raise ValueError("Last band must end before (Fs/2)Hz")
        edges = []
        for band in bands:
            edges.extend(band)
        if len(edges) != 2 * len(gains):
            raise ValueError("Invalid bands/gains values")
        super().__init__(signal.remez(ntaps, edges, gains, fs=Fs), [1], storeState, zeroPhase, initOut)


class cascade(filter):
    """Several stateful linear filters applied back to back, ``x -> f1 -> f2 -> ...`` (what
    ``sig.filter(f1).filter(f2)`` does in the reference, comm.py:80-92 + filters.py:64-70), as ONE
    filter -- an addition to the reference's API for long streams.

    A stable IIR stage forgets: its impulse response falls below any tolerance after a few hundred
    samples.  The cascade's impulse response h = h1 * h2 * ... is therefore, to ``tol`` of its
    absolute sum, a FINITE sequence, and the whole cascade is one FIR filter with taps h[:K] -- which
    the library applies by overlap-save FFT in a single pass over the signal (csrc/filter.cu,
    fir_fft_kernel): BASELINE configs[3]'s 1023-tap Remez followed by an 8th-order Butterworth becomes
    one 1172-tap filter (tol = 1e-9), and the intermediate signal never exists.  The reference's
    per-stage initial conditions (every stage starts from its own unscaled ``lfilter_zi``,
    filters.py:45) are carried over exactly: the zero-input response of those states through the
    cascade is the equivalent filter's initial delay line.  Coefficient work (impulse and zero-input
    responses, float64 scipy) stays on the host like all filter design.

    Raises ValueError when a stage is not a stateful LTI filter or the cascade does not die out
    within ``max_taps`` samples."""

    def __init__(self, filts, tol=1e-9, max_taps=2049, states=None):
        """``states``: the stages' current delay lines (complex arrays as from ``getState``; None for a
        stage that has not run yet) when the cascade takes over a stream in mid-flight; default: every
        stage fresh."""
        filts = list(filts)
        if not filts:
            raise ValueError("Atleast one filter must be given")
        for f in filts:
            if not isinstance(f, filter) or f._zeroPhase or not f._storeState or f._initOut is not None:
                raise ValueError("a cascade is made of stateful filters (no zeroPhase, no initOut)")
        self._stages = filts
        self._hist = None                 # the last taps-1 raw input samples (cuda), for release()
        span = 4 * max_taps
        h = np.zeros(span)
        h[0] = 1.0
        zir = np.zeros(span, dtype=np.complex128)
        for idx, f in enumerate(filts):
            bb, aa = np.asarray(f._b, dtype=np.float64), np.atleast_1d(np.asarray(f._a, dtype=np.float64))
            h = signal.lfilter(bb, aa, h)
            if max(len(bb), len(aa)) > 1:
                st = None if states is None else states[idx]
                zi0 = signal.lfilter_zi(bb, aa) if st is None else np.asarray(st, dtype=np.complex128)
                zir, _ = signal.lfilter(bb, aa, zir, zi=zi0.astype(np.complex128))
            else:
                zir = signal.lfilter(bb, aa, zir)
        tail = np.cumsum(np.abs(h)[::-1])[::-1]
        small = np.nonzero(tail <= tol * tail[0])[0]
        if small.size == 0 or not np.all(np.isfinite(h)):
            raise ValueError("the cascade's impulse response does not die out within %d samples" % span)
        k = max(int(small[0]), 2)
        # the zero-input response of the initial states must fit the equivalent filter's delay line too
        zmax = np.max(np.abs(zir)) if zir.size else 0.0
        if zmax > 0:
            late = np.nonzero(np.abs(zir) > tol * zmax)[0]
            if late.size:
                k = max(k, int(late[-1]) + 2)
        if k > max_taps:
            raise ValueError("the cascade needs %d equivalent taps (limit %d)" % (k, max_taps))
        self._zir = zir[:k - 1].astype(np.complex128)
        super().__init__(h[:k].copy(), [1.0], storeState=True)

    @property
    def stages(self):
        return list(self._stages)

    def _handle(self, dev=None):
        fresh = self._h is None
        h = super()._handle(dev)
        if fresh and self._h is not None and not self._used:
            # every stage starts from its own lfilter_zi (filters.py:45), not from the equivalent
            # filter's all-ones history
            self.setState(self._zir)
        return h

    def lookback(self):
        return self._bd.size - 1

    def _apply_dev(self, xd, host_x=None, _queued=False):
        y = super()._apply_dev(xd, host_x=host_x, _queued=_queued)
        k1 = self._bd.size - 1
        if xd.numel() >= k1:
            self._hist = xd[xd.numel() - k1:].clone()
        elif xd.numel() > 0:
            t = _dev.torch()
            old = self._hist if self._hist is not None else t.zeros(0, dtype=xd.dtype, device=xd.device)
            self._hist = t.cat([old.to(xd.dtype), xd])[-k1:]
        self._ran = getattr(self, "_ran", 0) + int(xd.numel())
        return y

    def release(self):
        """Hand the carried state back to the stages (they are about to be used on their own): each
        stage's delay line is what the stage-by-stage run leaves after the last taps-1 input samples
        -- exact for an FIR stage, and for an IIR stage as exact as the cascade itself, whose taps end
        where that stage's memory of older samples has died out.  Host float64, a few thousand samples."""
        stages = self._stages
        for f in stages:
            if f.__dict__.get("_cascade") is self:
                f._cascade = None
        if self._hist is None:
            return
        if getattr(self, "_ran", 0) < self._bd.size - 1:
            raise RuntimeError("a fused cascade cannot hand back the state of a stream shorter than its taps")
        sig = _dev.to_host(self._hist).astype(np.complex128)
        for f in stages:
            bb, aa = np.asarray(f._b, dtype=np.float64), np.atleast_1d(np.asarray(f._a, dtype=np.float64))
            if max(len(bb), len(aa)) > 1:
                sig, zf = signal.lfilter(bb, aa, sig, zi=np.zeros(max(len(bb), len(aa)) - 1, dtype=np.complex128))
                f.setState(zf)
            else:
                sig = signal.lfilter(bb, aa, sig)
        self._hist = None


class blackmanHarrisConv:
    """Blackman-Harris by 'same' convolution (filters.py:145-174); stateless.

    convolve(sig, w, 'same')[i] = sum_k w[k] sig[i + (n-1)//2 - k] with zeros outside, i.e.
    a zero-state FIR whose output is advanced by (n-1)//2 samples."""

    _ddm_native = True

    def __init__(self, n=151):
        self._n = n
        self._fir = filter(signal.windows.blackmanharris(n), [1], storeState=False)

    def applyOn(self, sig):
        dev_in = _dev.is_tensor(sig) and sig.is_cuda
        xd = _dev.to_device(sig)
        t = _dev.torch()
        n = xd.numel()
        # 'same' keeps len(sig) samples centred on the full convolution (scipy.signal.convolve
        # sizes the output after its FIRST argument, whichever operand is longer)
        adv = (self._n - 1) // 2
        padded = t.cat([xd, t.zeros(adv, dtype=xd.dtype, device=xd.device)])
        y = self._fir._apply_dev(padded)[adv:adv + n]
        y = y.contiguous()
        return y if dev_in else _dev.to_host(y)


class medianFilter:
    """scipy.signal.medfilt(sig, n) (filters.py:322-326) on the GPU: window median with zero
    padding at the ends; real signals, odd n."""

    _ddm_native = True

    def __init__(self, n=5):
        self._n = n

    def applyOn(self, sig):
        if self._n % 2 != 1:
            raise ValueError("Each element of kernel_size should be odd.")
        dev_in = _dev.is_tensor(sig) and sig.is_cuda
        xd = _dev.to_device(sig)
        if xd.is_complex():
            raise TypeError("medianFilter works on real signals")
        y = _dev.empty_like_kind(xd.numel(), False, xd.device.index)
        _lib.check(_lib.lib().ddm_medfilt(xd.device.index, _dev.ptr(xd), xd.numel(), int(self._n), _dev.ptr(y),
                                          _dev.stream_ptr(xd.device.index)), "ddm_medfilt")
        return y if dev_in else _dev.to_host(y)
