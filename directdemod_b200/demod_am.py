"""AM demodulators (directdemod/demod_am.py) on the GPU."""

from __future__ import annotations

import ctypes as C

from . import _dev, _lib, filters, fftops


class demod_am:
    """Envelope by Hilbert transform: abs(hilbert(x)) over the whole array -- circular and
    length dependent, exactly like scipy's (demod_am.py:18-29)."""

    _ddm_native = True

    def demod(self, sig):
        if _dev.is_tensor(sig) and sig.is_cuda:
            return fftops.hilbert_envelope(_dev.to_device(sig))
        return _dev.to_host(fftops.hilbert_envelope(_dev.to_device(sig)))

    def demodChunked(self, sig, chunk):
        """decode_noaa.__getAM in one call: the envelope per ``chunk``-sample piece, all pieces
        batched into one pass of FFT launches."""
        if _dev.is_tensor(sig) and sig.is_cuda:
            return fftops.hilbert_envelope(_dev.to_device(sig), chunk)
        return _dev.to_host(fftops.hilbert_envelope(_dev.to_device(sig), chunk))


class demod_amFLT:
    """Envelope by low-pass: butter(Fs, cutoff)(abs(x)), stateful (demod_am.py:35-62)."""

    _ddm_native = True

    def __init__(self, Fs, cutoff):
        self._filter = filters.butter(Fs, cutoff)

    def demod(self, sig):
        dev_in = _dev.is_tensor(sig) and sig.is_cuda
        xd = _dev.to_device(sig)
        mag = _dev.empty_like_kind(xd.numel(), False, xd.device.index)
        _lib.check(_lib.lib().ddm_abs(xd.device.index, _dev.ptr(xd), xd.numel(), int(xd.is_complex()),
                                      _dev.ptr(mag), _dev.stream_ptr(xd.device.index)), "ddm_abs")
        y = self._filter._apply_dev(mag)
        return y if dev_in else _dev.to_host(y)
