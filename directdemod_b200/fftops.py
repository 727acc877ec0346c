"""Python face of the FFT-defined kernels (ddm_am_hilbert / ddm_resample in include/ddemod.h):
one ddm_fft context per device, created on first use."""

from __future__ import annotations

import ctypes as C

from . import _dev, _lib

_ctx = {}


def _context(dev):
    h = _ctx.get(dev)
    if h is None:
        _dev.require_cuda()
        h = C.c_void_p()
        _lib.check(_lib.lib().ddm_fft_create(dev, C.byref(h)), "ddm_fft_create")
        _ctx[dev] = h
    return h


def hilbert_envelope(xd, chunk=None):
    """abs(scipy.signal.hilbert(x)) of a cuda float32 tensor; with ``chunk`` the transform is
    applied per chunk like decode_noaa.__getAM (decode_noaa.py:647)."""
    t = _dev.torch()
    if xd.is_complex():
        # scipy.signal.hilbert raises for complex input
        raise ValueError("x must be real.")
    xd = xd.to(t.float32).contiguous()
    n = xd.numel()
    out = _dev.empty_like_kind(n, False, xd.device.index)
    _lib.check(_lib.lib().ddm_am_hilbert(_context(xd.device.index), _dev.ptr(xd), n,
                                         int(chunk) if chunk else max(n, 1), _dev.ptr(out),
                                         _dev.stream_ptr(xd.device.index)), "ddm_am_hilbert")
    return out


def resample(xd, num):
    """scipy.signal.resample(x, num) of a cuda float32 / complex64 tensor."""
    num = int(num)
    n = xd.numel()
    if n == 0:
        raise ValueError("cannot resample an empty signal")
    out = _dev.empty_like_kind(num, xd.is_complex(), xd.device.index)
    _lib.check(_lib.lib().ddm_resample(_context(xd.device.index), _dev.ptr(xd.contiguous()), n,
                                       int(xd.is_complex()), num, _dev.ptr(out),
                                       _dev.stream_ptr(xd.device.index)), "ddm_resample")
    return out
