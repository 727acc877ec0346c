"""FM demodulators (directdemod/demod_fm.py) on the GPU (ddm_fm_demod / ddm_fm_angle_diff)."""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib


class demod_fm:
    """Polar discriminator angle(x[n] conj x[n-1]) (demod_fm.py:29-51).

    With storeState the last sample of a call is carried, so the first call returns N-1
    samples and later calls N (Experiment 5).  ``demod`` takes a numpy array (returns
    float64) or a cuda complex64 tensor (returns a cuda float32 tensor)."""

    _ddm_native = True

    def __init__(self, storeState=True):
        self._storeState = bool(storeState)
        self._last = None       # 1-element cuda complex64 tensor
        self._chain = None      # fused chain currently holding the carried sample

    @property
    def _fresh(self):
        return self._last is None and self._chain is None

    def _release_chain(self):
        ch = self._chain
        if ch is not None:
            self._chain = None
            _, last = ch.export_state()
            if last is not None:
                t = _dev.torch()
                self._last = t.tensor([last], dtype=t.complex64, device="cuda:%d" % ch.device)

    def demod(self, sig):
        if _dev.is_tensor(sig) and sig.is_cuda:
            return self._demod_dev(_dev.to_device(sig))
        return _dev.to_host(self._demod_dev(_dev.to_device(sig)))

    def _demod_dev(self, xd):
        if not xd.is_complex():
            xd = xd.to(_dev.torch().complex64)
        self._release_chain()
        n = xd.numel()
        if self._storeState and n == 0:
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")   # demod_fm.py:44
        prev = self._last if self._storeState else None
        m = n if prev is not None else max(n - 1, 0)
        out = _dev.empty_like_kind(m, False, xd.device.index)
        got = C.c_int64()
        _lib.check(_lib.lib().ddm_fm_demod(
            xd.device.index, _dev.ptr(xd), n, _dev.ptr(prev) if prev is not None else C.c_void_p(0),
            _dev.ptr(out), C.byref(got), _dev.stream_ptr(xd.device.index)), "ddm_fm_demod")
        if self._storeState:
            self._last = xd[-1:].clone()
        return out


class demod_fmAD:
    """diff(unwrap(angle(x))) with the last angle carried (demod_fm.py:57-96)."""

    _ddm_native = True

    def __init__(self, storeState=True):
        self._storeState = bool(storeState)
        self._last = None       # 1-element cuda complex64 tensor (the sample whose angle is carried)

    def demod(self, sig):
        if _dev.is_tensor(sig) and sig.is_cuda:
            return self._demod_dev(_dev.to_device(sig))
        return _dev.to_host(self._demod_dev(_dev.to_device(sig)))

    def _demod_dev(self, xd):
        t = _dev.torch()
        if not xd.is_complex():
            xd = xd.to(t.complex64)
        n = xd.numel()
        if self._storeState and n == 0:
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")   # demod_fm.py:88
        prev = self._last if self._storeState else None
        m = n if prev is not None else max(n - 1, 0)
        out = _dev.empty_like_kind(m, False, xd.device.index)
        new_last = t.empty(1, dtype=t.complex64, device=xd.device) if self._storeState else None
        got = C.c_int64()
        _lib.check(_lib.lib().ddm_fm_angle_diff(
            xd.device.index, _dev.ptr(xd), n, _dev.ptr(prev) if prev is not None else C.c_void_p(0),
            _dev.ptr(out), _dev.ptr(new_last) if new_last is not None else C.c_void_p(0),
            C.byref(got), _dev.stream_ptr(xd.device.index)), "ddm_fm_angle_diff")
        if self._storeState:
            self._last = new_last
        return out
