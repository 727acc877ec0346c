"""Handle-level Python wrapper of the fused chain (ddm_chain_* in include/ddemod.h).

One FusedChain == the state the reference spreads over a chunker ("freqoffset", "bwlim"
variables, chunker.py:54-84), a filter object (filter.__zi, filters.py:45,69) and a
demod_fm object (demod_fm.__last, demod_fm.py:44,48), for the chain
offsetFreq -> FIR -> bwLim(non strict) -> [demod_fm] of decode_noaa.py:623.

commSignal (comm.py) builds these lazily from the fluent chain; bench.py and the
multi-GPU driver use them directly.  torch is only used to own device memory/streams.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _torch():
    import torch
    return torch


def _stream_ptr(device_index):
    torch = _torch()
    return C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


def _raw_stream_getter():
    """device index -> current stream handle as an int.  torch's C-level getter when this build has
    it (no Stream object per call on the chunk loops), else the public route."""
    torch = _torch()
    fast = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if fast is not None:
        return fast
    return lambda dev: torch.cuda.current_stream(dev).cuda_stream


class FusedChain:
    def __init__(self, taps, decim, freq_offset, samp_rate, demod=True, device=None, in_format="cf32"):
        """in_format: "cf32" (complex64 samples) or "cu8" (interleaved unsigned 8-bit I/Q exactly as
        an RTL-SDR recording stores them; the kernel applies source.py:117's ``- 127.5`` itself and
        moves a quarter of the bytes)."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("directdemod_b200 needs a CUDA device; there is no CPU fallback")
        self._l = _lib.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        taps = np.ascontiguousarray(np.asarray(taps, dtype=np.float64))
        if taps.ndim != 1 or taps.size < 1:
            raise ValueError("taps must be a non-empty 1-D array")
        self._taps = taps
        self.ntaps = int(taps.size)
        self.decim = int(decim)
        self.freq_offset = float(freq_offset)
        self.samp_rate = float(samp_rate)
        self.demod = bool(demod)
        if in_format not in ("cf32", "cu8"):
            raise ValueError("in_format must be 'cf32' or 'cu8'")
        self.in_format = in_format
        h = C.c_void_p()
        _lib.check(self._l.ddm_chain_create(
            self.device, taps.ctypes.data_as(C.POINTER(C.c_double)), self.ntaps, self.decim,
            self.freq_offset, self.samp_rate,
            _lib.CHAIN_OUT_FM if demod else _lib.CHAIN_OUT_IQ,
            _lib.IN_CU8 if in_format == "cu8" else _lib.IN_CF32, C.byref(h)),
            "ddm_chain_create")
        self._h = h
        n = C.c_int64()
        _lib.check(self._l.ddm_chain_halo_len(self._h, C.byref(n)), "ddm_chain_halo_len")
        self.halo_len = int(n.value)
        # host mirror of the carried position (the library is the authority -- see `position`): saves
        # two C calls per chunk on the 93-chunk loops of the decoders
        self._pos = (0, 0, False)
        self._fn_apply = self._l.ddm_chain_apply_dev
        self._raw_stream = _raw_stream_getter()
        self._dev_str = "cuda:%d" % self.device

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._l.ddm_chain_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- carried state ----------------------------------------------------------------
    def reset(self):
        _lib.check(self._l.ddm_chain_reset(self._h), "ddm_chain_reset")
        self._pos = (0, 0, False)

    @property
    def position_cached(self):
        """The position as mirrored on the host (no library call)."""
        return self._pos

    def _positions_in(self, n):
        off = self._pos[1]
        return 0 if n <= off else (n - off + self.decim - 1) // self.decim

    def _advance(self, n):
        """comm.py:124 and the sample counter of comm.py:76, mirrored after a chunk of n samples."""
        n0, off, hp = self._pos
        m = self._positions_in(n)
        d = self.decim
        self._pos = (n0 + n, (d - (n - off) % d) % d, hp or m > 0)

    @property
    def position(self):
        """(global sample index, decimation offset, has-previous-sample flag)."""
        n0, off, hp = C.c_int64(), C.c_int64(), C.c_int()
        _lib.check(self._l.ddm_chain_get_position(self._h, C.byref(n0), C.byref(off), C.byref(hp)),
                   "ddm_chain_get_position")
        return int(n0.value), int(off.value), bool(hp.value)

    def set_position(self, n0, dec_off, has_prev, halo=None):
        """Place the chain at global index n0; `halo` = the halo_len raw samples before it
        (cuda complex64 tensor) or None for the reference's initial condition."""
        ptr = C.c_void_p(0)
        if halo is not None:
            torch = _torch()
            want = torch.uint8 if self.in_format == "cu8" else torch.complex64
            count = 2 * self.halo_len if self.in_format == "cu8" else self.halo_len
            if halo.dtype != want or not halo.is_cuda or halo.numel() != count:
                raise ValueError("halo must be a cuda %s tensor of halo_len samples" % want)
            halo = halo.contiguous()
            ptr = C.c_void_p(halo.data_ptr())
        _lib.check(self._l.ddm_chain_set_position(self._h, int(n0), int(dec_off), int(bool(has_prev)),
                                                  ptr, _stream_ptr(self.device)),
                   "ddm_chain_set_position")
        self._pos = (int(n0), int(dec_off), bool(has_prev))

    def get_halo(self):
        torch = _torch()
        if self.in_format == "cu8":
            out = torch.empty((self.halo_len, 2), dtype=torch.uint8, device="cuda:%d" % self.device)
        else:
            out = torch.empty(self.halo_len, dtype=torch.complex64, device="cuda:%d" % self.device)
        _lib.check(self._l.ddm_chain_get_halo(self._h, C.c_void_p(out.data_ptr()),
                                              _stream_ptr(self.device)), "ddm_chain_get_halo")
        return out

    def export_state(self):
        """The carried state in the reference's terms: (zi complex128[ntaps-1] == filter.__zi,
        last complex or None == demod_fm.__last).  Used when a stream moves from the fused
        kernel to the stand-alone operators."""
        zi = np.zeros(2 * max(self.ntaps - 1, 1))
        last = np.zeros(2)
        _lib.check(self._l.ddm_chain_export_state(
            self._h, zi.ctypes.data_as(C.POINTER(C.c_double)), last.ctypes.data_as(C.POINTER(C.c_double)),
            _stream_ptr(self.device)), "ddm_chain_export_state")
        z = (zi[0::2] + 1j * zi[1::2])[:self.ntaps - 1]
        return z, (complex(last[0], last[1]) if self.position[2] else None)

    def count_for(self, n, dec_off, has_prev):
        """Outputs a chunk of n samples produces at decimation phase ``dec_off`` (pure arithmetic)."""
        n, dec_off = int(n), int(dec_off)
        m = 0 if n <= dec_off else (n - dec_off + self.decim - 1) // self.decim
        if self.demod and not has_prev:
            m = max(m - 1, 0)
        return m

    def out_count(self, n):
        m = self._positions_in(int(n))
        if self.demod and not self._pos[2]:
            m = max(m - 1, 0)
        return m

    # -- data path --------------------------------------------------------------------
    def apply(self, x, out=None):
        """x: cuda complex64 tensor (one chunk).  Returns a cuda tensor (float32 FM output or
        complex64 IQ) holding exactly the samples the reference chain returns for it."""
        torch = _torch()
        if self.in_format == "cu8":
            if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.uint8):
                raise TypeError("apply() wants a cuda uint8 tensor of interleaved I/Q for a cu8 chain")
            if x.numel() % 2:
                raise TypeError("interleaved I/Q needs an even number of bytes")
            n = x.numel() // 2
        else:
            if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.complex64):
                raise TypeError("apply() wants a cuda complex64 tensor; use apply_host() for numpy")
            if x.dim() != 1:
                raise TypeError("The signal array must be 1-D")
            n = x.numel()
        if x.device.index != self.device:
            raise ValueError("tensor is on cuda:%d, chain on cuda:%d" % (x.device.index, self.device))
        if not x.is_contiguous():
            x = x.contiguous()
        m = self.out_count(n)
        dt = torch.float32 if self.demod else torch.complex64
        exact = out is None
        if exact:
            out = torch.empty(m, dtype=dt, device=x.device)
        elif out.dtype != dt or out.numel() < m or not out.is_contiguous():
            raise ValueError("out tensor must be contiguous %s with >= %d elements" % (dt, m))
        got = C.c_int64()
        rc = self._fn_apply(self._h, x.data_ptr(), n, out.data_ptr(), out.numel(), C.byref(got),
                            self._raw_stream(self.device))
        if rc != 0:
            _lib.check(rc, "ddm_chain_apply_dev")
        self._advance(n)
        if got.value != m:
            raise RuntimeError("host mirror of the chain position is out of step (%d != %d)" % (got.value, m))
        return out if exact else out[:m]

    def apply_batch(self, x2d, out=None):
        """x2d: cuda complex64 tensor [captures, n], each row an independent capture demodulated as
        a fresh stream (reference initial state) -- all rows in one launch.  Returns [captures, m].
        The chain is reset by this call."""
        torch = _torch()
        if not (isinstance(x2d, torch.Tensor) and x2d.is_cuda and x2d.dim() == 2):
            raise TypeError("apply_batch() wants a 2-D cuda tensor [captures, samples]")
        if self.in_format == "cu8":
            raise TypeError("apply_batch() takes complex64 captures")
        x2d = x2d.contiguous()
        batch, n = x2d.shape
        self.reset()
        m = self.out_count(n)
        dt = torch.float32 if self.demod else torch.complex64
        if out is None:
            out = torch.empty((batch, m), dtype=dt, device=x2d.device)
        got = C.c_int64()
        _lib.check(self._l.ddm_chain_apply_batch_dev(
            self._h, C.c_void_p(x2d.data_ptr()), n, batch, x2d.stride(0), C.c_void_p(out.data_ptr()),
            out.stride(0), C.byref(got), _stream_ptr(self.device)), "ddm_chain_apply_batch_dev")
        return out[:, :got.value]

    def apply_host(self, x, out=None):
        """x: host complex64 array (numpy, or a pinned torch tensor).  Copies in, runs the
        chain, copies the result back; returns a numpy array."""
        if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
            x = x.numpy()
        if self.in_format == "cu8":
            x = np.ascontiguousarray(x, dtype=np.uint8).reshape(-1)
            if x.size % 2:
                raise TypeError("interleaved I/Q needs an even number of bytes")
            n = x.size // 2
        else:
            x = np.ascontiguousarray(x, dtype=np.complex64)
            if x.ndim != 1:
                raise TypeError("The signal array must be 1-D")
            n = x.size
        m = self.out_count(n)
        dt = np.float32 if self.demod else np.complex64
        if out is None:
            out = np.empty(m, dtype=dt)
        elif hasattr(out, "numpy") and not isinstance(out, np.ndarray):
            out = out.numpy()
        if out.dtype != dt or out.size < m or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous %s array with >= %d elements" % (dt, m))
        got = C.c_int64()
        _lib.check(self._l.ddm_chain_apply_host(self._h, C.c_void_p(x.ctypes.data), n,
                                                C.c_void_p(out.ctypes.data), out.size,
                                                C.byref(got), _stream_ptr(self.device)),
                   "ddm_chain_apply_host")
        self._advance(n)
        return out[:got.value]
