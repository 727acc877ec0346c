"""Signal container with chainable operators (directdemod/comm.py:15-181), GPU backed.

Same constructor, properties and methods as the reference's ``commSignal``.  The samples
live on the GPU between operators; ``.signal`` hands back a numpy array in the reference's
dtype.  Operators that form the hot chain

    offsetFreq(f) -> filter(<stateful FIR>) -> bwLim(rate) -> funcApply(demod_fm().demod)

(decode_noaa.py:623, decode_fm.py:64-68, decode_afsk1200.py:79-91) are queued and executed
as ONE fused kernel launch per chunk (directdemod_b200/fused.py, csrc/chain.cu); consecutive
stateful filters over a long chunk -- ``filter(fir).filter(iir)`` -- run as one equivalent filter in
one overlap-save pass (filters.cascade); anything else runs operator by operator on the device.  All carried state (mixer sample index and
decimation phase in the chunker, filter delay line, FM last sample) keeps the reference's
meaning and can move between the fused and the stand-alone kernels at any chunk boundary.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib, constants
from . import demod_fm as _demod_fm
from . import filters as _filters
from .fused import FusedChain
from .source import RawIQ8


def _is_cuda_tensor(x):
    return _dev.is_tensor(x) and x.is_cuda


# chunks from this length on run consecutive stateful filters as one overlap-save pass (the FFT path of
# the long-FIR kernel starts there; below it the stage-by-stage kernels are the faster tool)
_CASCADE_MIN = 1 << 20


def _cascadable(f):
    """A stage filters.cascade can absorb: stateful LTI, and -- for a recursive filter -- one the library
    runs segment-parallel anyway (the bit-exact sequential replay of an ill-conditioned filter is a
    promise the equivalent FIR cannot keep)."""
    ok = f.__dict__.get("_cascadable")
    if ok is None:
        ok = bool(type(f) is not _filters.cascade and f._storeState and not f._zeroPhase and f._initOut is None)
        if ok and not f.isFIR:
            try:
                f.lookback()
            except ValueError:
                ok = False
        f._cascadable = ok
    return ok and not f._needs_lfiltic and f._chain is None


class commSignal:
    def __init__(self, sampRate, sig=np.array([]), chunker=None):
        """sampRate: Hz, forced to int, must be > 0 (ValueError); sig: 1-D array (TypeError
        otherwise), copied; chunker: chunker object when the signal is processed in chunks."""
        self._chunker = chunker
        self._len = len(sig)
        self._sampRate = int(sampRate)
        if self._sampRate <= 0:
            raise ValueError("The sampling rate must be greater than zero")
        self._pending = []
        self._host = None
        self._dev = None
        self._parts = []             # device pieces appended by extend(), joined on demand
        self._raw8 = None            # RawIQ8 block not yet converted (source.readRaw)
        self._shared = False         # _dev aliases the caller's tensor (copy before writing)
        self._pristine = False       # _dev is an untouched float32/complex64 snapshot of a host array
        if isinstance(sig, RawIQ8):
            self._raw8 = sig
            self._complex = True
        elif _is_cuda_tensor(sig):
            if sig.dim() != 1:
                raise TypeError("The signal array must be 1-D")
            # the reference copies its input (np.array(sig)); here the copy is deferred until an
            # operator would write into the caller's tensor (only the un-fused mixer works in place)
            self._dev = _dev.to_device(sig)
            self._shared = self._dev.data_ptr() == sig.data_ptr()
            self._complex = bool(sig.is_complex())
        elif self._snapshot_to_device(sig):
            pass
        else:
            arr = np.array(sig)
            if arr.ndim != 1:
                raise TypeError("The signal array must be 1-D")
            self._host = arr
            self._complex = bool(np.iscomplexobj(arr))

    def _snapshot_to_device(self, sig):
        """The reference's constructor copies its input (np.array(sig), comm.py:38).  For a large
        float32 / complex64 array -- what the sources hand out per chunk -- the copy goes straight to
        the device instead of through a second host array first (a 20 M-sample chunk: 33 ms of host
        memcpy saved); the snapshot semantics and, until an operator runs, the dtype seen through
        ``.signal`` stay the reference's."""
        if not (isinstance(sig, np.ndarray) and sig.ndim == 1 and sig.size >= (1 << 16)
                and sig.dtype in (np.complex64, np.float32) and sig.flags.c_contiguous and sig.flags.writeable):
            return False
        t = _dev.torch()
        if not t.cuda.is_available():
            return False
        _dev.require_cuda()
        self._dev = t.from_numpy(sig).to("cuda")          # synchronous for pageable memory: a snapshot
        self._complex = sig.dtype == np.complex64
        self._pristine = True
        return True

    # ---- properties -----------------------------------------------------------------
    @property
    def length(self):
        return self._len

    @property
    def sampRate(self):
        return self._sampRate

    @property
    def signal(self):
        """The samples as a numpy array (float64 / complex128 once an operator has run)."""
        self._flush()
        if self._raw8 is not None:
            self._host = self._raw8.to_complex64()
            self._raw8 = None
        if self._host is None or self._parts:
            if self._pristine and not self._parts:
                self._host = self._dev.cpu().numpy()       # the constructor's copy, in its own dtype
            else:
                self._host = _dev.to_host(self._device_array())
            self._dev = None          # the caller may now mutate the array it was given
        return self._host

    @property
    def deviceSignal(self):
        """The samples as a cuda tensor (float32 / complex64), without a host round trip."""
        self._flush()
        return self._device_array()

    def _device_array(self):
        if self._raw8 is not None:
            # bytes -> cf32 on the device (ddm_cu8_to_cf32), 2 B/sample over PCIe
            t = _dev.require_cuda()
            raw = t.from_numpy(np.array(self._raw8.data)).to("cuda")       # (copy: memmaps are read-only)
            out = _dev.empty_like_kind(raw.shape[0], True)
            _lib.check(_lib.lib().ddm_cu8_to_cf32(out.device.index, _dev.ptr(raw), raw.shape[0], _dev.ptr(out),
                                                  _dev.stream_ptr(out.device.index)), "ddm_cu8_to_cf32")
            self._raw8 = None
            self._dev = out
            self._host = None
            return self._dev
        if self._parts:
            t = _dev.torch()
            parts = self._parts
            self._parts = []
            if any(p.is_complex() for p in parts):
                parts = [p.to(t.complex64) for p in parts]
            if len(parts) == 1:
                self._dev = parts[0]
                self._shared = True      # still the tensor of the signal it was taken from
            else:
                self._dev = t.cat(parts)
                self._shared = False
            self._host = None
        elif self._dev is None:
            self._dev = _dev.to_device(self._host)
            self._host = None
        return self._dev

    # ---- operators ------------------------------------------------------------------
    def offsetFreq(self, freqOffset):
        """x[n] *= exp(-j 2 pi f (n0 + n) / fs) with the global sample index n0 carried in the
        chunker variable "freqoffset" (comm.py:63-78).  f may be a per-sample array."""
        offset = 0
        if self._chunker is not None:
            offset = self._chunker.get(constants.CHUNK_FREQOFFSET, 0)
            self._chunker.set(constants.CHUNK_FREQOFFSET, offset + self.length)
        if not self._complex:
            # the reference multiplies a real array by a complex one in place -> numpy refuses
            raise TypeError("Cannot cast ufunc 'multiply' output from dtype('complex128') to a real dtype")
        if not isinstance(freqOffset, (int, float)) and np.ndim(freqOffset) != 0:
            f = np.ascontiguousarray(np.asarray(freqOffset, dtype=np.float64).ravel())
            if f.size != self.length:
                raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,)"
                                 % (self.length, f.size))
            self._flush()
            self._run_mix_var(f, offset)
            return self
        if self._pending:            # a mixer can only open a fused chain
            self._flush()
        self._pending.append(("mix", float(freqOffset), int(offset), self._sampRate))
        return self

    def filter(self, filt):
        """Apply a filter object (anything with ``applyOn``), comm.py:80-92."""
        if isinstance(filt, _filters.filter):
            fusable = (filt.isFIR and filt._storeState and not filt._needs_lfiltic and self._complex
                       and all(op[0] == "mix" for op in self._pending))
            # a second stateful filter right behind a queued one may run with it as ONE equivalent
            # filter (filters.cascade) when the chunk is long enough: keep both queued
            chained = (self._pending and self._pending[-1][0] == "filter" and self._len >= _CASCADE_MIN
                       and _cascadable(filt) and _cascadable(self._pending[-1][1]))
            if not fusable and not chained:
                self._flush()
            self._claim(filt)
            self._pending.append(("filter", filt))
            return self
        self._flush()
        if getattr(filt, "_ddm_native", False):
            self._set_device(filt.applyOn(self._device_array()))
        else:
            self.updateSignal(filt.applyOn(self.signal))
        return self

    def bwLim(self, tsampRate, strict=False, uniq="abcd"):
        """Limit the bandwidth by decimation (comm.py:94-130).  Non strict: keep every
        int(fs/t)-th sample, phase carried in chunker variable "bwlim"+uniq, new rate
        int(fs/j).  strict: FFT resample of this chunk to exactly t (no carry)."""
        if self._sampRate < tsampRate:
            raise ValueError("The target sampling rate must be less than current sampling rate")
        if strict:
            from . import fftops
            self._flush()
            num = int(tsampRate * self.length / self.sampRate)
            self._set_device(fftops.resample(self._device_array(), num))
            self._sampRate = tsampRate
            return self
        jump = int(self.sampRate / tsampRate)
        offset = 0
        if self._chunker is not None:
            offset = self._chunker.get(constants.CHUNK_BWLIM + uniq, 0)
            nxt = (jump - (self.length - offset) % jump) % jump
            self._chunker.set(constants.CHUNK_BWLIM + uniq, nxt)
        if jump == 1 and offset == 0:
            return self              # x[0::1]: nothing to move (decode_noaa.py:623, 60235 -> 40960)
        if not (self._pending and self._pending[-1][0] == "filter"):
            self._flush()
        self._pending.append(("decim", jump, int(offset)))
        self._len = len(range(offset, self._len, jump))
        self._sampRate = int(self.sampRate / jump)
        return self

    def funcApply(self, func):
        """Apply a callable to the signal array (comm.py:132-144)."""
        owner = getattr(func, "__self__", None)
        if isinstance(owner, _demod_fm.demod_fm) and getattr(func, "__name__", "") == "demod":
            if not self._complex:
                self._flush()
                self._set_device(owner.demod(self._device_array()))
                return self
            if owner._storeState and self._len == 0:
                self._flush()
                raise IndexError("index -1 is out of bounds for axis 0 with size 0")
            self._claim(owner)
            has_prev = owner._storeState and not owner._fresh
            self._pending.append(("fm", owner))
            self._len = self._len if has_prev else max(self._len - 1, 0)
            self._complex = False
            self._flush()            # the discriminator closes the fusable pattern
            return self
        self._flush()
        if getattr(owner, "_ddm_native", False):
            self._set_device(func(self._device_array()))
        else:
            self.updateSignal(func(self.signal))
        return self

    def extend(self, sig):
        """Append another commSignal (comm.py:146-164)."""
        if self.length == 0:
            self._sampRate = sig.sampRate
        if not self._sampRate == sig.sampRate:
            raise TypeError("Signals must have same sampling rate to be extended")
        if not isinstance(sig, commSignal):
            self.updateSignal(np.concatenate([self.signal, sig.signal]))
            return self
        # the reference re-concatenates the whole accumulated array on every call (O(chunks^2)
        # bytes); here the pieces stay on the device and are joined once, when first read
        self._flush()
        sig._flush()
        # the piece is shared, not copied: `sig` is told that its tensor now has a second owner, so the
        # one operator that writes in place (the stand-alone mixer) copies first
        piece = sig._device_array()
        sig._shared = True
        if self._len == 0:
            self._parts = [piece]
            self._complex = bool(piece.is_complex())
        else:
            self._complex = self._complex or bool(piece.is_complex())
            if not self._parts:
                self._parts = [self._device_array()]
            self._parts.append(piece)
        self._dev = None
        self._host = None
        self._pristine = False
        self._len += int(piece.numel())
        return self

    def updateSignal(self, sig):
        """Replace the samples (comm.py:166-181)."""
        self._pending = []
        self._parts = []
        if _is_cuda_tensor(sig):
            if sig.dim() != 1:
                raise TypeError("The signal array must be 1-D")
            self._set_device(_dev.to_device(sig))
            return self
        arr = np.array(sig)
        if arr.ndim != 1:
            raise TypeError("The signal array must be 1-D")
        self._pristine = False
        self._host = arr
        self._dev = None
        self._raw8 = None
        self._len = len(arr)
        self._complex = bool(np.iscomplexobj(arr))
        return self

    # ---- execution ------------------------------------------------------------------
    def _set_device(self, tensor):
        self._pristine = False
        self._parts = []
        self._raw8 = None
        self._shared = False
        self._dev = tensor
        self._host = None
        self._len = int(tensor.numel())
        self._complex = bool(tensor.is_complex())

    def _claim(self, obj):
        """A stateful operator object may be queued in one signal at a time; run whoever
        queued it before so state is consumed in call order."""
        other = getattr(obj, "_pending_owner", None)
        if other is not None and other is not self:
            other._flush()
        obj._pending_owner = self

    def _flush(self):
        ops = self._pending
        if not ops:
            return
        self._pending = []
        self._pristine = False
        x = None
        i = 0
        if self._raw8 is not None:
            took, y = self._try_fused(ops, 0, None, raw=self._raw8)
            if took:
                self._raw8 = None
                x = y
                i = took
        if x is None:
            x = self._device_array()
        try:
            while i < len(ops):
                took, y = self._try_fused(ops, i, x)
                if took:
                    x = y
                    i += took
                    continue
                took, y = self._try_cascade(ops, i, x)
                if took:
                    x = y
                    i += took
                    continue
                x = self._run_single(ops[i], x)
                i += 1
        finally:
            for op in ops:
                if op[0] in ("filter", "fm") and getattr(op[1], "_pending_owner", None) is self:
                    op[1]._pending_owner = None
        if self._dev is None or x.data_ptr() != self._dev.data_ptr():
            self._shared = False     # the operators produced a fresh tensor
        self._dev = x
        self._host = None
        self._len = int(x.numel())
        self._complex = bool(x.is_complex())

    def _try_fused(self, ops, i, x, raw=None):
        """Match [mix] filter [decim] [fm] at ops[i:] and run it as one fused launch.  With
        ``raw`` (a RawIQ8 block) the chain ingests the unsigned 8-bit bytes directly."""
        j = i
        mix = filt = dec = fm = None
        if j < len(ops) and ops[j][0] == "mix":
            mix = ops[j]
            j += 1
        if j < len(ops) and ops[j][0] == "filter":
            filt = ops[j][1]
            j += 1
        else:
            return 0, None
        if j < len(ops) and ops[j][0] == "decim":
            dec = ops[j]
            j += 1
        if j < len(ops) and ops[j][0] == "fm":
            fm = ops[j][1]
            j += 1
        if dec is None or dec[1] < 2 or (raw is None and not x.is_complex()):
            return 0, None          # without a decimator the stand-alone FIR kernel is the better tool
        fmt = "cu8" if raw is not None else "cf32"
        dev_index = _dev.device_index() if raw is not None else x.device.index
        if not (filt.isFIR and filt._storeState and not filt._needs_lfiltic):
            return 0, None
        if fm is not None and not fm._storeState:
            return 0, None
        freq = mix[1] if mix is not None else 0.0
        fs = mix[3] if mix is not None else 0
        n0 = mix[2] if mix is not None else None
        jump, off = dec[1], dec[2]
        ch = filt._chain
        key = (float(freq), int(fs), int(jump), fm is not None, fmt)
        if ch is not None:
            if getattr(ch, "_key", None) != key or ch.device != dev_index:
                return 0, None
            pos_n0, pos_off, pos_prev = ch.position_cached
            if (n0 is not None and n0 != pos_n0) or off != pos_off:
                return 0, None
            if fm is not None and not (fm._chain is ch or (fm._fresh and not pos_prev)):
                return 0, None
        else:
            if not filt._fresh or (n0 is not None and n0 != 0):
                return 0, None
            if fm is not None and not fm._fresh:
                return 0, None
            if raw is not None:
                _dev.require_cuda()
            ch = FusedChain(filt._bd, jump, freq, fs if mix is not None else 1.0,
                            demod=fm is not None, device=dev_index, in_format=fmt)
            ch._key = key
            ch.set_position(0, off, False)
        if raw is not None:
            t = _dev.torch()
            y = ch.apply(t.from_numpy(np.array(raw.data)).to("cuda:%d" % dev_index))
        else:
            y = ch.apply(x)
        filt._chain = ch
        filt._used = True
        if fm is not None:
            fm._chain = ch
            fm._last = None
        return j - i, y

    def _try_cascade(self, ops, i, x):
        """Two or more queued stateful filters in a row over a long chunk: one equivalent filter
        (filters.cascade), one overlap-save pass, no intermediate signal.  The stages' states move into
        the cascade (from their current values, so a stream may switch over in mid-flight) and come
        back the moment a stage is used on its own again (filter._sync_pending -> cascade.release)."""
        j = i
        stages = []
        while j < len(ops) and ops[j][0] == "filter" and _cascadable(ops[j][1]):
            stages.append(ops[j][1])
            j += 1
        if len(stages) < 2 or x.numel() < _CASCADE_MIN:
            return 0, None
        key = tuple(id(f) for f in stages)
        cas = stages[0].__dict__.get("_cascade")
        if cas is None or cas._key != key or any(f.__dict__.get("_cascade") is not cas for f in stages):
            if stages[0].__dict__.get("_no_cascade") == key:
                return 0, None
            for f in stages:                                   # leave any other cascade first
                other = f.__dict__.get("_cascade")
                if other is not None:
                    other.release()
            try:
                cas = _filters.cascade(stages, states=[None if f._fresh else f.getState() for f in stages])
            except ValueError:
                stages[0]._no_cascade = key                    # does not die out in time: stage by stage
                return 0, None
            cas._key = key
            for f in stages:
                f._cascade = cas
        y = cas._apply_dev(x, _queued=True)
        for f in stages:
            f._used = True
        return j - i, y

    def _run_single(self, op, x):
        kind = op[0]
        if kind == "mix":
            if self._shared:
                x = x.clone()
                self._shared = False
            _lib.check(_lib.lib().ddm_mix_cf32(x.device.index, _dev.ptr(x), x.numel(), op[1], float(op[3]),
                                               op[2], _dev.stream_ptr(x.device.index)), "ddm_mix_cf32")
            return x
        if kind == "filter":
            return op[1]._apply_dev(x, _queued=True)
        if kind == "decim":
            jump, off = op[1], op[2]
            n = x.numel()
            m = len(range(off, n, jump))
            out = _dev.empty_like_kind(m, x.is_complex(), x.device.index)
            got = C.c_int64()
            _lib.check(_lib.lib().ddm_stride_copy(x.device.index, _dev.ptr(x), n, 8 if x.is_complex() else 4,
                                                  off, jump, _dev.ptr(out), C.byref(got),
                                                  _dev.stream_ptr(x.device.index)), "ddm_stride_copy")
            return out
        if kind == "fm":
            return op[1]._demod_dev(x)
        raise RuntimeError("unknown queued operator %r" % (kind,))

    def _run_mix_var(self, f, offset):
        t = _dev.torch()
        x = self._device_array()
        self._pristine = False
        if self._shared:
            x = x.clone()
            self._dev = x
            self._shared = False
        fd = t.from_numpy(f).to(x.device)
        _lib.check(_lib.lib().ddm_mix_var_cf32(x.device.index, _dev.ptr(x), _dev.ptr(fd), x.numel(),
                                               float(self._sampRate), int(offset),
                                               _dev.stream_ptr(x.device.index)), "ddm_mix_var_cf32")
