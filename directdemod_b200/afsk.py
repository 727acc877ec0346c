"""AFSK1200 mark/space correlator bank (directdemod/decode_afsk1200.py:106-142) on the GPU."""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib


def bank_taps(bw, baud=1200, mark=1200, space=2200):
    """The four correlator kernels exactly as decode_afsk1200.py:106-126 builds them."""
    nbuf = int(np.round(bw / baud))
    i = np.arange(nbuf)
    mark_ang = (i * 1.0 / bw) / (1 / mark) * 2 * np.pi
    space_ang = (i * 1.0 / bw) / (1 / space) * 2 * np.pi
    return np.ascontiguousarray(np.stack([np.cos(mark_ang), np.sin(mark_ang),
                                          np.cos(space_ang), np.sin(space_ang)]), dtype=np.float64), nbuf


def mark_space_bank(sig, bw, baud=1200, mark=1200, space=2200):
    """out[s] = mi^2 + mq^2 - si^2 - sq^2 over sig[s:s+buffer_size]; the last buffer_size outputs
    stay zero (decode_afsk1200.py:129-142).  numpy in -> float64 numpy out; cuda in -> cuda f32."""
    t = _dev.require_cuda()
    dev_in = _dev.is_tensor(sig) and sig.is_cuda
    if dev_in:
        x = sig.contiguous()
        if x.dtype not in (t.float32, t.float64):
            x = x.to(t.float32)
    else:
        x = t.from_numpy(np.ascontiguousarray(np.asarray(sig, dtype=np.float64))).to("cuda")
    taps, nbuf = bank_taps(bw, baud, mark, space)
    out = t.empty(x.numel(), dtype=t.float32, device=x.device)
    _lib.check(_lib.lib().ddm_bank4(x.device.index, _dev.ptr(x), x.numel(), int(x.dtype == t.float64),
                                    taps.ctypes.data_as(C.POINTER(C.c_double)), nbuf, _dev.ptr(out),
                                    _dev.stream_ptr(x.device.index)), "ddm_bank4")
    return out if dev_in else _dev.to_host(out)
