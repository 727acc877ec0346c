"""Sync correlation and peak picking on the GPU: the arithmetic of
decode_noaa.__correlate / __correlateAndFindPeaks (decode_noaa.py:659-767) behind small
functions, used by directdemod_b200.decode_noaa and callable on their own."""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _dev, _lib, constants


def sync_needle(sync_bits, samp_rate, positive=True):
    """decode_noaa.py:689-694: the sync word stretched to the sampling rate; positive:
    (bits*233 + 11)/255, otherwise bits - 0.5."""
    rep = round(samp_rate * constants.NOAA_T)
    if positive:
        return ((np.repeat(sync_bits, rep) * 233) + 11) / 255
    return np.repeat(sync_bits, rep) - 0.5


def correlate(hay, needle, normalised=True):
    """signal.correlate(hay, needle, 'same'), optionally normalised like
    decode_noaa.__correlate.  hay: numpy array or cuda float32/float64 tensor (real).
    Returns a cuda float64 tensor."""
    t = _dev.require_cuda()
    if not (_dev.is_tensor(hay) and hay.is_cuda):
        a = np.ascontiguousarray(np.asarray(hay, dtype=np.float64))
        if a.ndim != 1:
            raise TypeError("The signal array must be 1-D")
        hay = t.from_numpy(a).to("cuda")
    if hay.is_complex():
        raise TypeError("correlation of complex signals is not part of the hot path")
    if hay.dtype not in (t.float32, t.float64):
        hay = hay.to(t.float64)
    hay = hay.contiguous()
    needle = np.ascontiguousarray(np.asarray(needle, dtype=np.float64))
    out = t.empty(hay.numel(), dtype=t.float64, device=hay.device)
    _lib.check(_lib.lib().ddm_correlate(
        hay.device.index, _dev.ptr(hay), hay.numel(), int(hay.dtype == t.float64),
        needle.ctypes.data_as(C.POINTER(C.c_double)), needle.size, int(bool(normalised)), _dev.ptr(out),
        _dev.stream_ptr(hay.device.index)), "ddm_correlate")
    return out


def _pick_peaks_host_scan(cor, avgpk, min_dist, expected):
    """Candidates to the host, then the sequential scan (ddm_compact_above + ddm_group_peaks): the
    general form, used for degenerate minimum distances and as the cross-check of ddm_pick_peaks."""
    t = _dev.torch()
    l = _lib.lib()
    n = cor.numel()
    dev = cor.device.index
    st = _dev.stream_ptr(dev)
    cap = max(4096, 64 * expected)
    while True:
        idx = t.empty(cap, dtype=t.int64, device=cor.device)
        val = t.empty(cap, dtype=t.float64, device=cor.device)
        cnt = C.c_int64()
        _lib.check(l.ddm_compact_above(dev, _dev.ptr(cor), n, avgpk, _dev.ptr(idx), _dev.ptr(val), cap,
                                       C.byref(cnt), st), "ddm_compact_above")
        if cnt.value <= cap:
            break
        cap = int(cnt.value)
    m = int(cnt.value)
    if m == 0:
        raise TypeError("unsupported operand type(s) for -: 'NoneType' and 'int'")
    stage = _pinned(2 * m)
    stage[:m].copy_(idx[:m].view(t.float64), non_blocking=True)
    stage[m:2 * m].copy_(val[:m], non_blocking=True)
    t.cuda.current_stream(dev).synchronize()
    both = stage[:2 * m].numpy()
    idx_h = both[:m].view(np.int64)
    val_h = both[m:2 * m]
    peaks = np.empty(m, dtype=np.int64)
    npk = C.c_int64()
    _lib.check(l.ddm_group_peaks(idx_h.ctypes.data_as(C.POINTER(C.c_int64)),
                                 val_h.ctypes.data_as(C.POINTER(C.c_double)), m, float(min_dist),
                                 peaks.ctypes.data_as(C.POINTER(C.c_int64)), m, C.byref(npk)),
               "ddm_group_peaks")
    return peaks, npk


_stage = {"buf": None}


def _pinned(count):
    """A pinned float64 host buffer of at least ``count`` elements, grown geometrically."""
    t = _dev.torch()
    buf = _stage["buf"]
    if buf is None or buf.numel() < count:
        buf = t.empty(max(count + count // 4, 1 << 16), dtype=t.float64, pin_memory=True)
        _stage["buf"] = buf
    return buf


def pick_peaks(cor, samp_rate, needle_len):
    """Threshold + group-maximum scan of decode_noaa.py:710-751 on a cuda float64 tensor.
    Returns (sorted int64 numpy array of sync START positions, threshold)."""
    t = _dev.torch()
    l = _lib.lib()
    n = cor.numel()
    dev = cor.device.index
    st = _dev.stream_ptr(dev)
    expected = int(2 * (n / samp_rate)) + 2
    if expected > n:
        raise ValueError("kth(=%d) out of bounds (%d)" % (n - expected, n))     # np.argpartition's error
    top, bottom = C.c_double(), C.c_double()
    _lib.check(l.ddm_topk_sums(dev, _dev.ptr(cor), n, expected, C.byref(top), C.byref(bottom), st),
               "ddm_topk_sums")
    avgpk = top.value / expected
    avgpk -= constants.NOAA_PEAKHEIGHTWIGGLE * (avgpk - (bottom.value / expected))
    min_dist = float(constants.NOAA_MINPEAKDIST * samp_rate)
    if min_dist > 1.0:
        # threshold test + group-maximum scan entirely on the device (ddm_pick_peaks)
        cap = max(4096, 8 * expected)
        while True:
            peaks = np.empty(cap, dtype=np.int64)
            npk = C.c_int64()
            rc = l.ddm_pick_peaks(dev, _dev.ptr(cor), n, avgpk, min_dist,
                                  peaks.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(npk), st)
            if rc == _lib.ERR_CAPACITY:
                cap = int(npk.value)
                continue
            _lib.check(rc, "ddm_pick_peaks")
            break
        if npk.value == 0:
            # the reference appends currentMaxIndex == None and fails on None - int
            raise TypeError("unsupported operand type(s) for -: 'NoneType' and 'int'")
    else:
        peaks, npk = _pick_peaks_host_scan(cor, avgpk, min_dist, expected)
    peaks = peaks[:npk.value] - int(needle_len / 2)
    return np.sort(peaks), avgpk
