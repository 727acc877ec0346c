"""AFSK1200 front end on the GPU (directdemod/decode_afsk1200.py:74-158): the chunked
mixer/FIR/decimator chain, FM demodulation, the 700-2700 Hz Butterworth band-pass, the mark/space
correlator bank and the bit-edge correlation.  What follows in the reference -- peakdetect,
NRZI, HDLC flag search, bit unstuffing, CRC -- is small branchy host logic and stays with the
caller (out of scope for the GPU)."""

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


def bit_edges(binary_filter, bw, baud=1200):
    """decode_afsk1200.py:151-158: correlate sign(binary_filter) with a -1/+1 step kernel of one
    baud ('same'), divided by the samples per baud.  cuda f32 in -> cuda f64 out; numpy -> numpy."""
    from . import sync
    t = _dev.require_cuda()
    dev_in = _dev.is_tensor(binary_filter) and binary_filter.is_cuda
    x = binary_filter if dev_in else t.from_numpy(np.ascontiguousarray(np.asarray(binary_filter, dtype=np.float32))).to("cuda")
    x = x.to(t.float32).contiguous()
    spb = int(bw // baud)
    kernel = np.ones(spb)
    kernel[:spb // 2] = -1
    sg = t.empty_like(x)
    _lib.check(_lib.lib().ddm_sign(x.device.index, _dev.ptr(x), x.numel(), _dev.ptr(sg),
                                   _dev.stream_ptr(x.device.index)), "ddm_sign")
    changes = sync.correlate(sg, kernel, normalised=False) / spb
    return changes if dev_in else changes.cpu().numpy()


def front_end(sigsrc, offset, bw, baud=1200, mark=1200, space=2200, exact_iir=False):
    """The GPU part of decode_afsk1200.getMsg (decode_afsk1200.py:62-158).

    Returns (audio commSignal after the band-pass, binary_filter cuda f32, changes cuda f64).
    The 12th-order band-pass runs segment-parallel unless ``exact_iir``: its input already
    carries the fp32 rounding of the FM stage, which re-rolls scipy's roundoff noise (4e-5
    relative for this filter) whatever the mode, and only signs are used below."""
    from . import chunker, comm, constants, demod_fm, filters
    sig = comm.commSignal(sigsrc.sampFreq)
    chunkerObj = chunker.chunker(sigsrc)
    bhFilter = filters.blackmanHarris(151)
    fmDemodObj = demod_fm.demod_fm()
    for i in chunkerObj.getChunks:
        chunkSig = comm.commSignal(sigsrc.sampFreq, sigsrc.read(*i), chunkerObj)
        chunkSig.offsetFreq(offset)
        chunkSig.filter(bhFilter)
        chunkSig.bwLim(bw)
        sig.extend(chunkSig)
    sig.funcApply(fmDemodObj.demod)
    bp = filters.butter(sig.sampRate, mark - 500, space + 500, typeFlt=constants.FLT_BP)
    if not exact_iir:
        bp.setIIRMode(1)
    sig.filter(bp)
    audio = sig.deviceSignal
    binary_filter = mark_space_bank(audio, bw, baud, mark, space)
    changes = bit_edges(binary_filter, bw, baud)
    return sig, binary_filter, changes
