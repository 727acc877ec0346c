"""Narrow-band FM audio decoder (directdemod/decode_fm.py:41-72): the NOAA audio chain with
bw = 30 kHz and a strict (FFT) resample to the audio rate, one fused launch per chunk."""

from __future__ import annotations

from . import chunker, comm, demod_fm, filters


def _read(sigsrc, a, b):
    """Raw 8-bit block when the source offers it (u8 ingest in the fused kernel), else samples."""
    raw = getattr(sigsrc, "readRaw", None)
    return raw(a, b) if raw is not None else sigsrc.read(a, b)


class decode_fm:
    def __init__(self, sigsrc, offset, bw=None, audioFreq=None):
        self._bw = 30000 if bw is None else bw
        self._sigsrc = sigsrc
        self._offset = offset
        self._audioFreq = 15000 if audioFreq is None else audioFreq
        self._strictness = True

    @property
    def getAudio(self):
        audioOut = comm.commSignal(self._audioFreq)
        bhFilter = filters.blackmanHarris(151)
        fmDemdulator = demod_fm.demod_fm()
        chunkerObj = chunker.chunker(self._sigsrc)      # (the reference reads a global here, decode_fm.py:57)
        for i in chunkerObj.getChunks:
            sig = comm.commSignal(self._sigsrc.sampFreq, _read(self._sigsrc, *i), chunkerObj) \
                .offsetFreq(self._offset).filter(bhFilter).bwLim(self._bw, uniq="First") \
                .funcApply(fmDemdulator.demod).bwLim(self._audioFreq, self._strictness)
            audioOut.extend(sig)
        return audioOut
