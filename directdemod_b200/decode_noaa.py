"""NOAA APT sync path (directdemod/decode_noaa.py:600-880) on the GPU: chunked FM audio
(__audio), chunked Hilbert AM (__getAM), normalised sync correlation with peak picking
(__correlate / __correlateAndFindPeaks), getCrudeSync, getAccurateSync and the usefulness test.

Same class name, constructor and method/property names as the reference.  The image
assembly (getImage and what hangs off it: channelID, getColor, getMapImage) is the consumer
of this path and is not part of it; those members raise NotImplementedError.
"""

from __future__ import annotations

import logging

import numpy as np

from . import chunker, comm, constants, demod_am, demod_fm, filters, sync


class decode_noaa:
    def __init__(self, sigsrc, offset, bw=None):
        """sigsrc: IQ source with ``sampFreq``, ``length`` and ``read(a, b)``; offset: frequency
        offset in Hz; bw: FM bandwidth (default constants.NOAA_FMBW)."""
        self._bw = constants.NOAA_FMBW if bw is None else bw
        self._sigsrc = sigsrc
        self._offset = offset
        self._extractedAudio = None
        self._syncA = None
        self._syncB = None
        self._asyncA = None
        self._asyncB = None
        self._audOut = None
        self._asyncApk = None
        self._asyncAtime = None
        self._asyncBpk = None
        self._asyncBtime = None
        self._useNormCorrelate = None
        self._useful = 0
        self._syncCrudeSampRate = None

    # ---- not on the hot path ----------------------------------------------------------
    @property
    def channelID(self):
        raise NotImplementedError("image assembly (decode_noaa.getImage) is outside the accelerated path")

    @property
    def getImage(self):
        raise NotImplementedError("image assembly (decode_noaa.getImage) is outside the accelerated path")

    # ---- properties -------------------------------------------------------------------
    @property
    def useful(self):
        """1 if 10 consecutive syncs 0.5 s apart were found (decode_noaa.py:70-82)."""
        if self._syncA is None or self._syncB is None:
            self.getCrudeSync()
        return self._useful

    @property
    def getAudio(self):
        if self._extractedAudio is None:
            self._extractedAudio = self._audio()
        return self._extractedAudio

    # ---- stages -----------------------------------------------------------------------
    def _audio(self, audioFreq=constants.NOAA_AUDSAMPRATE, strictness=True):
        """decode_noaa.py:600-629: per PROC_CHUNKSIZE chunk, offsetFreq -> blackmanHarris(151) ->
        bwLim(bw) -> demod_fm -> bwLim(audioFreq, strictness); one fused launch per chunk."""
        logging.info('Beginning FM demodulation to get audio in chunks')
        audioOut = comm.commSignal(audioFreq)
        bhFilter = filters.blackmanHarris(151)
        fmDemdulator = demod_fm.demod_fm()
        chunkerObj = chunker.chunker(self._sigsrc)
        for num, i in enumerate(chunkerObj.getChunks):
            logging.info('Processing chunk %d of %d chunks', num + 1, len(chunkerObj.getChunks))
            sig = comm.commSignal(self._sigsrc.sampFreq, self._sigsrc.read(*i), chunkerObj) \
                .offsetFreq(self._offset).filter(bhFilter).bwLim(self._bw, uniq="First") \
                .funcApply(fmDemdulator.demod).bwLim(audioFreq, strictness)
            audioOut.extend(sig)
        logging.info('FM demodulation successfully complete')
        self._audOut = audioOut
        return audioOut

    def _getAM(self, sig):
        """decode_noaa.py:631-657: abs(hilbert) per 240 000-sample chunk; every chunk of the
        capture goes through the FFT engine as one batch."""
        logging.info('Beginning AM demodulation in chunks')
        env = demod_am.demod_am().demodChunked(sig.deviceSignal, 60000 * 4)
        logging.info('AM demodulation completed')
        return comm.commSignal(sig.sampRate, env)

    def _correlateAndFindPeaks(self, sig, syncBits, getExtraInfo=False, useNormCorrelate=True,
                               useFilter=False, usePosNeedle=True, filterType=None):
        """decode_noaa.py:677-767."""
        needle = sync.sync_needle(syncBits, sig.sampRate, usePosNeedle)
        hay = sig.deviceSignal
        if useFilter:
            if filterType is None:
                filterType = filters.hamming(492, zeroPhase=True)
            hay = filterType.applyOn(hay)
        cor = sync.correlate(hay, needle, normalised=useNormCorrelate)
        peaks, _ = sync.pick_peaks(cor, sig.sampRate, len(needle))
        if not getExtraInfo:
            return peaks
        timeSyncs, pkHeights = [], []
        m = int(len(needle))
        sigd = sig.deviceSignal
        for i in peaks:
            i = int(i)
            if i + 2 * m < sig.length:
                timeSyncs.append(float(sigd[i + m:i + 2 * m].double().mean().item()))
            else:
                timeSyncs.append(None)
            pkHeights.append(float(cor[i + int(m / 2)].item()))
        return peaks, pkHeights, timeSyncs

    def getCrudeSync(self):
        """decode_noaa.py:769-806."""
        if self._syncA is None or self._syncB is None:
            sig = self._audio(constants.NOAA_CRUDESYNCSAMPRATE, False)
            sig = self._getAM(sig)
            self._syncCrudeSampRate = sig.sampRate
            logging.info('Beginning SyncA detection')
            self._syncA = self._correlateAndFindPeaks(sig, constants.NOAA_SYNCA)
            logging.info('Beginning SyncB detection')
            self._syncB = self._correlateAndFindPeaks(sig, constants.NOAA_SYNCB)
            n = constants.NOAA_DETECTCONSSYNCSNUM
            syncAdiff = np.abs(np.diff(self._syncA) - (self._syncCrudeSampRate * 0.5))
            minSyncAdiff = np.min([np.max(syncAdiff[i:i + n]) for i in range(len(syncAdiff) - n + 1)])
            syncBdiff = np.abs(np.diff(self._syncB) - (self._syncCrudeSampRate * 0.5))
            minSyncBdiff = np.min([np.max(syncBdiff[i:i + n]) for i in range(len(syncBdiff) - n + 1)])
            if minSyncAdiff < constants.NOAA_DETECTMAXCHANGE or minSyncBdiff < constants.NOAA_DETECTMAXCHANGE:
                logging.info('NOAA Signal was found')
                self._useful = 1
            else:
                logging.info('NOAA Signal was not found')
        return [self._syncA, self._syncB]

    def _accurate(self, csync, syncBits, width, useNormCorrelate):
        out, pk, tm = [], [], []
        for i in csync:
            startI = int(i) - int(width)
            endI = int(i) + int(width)
            if startI < 0 or endI > self._sigsrc.length:
                continue
            sig = comm.commSignal(self._sigsrc.sampFreq, self._sigsrc.read(startI, endI)) \
                .offsetFreq(self._offset).filter(filters.blackmanHarris(151, zeroPhase=True)) \
                .funcApply(demod_fm.demod_fm().demod).funcApply(demod_am.demod_am().demod)
            syncDet, PkHeights, TimeSync = self._correlateAndFindPeaks(
                sig, syncBits, getExtraInfo=True, useNormCorrelate=useNormCorrelate,
                usePosNeedle=useNormCorrelate, useFilter=True)
            out.append(syncDet[0] + startI)
            pk.append(PkHeights[0])
            tm.append(TimeSync[0])
        return out, pk, tm

    def getAccurateSync(self, useNormCorrelate=True):
        """decode_noaa.py:808-880: one full-rate window of +-3 sync lengths around every crude
        sync, mixed, zero-phase filtered, FM and AM demodulated, then correlated."""
        stale = (self._asyncA is None or self._asyncB is None or self._asyncBtime is None
                 or self._asyncAtime is None or self._asyncBpk is None or self._asyncApk is None
                 or not self._useNormCorrelate == useNormCorrelate)
        if stale:
            self._useNormCorrelate = useNormCorrelate
            if self._syncA is None or self._syncB is None:
                self.getCrudeSync()
            syncTime = constants.NOAA_T * len(constants.NOAA_SYNCA)
            searchSampleWidth = int(3 * syncTime * self._sigsrc.sampFreq)
            csyncA = self._syncA / self._syncCrudeSampRate * self._sigsrc.sampFreq
            csyncB = self._syncB / self._syncCrudeSampRate * self._sigsrc.sampFreq
            logging.info('Beginning Accurate SyncA detection')
            self._asyncA, self._asyncApk, self._asyncAtime = self._accurate(
                csyncA, constants.NOAA_SYNCA, searchSampleWidth, useNormCorrelate)
            logging.info('Beginning Accurate SyncB detection')
            self._asyncB, self._asyncBpk, self._asyncBtime = self._accurate(
                csyncB, constants.NOAA_SYNCB, searchSampleWidth, useNormCorrelate)
        return [self._asyncA, np.diff(self._asyncA), self._asyncApk, self._asyncAtime,
                self._asyncB, np.diff(self._asyncB), self._asyncBpk, self._asyncBtime]
