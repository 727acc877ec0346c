"""NOAA APT sync path (directdemod/decode_noaa.py:600-880) on the GPU: chunked FM audio
(__audio), chunked Hilbert AM (__getAM), normalised sync correlation with peak picking
(__correlate / __correlateAndFindPeaks), getCrudeSync, getAccurateSync and the usefulness test.

Same class name, constructor and method/property names as the reference.  ``getImage``
(decode_noaa.py:255-465), the consumer that defines the pixel-level acceptance, is assembled
here too: the whole-stream band-pass filtfilt, the Hilbert AM, every line's FFT resample (all
lines of equal length in one batch) and all medians run on the GPU; the colour-calibration
state machine stays on the host.  False colour and map overlay (getColor, getMapImage) are
cosmetic post-processing and out of scope.
"""

from __future__ import annotations

import logging

import numpy as np

import ctypes as C

from . import _dev, _lib, chunker, comm, constants, demod_am, demod_fm, fftops, filters, sync


def _row_medians(x, starts=None, rows=None, stride=0, length=1):
    """np.median over rows of the cuda float32 tensor ``x`` (ddm_row_medians) -> float64 numpy."""
    t = _dev.torch()
    if starts is not None:
        sd = t.as_tensor(np.ascontiguousarray(starts, dtype=np.int64), device=x.device)
        rows = sd.numel()
        sp = _dev.ptr(sd)
    else:
        sp = C.c_void_p(0)
    out = t.empty(int(rows), dtype=t.float64, device=x.device)
    _lib.check(_lib.lib().ddm_row_medians(x.device.index, _dev.ptr(x), sp, int(rows), int(stride), int(length),
                                          _dev.ptr(out), _dev.stream_ptr(x.device.index)), "ddm_row_medians")
    return out.cpu().numpy()


def _fifo_medians(chunks, fifo_len, device):
    """Medians of a FIFO that keeps the last ``fifo_len`` values, read after every append of one of
    ``chunks`` (float32-valued arrays): np.median(concatenate(chunks[:j+1])[-fifo_len:]) for every j.
    Full-length states are equal-length windows of the concatenation -> one ddm_row_medians call."""
    if not chunks:
        return np.empty(0)
    t = _dev.torch()
    cat = np.concatenate(chunks)
    ends = np.cumsum([c.size for c in chunks])
    meds = np.empty(len(chunks))
    for j in np.nonzero(ends < fifo_len)[0]:
        meds[j] = np.median(cat[:ends[j]]) if ends[j] else np.nan
    full = np.nonzero(ends >= fifo_len)[0]
    if full.size:
        catd = t.from_numpy(cat.astype(np.float32)).to(device)
        meds[full] = _row_medians(catd, starts=ends[full] - fifo_len, length=fifo_len)
    return meds


def _resample_rows(x, starts, n, num):
    """signal.resample of the rows x[s:s+n] -> [rows][num] cuda float32 (ddm_resample_rows)."""
    t = _dev.torch()
    sd = t.as_tensor(np.ascontiguousarray(starts, dtype=np.int64), device=x.device)
    out = t.empty((sd.numel(), int(num)), dtype=t.float32, device=x.device)
    _lib.check(_lib.lib().ddm_resample_rows(fftops._context(x.device.index), _dev.ptr(x), _dev.ptr(sd),
                                            sd.numel(), int(n), int(num), _dev.ptr(out),
                                            _dev.stream_ptr(x.device.index)), "ddm_resample_rows")
    return out


def _steadiest(sync, rate, n):
    """min over all runs of n consecutive sync spacings of the run's largest deviation from half a
    second (decode_noaa.py:795-799; the reference's list comprehension of ~1800 np.max calls is 4 ms of
    interpreter time per sync word -- one windowed maximum here)."""
    dev = np.abs(np.diff(sync) - (rate * 0.5))
    if len(dev) - n + 1 <= 0:
        return np.min([])               # the reference's error for a pass with too few syncs
    return np.min(np.max(np.lib.stride_tricks.sliding_window_view(dev, n), axis=1))


def _quantise_rows(image, backup):
    """The pixel rows of decode_noaa.py:432-461 in one vectorised pass.  ``image`` holds (row, slope,
    intercept) for the lines written once a telemetry frame had been seen (:440-452: row * slope +
    intercept), ``backup`` (row, low, high) for the first-guess levels used when none ever was (:434-438,
    :455-456: 255 (row - low) / (high - low)); each row goes through round, clip to 0..255, uint8 exactly
    as the reference does line by line.  Every row has the same length here (shorter lines raise earlier),
    so the reference's "keep the rows of the most common length" (:459-461) keeps them all; its error on a
    pass without a single whole line is kept."""
    def quantise(v):
        v = np.round(v)
        v[v < 0] = 0
        v[v > 255] = 255
        return v.astype(np.uint8)
    if len(image) == 0 and len(backup) == 0:
        raise ValueError("max() arg is an empty sequence")
    if len(image) == 0:
        rows = np.stack([r for r, _, _ in backup])
        low = np.array([lo for _, lo, _ in backup], dtype=np.float64)[:, None]
        high = np.array([hi for _, _, hi in backup], dtype=np.float64)[:, None]
        return quantise(255 * (rows - low) / (high - low))
    rows = np.stack([r for r, _, _ in image])
    slope = np.array([a for _, a, _ in image], dtype=np.float64)[:, None]
    icpt = np.array([b for _, _, b in image], dtype=np.float64)[:, None]
    return quantise(rows * slope + icpt)


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
        self._image = None
        self._chIDA = None
        self._chIDB = None
        self._low = self._high = self._slope = self._intercept = None

    # ---- image assembly ---------------------------------------------------------------
    @property
    def channelID(self):
        """[channel id A, channel id B] from the telemetry wedges (decode_noaa.py:53-66)."""
        if self._image is None:
            self.getImage
        return [self._chIDA, self._chIDB]

    @staticmethod
    def _fillSync(csync, maxLen):
        """decode_noaa.py:467-508: keep the syncs spaced by the modal distance (+-200), then
        extrapolate backwards to the start and fill every gap forwards up to maxLen."""
        spacing = np.diff(csync)
        counts = list(spacing)
        mode = max(set(spacing), key=counts.count)
        wiggle = 200
        valid, seen = [], set()              # (a set beside the list: `v not in valid` alone is 65 ms per pass)
        for i in range(len(csync) - 1):
            if abs(csync[i + 1] - csync[i] - mode) < wiggle:
                for v in (csync[i], csync[i + 1]):
                    if v not in seen:
                        seen.add(v)
                        valid.append(v)
        filled = valid[:]
        c = valid[0] - mode
        while c > wiggle:
            filled.append(c)
            c -= mode
        anchor = 0
        c = mode
        while valid[anchor] + c < maxLen:
            if (anchor + 1) < len(valid) and (abs(valid[anchor + 1] - c - valid[anchor]) < wiggle
                                              or c + valid[anchor] > valid[anchor + 1]):
                anchor += 1
                c = mode
            else:
                filled.append(valid[anchor] + c)
                c += mode
        return list(np.sort(filled))

    @property
    def getImage(self):
        """Matrix of uint8 pixel values, one row per APT line (decode_noaa.py:255-465)."""
        if self._image is not None:
            return self._image
        if self._audOut is None or self._syncA is None or self._syncB is None:
            self.getCrudeSync()
        logging.info('Beginning image extraction')
        t = _dev.torch()
        aud = self._audOut
        fs = aud.sampRate
        # :274 whole-stream zero-phase band-pass.  Its float64 tf-form recursion has a 3e-4 roundoff
        # floor and its input already differs from the reference's by the fp32
        # rounding of the FM stage, so the sequential bit-replay buys nothing here: run parallel.
        bp = filters.butter(fs, 400, 4400, typeFlt=constants.FLT_BP, zeroPhase=True)
        bp.setIIRMode(1)
        aud.filter(bp)
        am = self._getAM(aud)
        sig = am.deviceSignal                       # cuda float32, stays on the device
        n = am.length
        csyncA = self._syncA / self._syncCrudeSampRate * fs
        csyncB = self._syncB / self._syncCrudeSampRate * fs
        ucsync = csyncA[:]
        csyncA = self._fillSync(csyncA, n)
        csyncB = self._fillSync(csyncB, n)
        if csyncB[0] < csyncA[0]:
            csyncB.pop(0)
        if csyncB[-1] < csyncA[-1]:
            csyncA.pop(-1)
        if not len(csyncA) == len(csyncB):
            logging.error("Number of syncA and syncB unequal")
            csyncB = np.array(csyncA) + int(0.25 * fs)

        numPixels = int(0.5 / constants.NOAA_T)      # 2080
        half = numPixels // 2
        # :317-320 first guess of the black / white levels from the whole pass
        seg = n // numPixels
        if seg < 1:
            # a pass shorter than one line: the reference takes medians of empty rows (NaN, with numpy's
            # "Mean of empty slice" warning) and percentiles of those
            self._low = self._high = float("nan")
        else:
            (self._low, self._high) = np.percentile(_row_medians(sig, rows=numPixels, stride=seg, length=seg),
                                                    (0.5, 99.5))

        # ---- geometry of every line, then all GPU work in batches --------------------
        lines = []
        for k in range(len(csyncA)):
            sA, sB = int(csyncA[k]), int(csyncB[k])
            eA = sB
            eB = sB + int(0.25 * fs)
            if 1 + k < len(csyncA):
                eB = int(csyncA[k + 1])
            if eB > n or eA > n or sA < 0 or sB < 0:
                continue
            lines.append((k, sA, eA, sB, eB))
        # per-pixel medians of the resampled half lines, grouped by (length, target) so that each
        # group is one batched FFT resample; plus the raw samples of the 40 sync pixels of side A
        pix = {}
        sync_px = {}
        groups = {}
        for li, (k, sA, eA, sB, eB) in enumerate(lines):
            for side, (s0, e0) in (("A", (sA, eA)), ("B", (sB, eB))):
                ln = e0 - s0
                num = int(int(ln / (numPixels * 0.5)) * (numPixels * 0.5))
                groups.setdefault((ln, num), []).append((li, side, s0))
        nsync = len(constants.NOAA_SYNCA)
        for (ln, num), members in groups.items():
            if ln < 1 or num < half:
                for li, side, s0 in members:
                    pix[(li, side)] = None
                continue
            rows = _resample_rows(sig, [m[2] for m in members], ln, num)        # [rows][num]
            per = num // half
            med = _row_medians(rows.reshape(-1), rows=rows.shape[0] * half, stride=per, length=per)
            med = med.reshape(rows.shape[0], half)
            head = rows[:, :nsync * per].cpu().numpy().astype(np.float64)
            for r, (li, side, s0) in enumerate(members):
                pix[(li, side)] = med[r]
                if side == "A":
                    sync_px[li] = head[r].reshape(nsync, per)
        # calibration strips in front of each sync (:370-378)
        lenStrip = int((len(constants.NOAA_SYNCA) * constants.NOAA_T) * fs)
        lenStrip2 = int((len(constants.NOAA_SYNCB) * constants.NOAA_T) * fs)

        def strip_medians(starts, ln):
            vals = np.full(len(starts), np.nan)
            ok = [i for i, s0 in enumerate(starts) if s0 - ln >= 0]
            if ok and ln > 0:
                vals[ok] = _row_medians(sig, starts=[starts[i] - ln for i in ok], length=ln)
            done = set(ok)                      # (`i not in ok` on the list is 16 ms per 1800-line pass)
            for i, s0 in enumerate(starts):
                if i not in done:               # negative start: numpy slice semantics (usually empty -> nan)
                    piece = sig[s0 - ln:s0].cpu().numpy()
                    vals[i] = np.median(piece) if piece.size else np.nan
            return vals
        stripA = strip_medians([l[1] for l in lines], lenStrip)
        stripB = strip_medians([l[3] for l in lines], lenStrip2)

        # ---- host: colour-calibration state machine over the precomputed lines -------
        image, imageBuffer, backupImage = [], [], []
        low_rows = [j for j in range(len(constants.NOAA_SYNCA)) if constants.NOAA_SYNCA[j] == 0]
        high_rows = [j for j in range(len(constants.NOAA_SYNCA)) if constants.NOAA_SYNCA[j] != 0]
        fifoLen = constants.NOAA_COLORCORRECT_FIFOLEN
        # :358-366 two FIFOs of the last 10 000 samples seen under the low / high sync bits, read
        # through their medians after every line with a real (not filled-in) sync.  All lines are
        # known here, so a FIFO state is a window of one concatenated array: the full-length windows
        # are equal-length rows of one ddm_row_medians call, the few partial ones at the start of the
        # pass stay np.median (appending row by row and truncating each time == appending all,
        # truncating once).
        real_sync = set(ucsync)
        upd = []
        for li, (k, sA, eA, sB, eB) in enumerate(lines):
            if pix.get((li, "A")) is None or pix.get((li, "B")) is None:
                break                           # the loop below raises at this line
            if csyncA[k] in real_sync:
                upd.append(li)

        def fifo_medians(rows_sel):
            meds = _fifo_medians([sync_px[li][rows_sel].ravel() for li in upd], fifoLen, sig.device)
            return dict(zip(upd, meds))
        lowMed, highMed = fifo_medians(low_rows), fifo_medians(high_rows)

        def median_small(vals):
            """np.median of a short python list (the three-entry smoothing FIFOs)."""
            for v in vals:
                if v != v:
                    return np.nan
            s = sorted(vals)
            m = len(s)
            return float(s[m // 2]) if m % 2 else 0.5 * (s[m // 2 - 1] + s[m // 2])
        corrfifo, corrfifosig, corrfifosig2 = [], [], []
        ncorrfifo = 3
        lcorr = lcorrsig = None
        statecorr = 0
        valuesPixCorr, valuesSigCorr = [], []
        chidFifo1, chidFifo2 = [], []

        for li, (k, sA, eA, sB, eB) in enumerate(lines):
            if pix.get((li, "A")) is None or pix.get((li, "B")) is None:
                raise ValueError("cannot reshape array of size 0 into shape (%d,0)" % half)
            if li in lowMed:
                val11, val244 = lowMed[li], highMed[li]
                self._low = val11 - (val244 - val11) * (11 - 0) / (244 - 11)
                self._high = val11 - (val244 - val11) * (11 - 255) / (244 - 11)
            stripVal = stripA[li]
            corrfifo = (corrfifo + [255 * (stripVal - self._low) / (self._high - self._low)])[-ncorrfifo:]
            outcorr = median_small(corrfifo)
            corrfifosig = (corrfifosig + [stripVal])[-ncorrfifo:]
            outcorrsig = median_small(corrfifosig)
            corrfifosig2 = (corrfifosig2 + [stripB[li]])[-ncorrfifo:]
            outcorrsig2 = median_small(corrfifosig2)
            chidFifo1 = (chidFifo1 + [outcorrsig2])[-100:]
            chidFifo2 = (chidFifo2 + [outcorrsig])[-100:]
            # telemetry wedge tracker (:386-425): eight rising steps then a big drop = one frame
            if lcorr is None or abs(outcorr - lcorr) > 255.0 / 16:
                if statecorr == 0 and lcorrsig is not None:
                    valuesPixCorr = [lcorr, outcorr]
                    valuesSigCorr = [lcorrsig, outcorrsig]
                    statecorr = 1
                elif 1 <= statecorr <= 6:
                    if outcorr - valuesPixCorr[-1] > 2 * 255.0 / (8 * 3):
                        valuesPixCorr.append(outcorr)
                        valuesSigCorr.append(outcorrsig)
                        statecorr += 1
                    else:
                        statecorr = 0
                elif statecorr == 7:
                    if valuesPixCorr[-1] - outcorr > 2 * 255.0 / 3:
                        valuesPixCorr = [outcorr] + valuesPixCorr
                        valuesSigCorr = [outcorrsig] + valuesSigCorr
                        from scipy import stats
                        self._slope, self._intercept = stats.linregress(
                            valuesSigCorr, np.arange(9) * 255.0 / 8)[:2]
                        if len(chidFifo1) > 1 + 64 + 8:
                            self._chIDA = int(np.round((self._slope * np.median(chidFifo1[-1 - 64 - 8:-1 - 64])
                                                        + self._intercept) / (255.0 / 8)))
                            self._chIDB = int(np.round((self._slope * np.median(chidFifo2[-1 - 64 - 8:-1 - 64])
                                                        + self._intercept) / (255.0 / 8)))
                        chidFifo1, chidFifo2 = [], []
                    statecorr = 0
            lcorr, lcorrsig = outcorr, outcorrsig
            row = np.concatenate([pix[(li, "A")], pix[(li, "B")]])
            # the rows are quantised in one vectorised pass after the loop (row by row it is 17 ms per
            # pass); what is recorded here is the row and the two numbers its formula uses
            if self._slope is None or self._intercept is None:
                imageBuffer.append(row)
                backupImage.append((row, self._low, self._high))
            else:
                for old in imageBuffer:
                    image.append((old, self._slope, self._intercept))
                imageBuffer = []
                image.append((row, self._slope, self._intercept))
        self._image = _quantise_rows(image, backupImage)
        logging.info('Image extraction complete')
        return self._image

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
            raw = getattr(self._sigsrc, "readRaw", None)       # 8-bit sources: bytes straight into the kernel
            block = raw(*i) if raw is not None else self._sigsrc.read(*i)
            sig = comm.commSignal(self._sigsrc.sampFreq, block, chunkerObj) \
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
            minSyncAdiff = _steadiest(self._syncA, self._syncCrudeSampRate, n)
            minSyncBdiff = _steadiest(self._syncB, self._syncCrudeSampRate, n)
            if minSyncAdiff < constants.NOAA_DETECTMAXCHANGE or minSyncBdiff < constants.NOAA_DETECTMAXCHANGE:
                logging.info('NOAA Signal was found')
                self._useful = 1
            else:
                logging.info('NOAA Signal was not found')
        return [self._syncA, self._syncB]

    def _accurate(self, csync, syncBits, width, useNormCorrelate, batch=256):
        """All windows of one sync kind, ``batch`` at a time through row-batched kernels: the
        windows are independent and equally long, so mixer, zero-phase filters, FM, Hilbert AM and
        correlation each run ONCE per batch instead of once per window (decode_noaa.py:844-877).
        With a single candidate group per window (window < minimum peak distance) the reference's
        threshold + group scan reduces to the first maximum of the correlation."""
        t = _dev.torch()
        l = _lib.lib()
        fs = self._sigsrc.sampFreq
        starts = [int(i) - int(width) for i in csync
                  if int(i) - int(width) >= 0 and int(i) + int(width) <= self._sigsrc.length]
        n = 2 * int(width)
        needle = sync.sync_needle(syncBits, fs, useNormCorrelate)
        m = len(needle)
        if not starts or n - 1 >= constants.NOAA_MINPEAKDIST * fs or n - 1 <= 3 * 492:
            return self._accurate_loop(csync, syncBits, width, useNormCorrelate)
        bh = filters.blackmanHarris(151, zeroPhase=True)
        ham = filters.hamming(492, zeroPhase=True)
        out, pk, tm = [], [], []
        raw = getattr(self._sigsrc, "readRaw", None)
        for b0 in range(0, len(starts), batch):
            rows_s = starts[b0:b0 + batch]
            w = len(rows_s)
            # ---- gather the windows into a [w][n] cf32 matrix on the device ----
            if raw is not None:
                u8 = t.from_numpy(np.stack([raw(s0, s0 + n).data for s0 in rows_s])).to("cuda")
                x = t.empty((w, n), dtype=t.complex64, device=u8.device)
                _lib.check(l.ddm_cu8_to_cf32(x.device.index, _dev.ptr(u8), w * n, _dev.ptr(x),
                                             _dev.stream_ptr(x.device.index)), "ddm_cu8_to_cf32")
            else:
                blocks = [self._sigsrc.read(s0, s0 + n) for s0 in rows_s]
                if _dev.is_tensor(blocks[0]):
                    x = t.stack([bk.to(t.complex64) for bk in blocks]).contiguous()
                else:
                    x = t.from_numpy(np.stack(blocks).astype(np.complex64)).to("cuda")
            dev = x.device.index
            st = _dev.stream_ptr(dev)
            _lib.check(l.ddm_mix_rows_cf32(dev, _dev.ptr(x), w, n, float(self._offset), float(fs), st),
                       "ddm_mix_rows_cf32")
            y = t.empty_like(x)
            _lib.check(l.ddm_filter_filtfilt_rows_dev(bh._handle(), _dev.ptr(x), w, n, 1, _dev.ptr(y), st),
                       "ddm_filter_filtfilt_rows_dev")
            fm = t.empty((w, n - 1), dtype=t.float32, device=x.device)
            _lib.check(l.ddm_fm_demod_rows(dev, _dev.ptr(y), w, n, _dev.ptr(fm), st), "ddm_fm_demod_rows")
            am = fftops.hilbert_envelope(fm.reshape(-1), n - 1).reshape(w, n - 1)       # one transform per row
            hay = t.empty_like(am)
            _lib.check(l.ddm_filter_filtfilt_rows_dev(ham._handle(), _dev.ptr(am), w, n - 1, 0, _dev.ptr(hay), st),
                       "ddm_filter_filtfilt_rows_dev")
            # rows separated by m zeros: one flat correlation serves all rows (windows never reach a neighbour)
            c = m // 2
            stride = (n - 1) + m
            padded = t.zeros((w, stride), dtype=t.float32, device=x.device)
            padded[:, c:c + n - 1] = hay
            cor = sync.correlate(padded.reshape(-1), needle, normalised=useNormCorrelate)
            idx = t.empty(w, dtype=t.int64, device=x.device)
            val = t.empty(w, dtype=t.float64, device=x.device)
            # output i of row r sits at flat index r*stride + i + c
            _lib.check(l.ddm_rows_argmax(dev, C.c_void_p(cor.data_ptr() + 8 * c), w, stride, n - 1, _dev.ptr(idx),
                                         _dev.ptr(val), st), "ddm_rows_argmax")
            idx_h = idx.cpu().numpy()
            val_h = val.cpu().numpy()
            det = idx_h - int(m / 2)                                   # start of the sync inside the window
            # np.average(sig[i + m : i + 2 m]) of the AM signal where it fits (decode_noaa.py:757-758)
            fits = [r for r in range(w) if det[r] + 2 * m < n - 1]
            means = {}
            if fits:
                so = t.as_tensor(np.array([r * (n - 1) + det[r] + m for r in fits], dtype=np.int64), device=x.device)
                mo = t.empty(len(fits), dtype=t.float64, device=x.device)
                _lib.check(l.ddm_rows_mean(dev, _dev.ptr(am), _dev.ptr(so), len(fits), m, _dev.ptr(mo), st),
                           "ddm_rows_mean")
                means = dict(zip(fits, mo.cpu().numpy().tolist()))
            for r in range(w):
                out.append(int(det[r]) + rows_s[r])
                pk.append(float(val_h[r]))
                tm.append(means.get(r))
        return out, pk, tm

    def _accurate_loop(self, csync, syncBits, width, useNormCorrelate):
        """The reference's per-window formulation, operator by operator (kept as the general
        path and as the cross-check of the batched one)."""
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
