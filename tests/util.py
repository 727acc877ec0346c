"""Helpers shared by the parity tests."""

import numpy as np

from oracle import ddoracle as O

# Tolerance of BASELINE.json's north_star: fp32 compute vs the reference's float64 scipy
# path, relative RMS error <= 1e-5.
TOL = 1e-5


def wrap_rel_rms(got, want):
    """Relative RMS error of phase-valued data, compared modulo 2*pi (an FM sample sitting
    at +-pi may legitimately land on either side)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.angle(np.exp(1j * (got - want)))
    den = np.sqrt(np.mean(want ** 2))
    return float(np.sqrt(np.mean(err ** 2)) / den) if den > 0 else float(np.sqrt(np.mean(err ** 2)))


def noise_c64(seed, n, scale=40.0):
    rng = np.random.default_rng(seed)
    return ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * scale).astype(np.complex64)


def fm_tone_c64(seed, n, fs, f_off, f_mod, beta, amp=60.0, noise=2.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    ph = 2 * np.pi * f_off * t + beta * np.sin(2 * np.pi * f_mod * t)
    x = amp * np.exp(1j * ph) + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


def random_cuts(seed, n, k, small=0):
    """k random chunk boundaries over [0, n], optionally with a few tiny chunks mixed in."""
    rng = np.random.default_rng(seed)
    cuts = set(rng.integers(1, n, size=k).tolist()) if n > 1 else set()
    base = sorted(cuts)
    for c in base[:small]:
        for d in (1, 2, 3):
            if c + d < n:
                cuts.add(c + d)
    return [0] + sorted(cuts) + [n]


def oracle_chain(x, fs, f, taps, target, cuts, demod=True):
    st = O.ChainState(taps)
    parts = []
    rate = None
    for a, b in zip(cuts[:-1], cuts[1:]):
        y, rate = O.chain_chunk(x[a:b], fs, f, taps, target, st, demod=False)
        # The reference raises IndexError when a chunk decimates to nothing (demod_fm.py:44
        # indexes sig[-1]); the library defines that case as "no output, state unchanged",
        # which is what chunk invariance demands, so the checker skips the call there.
        if demod and len(y) > 0:
            y, st.fm_last = O.fm_discriminator(y, st.fm_last)
        elif demod:
            y = np.zeros(0)
        parts.append(y)
    return np.concatenate(parts), rate


def apt_iq(seed, seconds, fs=2048000, f_off=30000.0, dev_hz=17000.0, amp=60.0, noise=3.0):
    """Synthetic NOAA APT pass as complex64 IQ: 2 lines/s of 2080 words at 4160 words/s
    (sync A, space, image A, telemetry, sync B, space, image B, telemetry), amplitude-modulated
    on a 2400 Hz subcarrier, FM-modulated (+-dev_hz) on a carrier f_off above the tuner, AWGN."""
    rng = np.random.default_rng(seed)
    words_per_line = 2080
    n_lines = int(np.ceil(seconds * 2)) + 1
    sync_a = (np.array(O.NOAA_SYNCA[:39]) * 233 + 11)
    sync_b = (np.array(O.NOAA_SYNCB[:39]) * 233 + 11)
    lines = []
    for ln in range(n_lines):
        img_a = (128 + 100 * np.sin(np.arange(909) / 30.0 + ln / 5.0) + rng.integers(-10, 10, 909)).clip(0, 255)
        img_b = (100 + 80 * np.cos(np.arange(909) / 50.0 - ln / 7.0) + rng.integers(-10, 10, 909)).clip(0, 255)
        tel = np.full(45, 30 + 25 * ((ln // 8) % 8))
        lines.append(np.concatenate([sync_a, np.full(47, 11), img_a, tel, sync_b, np.full(47, 244), img_b, tel]))
    words = np.concatenate(lines).astype(np.float64) / 255.0
    assert len(lines[0]) == words_per_line
    n = int(seconds * fs)
    t = np.arange(n) / fs
    widx = np.minimum((t * 4160).astype(np.int64), len(words) - 1)
    audio = words[widx] * np.cos(2 * np.pi * 2400 * t)
    phase = 2 * np.pi * f_off * t + 2 * np.pi * dev_hz * np.cumsum(audio) / fs
    x = amp * np.exp(1j * phase)
    x += noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


class ArraySource:
    """Minimal IQ source with the reference's source interface (source.py:18-47): sampFreq,
    length, read(a, b)."""

    def __init__(self, x, fs):
        self._x = x
        self.sampFreq = fs
        self.length = len(x)

    def read(self, a, b=None):
        return self._x[a:b]


def afsk_wav_u8(path, seed=2, seconds=1.5, fs=2048000, dev_hz=3000.0):
    """C1 stand-in for samples/SDRSharp_..._IQ_autogain.wav (BASELINE configs[0]): a
    two-channel unsigned 8-bit WAV with the canonical 44-byte header holding AFSK 1200/2200 Hz
    FM-modulated at 0 Hz offset.  Written with scipy.io.wavfile, read back by source.IQwav."""
    import scipy.io.wavfile
    rng = np.random.default_rng(seed)
    n = int(seconds * fs)
    t = np.arange(n) / fs
    bits = rng.integers(0, 2, int(seconds * 1200) + 2)
    tone = np.where(bits[(t * 1200).astype(np.int64)] == 1, 1200.0, 2200.0)
    audio = np.sin(2 * np.pi * np.cumsum(tone) / fs)
    z = 90.0 * np.exp(1j * 2 * np.pi * dev_hz * np.cumsum(audio) / fs) + 4.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.stack([np.clip(np.rint(z.real + 127.5), 0, 255), np.clip(np.rint(z.imag + 127.5), 0, 255)], axis=1)
    scipy.io.wavfile.write(path, fs, iq.astype(np.uint8))
    return n


def tutorial_sequences(mods, wav_path, chunk_size):
    """The operator sequences of tutorial/1_fm.py:21-37, 2_filter.py:21-48 and 3_chunking.py:30-38,
    parametrised by the package (the reference's ``directdemod`` or ``directdemod_b200``).  The third
    one passes ``chunk_size`` so that a small file still exercises several chunks."""
    source, comm, chunker, constants, filters, demod_fm = mods
    out = {}
    # tutorial 1
    sigsrc = source.IQwav(wav_path)
    sig = comm.commSignal(sigsrc.sampFreq, sigsrc.read(0, sigsrc.length))
    sig.bwLim(30000)
    sig.funcApply(demod_fm.demod_fm().demod)
    out["t1"], out["t1_rate"] = np.asarray(sig.signal), sig.sampRate
    # tutorial 2
    sigsrc = source.IQwav(wav_path)
    sig = comm.commSignal(sigsrc.sampFreq, sigsrc.read(0, sigsrc.length))
    sig.filter(filters.blackmanHarris(151))
    sig.bwLim(30000)
    sig.funcApply(demod_fm.demod_fm().demod)
    out["t2_fm"] = np.asarray(sig.signal).copy()
    bFilter = filters.butter(sig.sampRate, 1200 - 1000, 2200 + 1000, typeFlt=constants.FLT_BP)
    sig.filter(bFilter)
    out["t2"], out["t2_rate"] = np.asarray(sig.signal), sig.sampRate
    # tutorial 3
    sigsrc = source.IQwav(wav_path)
    sigOut = comm.commSignal(sigsrc.sampFreq)
    bhFilter = filters.blackmanHarris(151)
    fmDemdulator = demod_fm.demod_fm()
    chunkerObj = chunker.chunker(sigsrc, chunk_size)
    for i in chunkerObj.getChunks:
        sig = comm.commSignal(sigsrc.sampFreq, sigsrc.read(*i), chunkerObj)
        sig.filter(bhFilter)
        sig.bwLim(30000)
        sig.funcApply(fmDemdulator.demod)
        sigOut.extend(sig)
    out["t3"], out["t3_rate"], out["t3_chunks"] = np.asarray(sigOut.signal), sigOut.sampRate, len(chunkerObj.getChunks)
    return out


def apt_iq_long(seed, seconds, fs=2048000, f_off=30000.0, dev_hz=17000.0, amp=60.0, noise=3.0, block_s=10.0):
    """apt_iq for long passes: generated block by block (carried phase) so that a 120 s pass needs
    2 GB for the result and ~1 GB of temporaries instead of 16 GB."""
    rng = np.random.default_rng(seed)
    n_lines = int(np.ceil(seconds * 2)) + 1
    sync_a = (np.array(O.NOAA_SYNCA[:39]) * 233 + 11)
    sync_b = (np.array(O.NOAA_SYNCB[:39]) * 233 + 11)
    lines = []
    for ln in range(n_lines):
        img_a = (128 + 100 * np.sin(np.arange(909) / 30.0 + ln / 5.0) + rng.integers(-10, 10, 909)).clip(0, 255)
        img_b = (100 + 80 * np.cos(np.arange(909) / 50.0 - ln / 7.0) + rng.integers(-10, 10, 909)).clip(0, 255)
        tel = np.full(45, 30 + 25 * ((ln // 8) % 8))
        lines.append(np.concatenate([sync_a, np.full(47, 11), img_a, tel, sync_b, np.full(47, 244), img_b, tel]))
    words = np.concatenate(lines).astype(np.float64) / 255.0
    n = int(seconds * fs)
    out = np.empty(n, dtype=np.complex64)
    blk = int(block_s * fs)
    acc = 0.0
    for a in range(0, n, blk):
        b = min(n, a + blk)
        idx = np.arange(a, b)
        t = idx / fs
        audio = words[np.minimum((t * 4160).astype(np.int64), len(words) - 1)] * np.cos(2 * np.pi * 2400 * t)
        cs = acc + np.cumsum(audio)
        acc = float(cs[-1])
        phase = 2 * np.pi * ((f_off * idx % fs) / fs) + 2 * np.pi * dev_hz * cs / fs
        z = amp * np.exp(1j * phase)
        z += noise * (rng.standard_normal(b - a) + 1j * rng.standard_normal(b - a))
        out[a:b] = z
    return out


def oracle_accurate_window(job):
    """One window of decode_noaa.getAccurateSync (decode_noaa.py:826-856) through the oracle
    primitives; top-level so that a multiprocessing pool can run many of them."""
    w, a, fs, bits = job
    taps = O.taps_blackman_harris(151)[0]
    ham = O.taps_hamming(492)[0]
    w, _ = O.mix(w, 30000.0, fs, 0)
    w = O.filt_zero_phase(taps, [1], w)
    w, _ = O.fm_discriminator(w, None, store_state=True)
    w = O.am_envelope(w)
    pk, _ = O.find_syncs(w, fs, bits, prefilter_taps=ham)
    return int(pk[0] + a)
