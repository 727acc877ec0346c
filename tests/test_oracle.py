"""Pin the oracle: reference notebook vectors, committed golden fixtures (generated from the
unmodified reference by oracle/gen_golden.py) and the first-principles scipy restatement."""

import numpy as np
import pytest

from oracle import ddoracle as O
from oracle import scipy_restated as R

HALF_PI = np.pi / 2


# ---------------- reference notebook known-answer vectors ---------------------------
def test_exp3_rolling_average_stateless_whole():
    # Experiment 3 cell 9
    b, a = O.taps_rolling_average(2)
    got = O.filt_stateless(b, a, np.arange(1, 20))
    assert np.allclose(got, np.arange(0.5, 19.0, 1.0), atol=1e-12)


def test_exp3_rolling_average_stateless_chunks_border_error():
    # Experiment 3 cell 11
    b, a = O.taps_rolling_average(2)
    assert np.allclose(O.filt_stateless(b, a, [10, 11, 12, 13, 14]), [5, 10.5, 11.5, 12.5, 13.5])
    assert np.allclose(O.filt_stateless(b, a, [15, 16, 17, 18, 19]), [7.5, 15.5, 16.5, 17.5, 18.5])


def test_exp3_rolling_average_stateful_chunks():
    # Experiment 3 cell 19: first sample is 1.0 (unscaled lfilter_zi), rest equals whole
    b, a = O.taps_rolling_average(2)
    zi = O.initial_zi(b, a)
    y1, zi = O.filt_stateful(b, a, [1, 2, 3, 4, 5, 6, 7, 8, 9], zi)
    y2, zi = O.filt_stateful(b, a, [10, 11, 12, 13, 14], zi)
    y3, zi = O.filt_stateful(b, a, [15, 16, 17, 18, 19], zi)
    assert np.allclose(y1, [1.0, 1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 7.5, 8.5])
    assert np.allclose(y2, [9.5, 10.5, 11.5, 12.5, 13.5])
    assert np.allclose(y3, [14.5, 15.5, 16.5, 17.5, 18.5])


def test_exp5_fm_demod_vectors():
    a = np.array([1 + 1j, 2 - 2j, 3 + 3j, 4 - 4j, 5 + 5j, 6 - 6j])
    s = HALF_PI
    # cell 6 (stateless whole)
    assert np.allclose(O.fm_discriminator(a, store_state=False)[0], [-s, s, -s, s, -s])
    # cell 8 (stateless chunks lose one sample per chunk)
    assert np.allclose(O.fm_discriminator(a[:3], store_state=False)[0], [-s, s])
    assert np.allclose(O.fm_discriminator(a[3:], store_state=False)[0], [s, -s])
    # cell 10 (stateful: first chunk N-1, later chunks N)
    y1, last = O.fm_discriminator(a[:3], None)
    y2, last = O.fm_discriminator(a[3:], last)
    assert np.allclose(y1, [-s, s])
    assert np.allclose(y2, [-s, s, -s])


def test_exp6_decimation_with_chunker_carry():
    # cell 7: chunks of 10 with carried offset == unchunked ; cell 5: without carry it drifts
    x = np.arange(100)
    whole, rate, _ = O.decimate(x, 40, 10, 0)
    assert rate == 10 and np.array_equal(whole, np.arange(0, 100, 4))
    off = 0
    parts = []
    for a, b in O.chunk_bounds(100, 10):
        y, r, off = O.decimate(x[a:b], 40, 10, off)
        parts.append(y)
    assert np.array_equal(np.concatenate(parts), whole)
    nocarry = np.concatenate([O.decimate(x[a:b], 40, 10, 0)[0] for a, b in O.chunk_bounds(100, 10)])
    assert list(nocarry[:6]) == [0, 4, 8, 10, 14, 18]


def test_exp4e_remez_band_flattening_and_errors():
    b, a = O.taps_remez(2048000, [[0, 2e5], [3e5, 5e5]], [1, 0], ntaps=31)
    assert len(b) == 31 and a == [1]
    with pytest.raises(ValueError):
        O.taps_remez(2048000, [], [])
    with pytest.raises(ValueError):
        O.taps_remez(1000, [[0, 100], [200, 500]], [1, 0])
    with pytest.raises(ValueError):
        O.taps_remez(2048000, [[0, 2e5], [3e5, 5e5]], [1])
    with pytest.raises(ValueError):
        O.taps_butter(48000, 1000, kind=O.FLT_BP)
    with pytest.raises(ValueError):
        O.taps_butter(48000, 1000, kind=7)


def test_window_gains_are_unnormalised():
    # blackmanHarris(151) DC gain 53.81, hamming(492) 265.22
    assert abs(np.sum(O.taps_blackman_harris(151)[0]) - 53.81) < 0.01
    assert abs(np.sum(O.taps_hamming(492)[0]) - 265.22) < 0.01


# ---------------- golden fixtures generated from the imported reference -------------
def test_chunker_golden(golden):
    g = golden("chunker")
    for key in g.files:
        ln, sz = [int(t[1:]) for t in key.split("_")]
        want = g[key]
        got = np.array(O.chunk_bounds(ln, sz), dtype=np.int64)
        assert np.array_equal(got, want), key


@pytest.mark.parametrize("name", ["chain_noise_d34", "chain_fmtone_d34", "chain_fmtone_d68", "chain_noise_d50"])
def test_chain_golden(golden, name):
    g = golden(name)
    x, fs, f, bw = g["x"], int(g["fs"]), float(g["f_off"]), int(g["bw"])
    taps = O.taps_blackman_harris(151)[0]
    for tag, cs in (("whole", len(x) + 1), ("c2500", 2500), ("c1111", 1111), ("c97", 97)):
        fm, rate = O.chain_stream(x, fs, f, taps, bw, cs)
        assert rate == int(g["rate"])
        assert fm.shape == g["fm_" + tag].shape
        assert np.array_equal(fm, g["fm_" + tag]), tag      # same scipy calls -> bit equal
        iq, _ = O.chain_stream(x, fs, f, taps, bw, cs, demod=False)
        assert np.array_equal(iq, g["iq_" + tag])
    # the chain is chunk-invariant up to float64 rounding
    assert O.rel_rms(g["fm_c97"], g["fm_whole"]) < 1e-9


def test_mixer_golden(golden):
    g = golden("mixer")
    for tag in "abc":
        fs, f, n0 = g["p_" + tag]
        y, n1 = O.mix(g["x"], f, int(fs), int(n0))
        assert y.dtype == np.complex64
        assert n1 == int(n0) + len(y)
        assert np.array_equal(y, g["y_" + tag])


def test_filters_golden(golden):
    g = golden("filters")
    cuts = list(g["cuts"])
    taps = {
        "bh151": O.taps_blackman_harris(151), "ham492": O.taps_hamming(492),
        "gauss51": O.taps_gaussian(51, 5), "roll7": O.taps_rolling_average(7),
        "remez255": O.taps_remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=255),
        "butlp8": O.taps_butter(2400000, 100000, n=8),
        "butbp6": O.taps_butter(60235, 400, 4400, n=6, kind=O.FLT_BP),
        "buthp3": O.taps_butter(48000, 3000, n=3, kind=O.FLT_HP),
    }
    for k, (b, a) in taps.items():
        assert np.array_equal(np.asarray(b, dtype=np.float64), g[k + "_b"]), k
        assert np.array_equal(np.asarray(a, dtype=np.float64), g[k + "_a"]), k
        for tag, x in (("c", g["xc"]), ("r", g["xr"])):
            zi = O.initial_zi(b, a)
            parts = []
            for i in range(len(cuts) - 1):
                y, zi = O.filt_stateful(b, a, x[cuts[i]:cuts[i + 1]], zi)
                parts.append(y)
            assert np.array_equal(np.concatenate(parts), g["%s_%s" % (k, tag)]), (k, tag)
    assert np.array_equal(O.filt_zero_phase(*taps["bh151"], g["xr"]), g["bh151_zp_r"])
    assert np.array_equal(O.filt_zero_phase(*taps["ham492"], g["xc"]), g["ham492_zp_c"])
    assert np.array_equal(O.filt_zero_phase(*taps["butbp6"], g["xr"]), g["butbp6_zp_r"])
    assert np.array_equal(O.filt_stateless(*taps["ham492"], g["xr"]), g["ham492_sl_r"])
    assert np.array_equal(O.filt_stateless(*taps["butlp8"], g["xc"]), g["butlp8_sl_c"])


def test_demod_golden(golden):
    g = golden("demod")
    x, cuts = g["x"], list(g["cuts"])
    last = None
    parts = []
    for i in range(len(cuts) - 1):
        y, last = O.fm_discriminator(x[cuts[i]:cuts[i + 1]], last)
        parts.append(y)
    assert np.array_equal(np.concatenate(parts), g["fm_state"])
    assert np.array_equal(O.fm_discriminator(x, store_state=False)[0], g["fm_whole"])
    la = None
    parts = []
    for a, b in ((0, 3), (3, 700), (700, 5000)):
        y, la = O.fm_angle_diff(x[a:b], la)
        parts.append(y)
    assert np.array_equal(np.concatenate(parts), g["fmad_state"])
    a = g["am_x"]
    assert np.array_equal(O.am_envelope(a), g["am_even"])
    assert np.array_equal(O.am_envelope(a[:2187]), g["am_odd"])
    assert np.array_equal(O.am_envelope(a[:30]), g["am_small"])
    b, aa = O.taps_butter(20800, 1200)
    zi = O.initial_zi(b, aa)
    y1, zi = O.filt_stateful(b, aa, np.abs(a[:1000]), zi)
    y2, zi = O.filt_stateful(b, aa, np.abs(a[1000:]), zi)
    assert np.array_equal(np.concatenate([y1, y2]), g["amflt"])


def test_bwlim_golden(golden):
    g = golden("bwlim")
    x = g["x"]
    off = 0
    parts = []
    for a, b in O.chunk_bounds(len(x), 1000):
        y, rate, off = O.decimate(x[a:b], 2048000, 60000, off)
        parts.append(y)
        assert rate == 60235
    assert np.array_equal(np.concatenate(parts), g["dec34"])
    y, rate = O.resample_strict(x, 60235, 20800)
    assert rate == int(g["strict_rate"][0]) and np.array_equal(y, g["strict_even"])
    assert np.array_equal(O.resample_strict(x[:5800], 60235, 40960)[0], g["strict_b"])
    assert np.array_equal(O.resample_strict(x[:4801], 48000, 12000)[0], g["strict_c"])
    with pytest.raises(ValueError):
        O.decimate(x, 1000, 2000)
    with pytest.raises(ValueError):
        O.resample_strict(x, 1000, 2000)


# ---------------- first-principles restatement of the scipy routines ----------------
def test_restated_lfilter_matches_scipy_fir_and_iir():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(300) + 1j * rng.standard_normal(300)
    for b, a in (O.taps_blackman_harris(31), O.taps_butter(48000, 3000, n=4),
                 O.taps_butter(60235, 400, 4400, n=3, kind=O.FLT_BP)):
        zi = O.initial_zi(b, a)
        assert np.allclose(R.lfilter_zi(b, a), zi, rtol=1e-9, atol=1e-12)
        y0, z0 = O.filt_stateful(b, a, x, zi)
        y1, z1 = R.lfilter(b, a, x, zi)
        assert O.rel_rms(y1, y0) < 1e-12 and np.allclose(z1, z0, rtol=1e-8, atol=1e-10)
        assert O.rel_rms(R.lfilter(b, a, x.real), O.filt_stateless(b, a, x.real)) < 1e-12


def test_restated_lfiltic_filtfilt_hilbert_resample_correlate():
    import scipy.signal as sps
    rng = np.random.default_rng(2)
    x = rng.standard_normal(500)
    b, a = O.taps_butter(48000, 3000, n=4)
    assert np.allclose(R.lfiltic(b, a, x[:6], [0.5] * 4), sps.lfiltic(b, a, x[:6], [0.5] * 4))
    for bb, aa in ((b, a), O.taps_blackman_harris(31)):
        assert O.rel_rms(R.filtfilt(bb, aa, x), O.filt_zero_phase(bb, aa, x)) < 1e-10
    for n in (500, 499, 30):
        assert O.rel_rms(R.hilbert(x[:n]), sps.hilbert(x[:n])) < 1e-13
    for n, num in ((500, 172), (500, 173), (499, 200), (481, 37)):
        assert O.rel_rms(R.resample(x[:n], num), sps.resample(x[:n], num)) < 1e-12
    k = rng.standard_normal(40)
    assert O.rel_rms(R.correlate_same(x, k), sps.correlate(x, k, mode="same")) < 1e-12
    assert O.rel_rms(R.convolve_same(x, np.ones(40)), np.convolve(x, np.ones(40), mode="same")) < 1e-12
    k = rng.standard_normal(41)
    assert O.rel_rms(R.correlate_same(x, k), sps.correlate(x, k, mode="same")) < 1e-12


def test_ncc_window_and_peaks_synthetic():
    fs = 4160 * 4
    needle = O.sync_needle(O.NOAA_SYNCA, fs)
    assert len(needle) == 160
    rng = np.random.default_rng(3)
    sig = 0.3 + 0.02 * rng.standard_normal(fs * 3)
    starts = [1000 + k * (fs // 2) for k in range(5)]
    for s in starts:
        sig[s:s + 160] = needle
    peaks, cor = O.find_syncs(sig, fs, O.NOAA_SYNCA)
    assert list(peaks) == starts
    # brute-force definition of the normalised correlation at one interior index
    i = starts[2] + 80
    win = sig[i - 80:i + 80]
    want = np.dot(win, needle) / np.sqrt(np.dot(win, win) * np.dot(needle, needle))
    assert abs(cor[i] - want) < 1e-9


def test_afsk_bank_matches_loop_definition():
    rng = np.random.default_rng(4)
    x = rng.standard_normal(300)
    bw = 48000
    out = O.afsk_bank(x, bw)
    nbuf = 40
    assert np.all(out[-nbuf:] == 0)
    i = np.arange(nbuf)
    for s in (0, 17, 259):
        seg = x[s:s + nbuf]
        mi = np.sum(seg * np.cos((i / bw) / (1 / 1200) * 2 * np.pi))
        mq = np.sum(seg * np.sin((i / bw) / (1 / 1200) * 2 * np.pi))
        si = np.sum(seg * np.cos((i / bw) / (1 / 2200) * 2 * np.pi))
        sq = np.sum(seg * np.sin((i / bw) / (1 / 2200) * 2 * np.pi))
        assert abs(out[s] - (mi * mi + mq * mq - si * si - sq * sq)) < 1e-9


# ---------------- the restatement against scipy on drawn shapes (hypothesis) ----------------
try:
    from hypothesis import given, settings, strategies as st
    _HAVE_HYPOTHESIS = True
except Exception:                                             # pragma: no cover
    _HAVE_HYPOTHESIS = False

if _HAVE_HYPOTHESIS:
    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(n=st.integers(2, 700), num=st.integers(1, 900), seed=st.integers(0, 2**16))
    def test_restated_resample_and_hilbert_on_drawn_lengths(n, num, seed):
        """scipy.signal.resample (comm.py:114, decode_noaa.py:350-351: real signals, down AND up, even/odd
        on both sides) and scipy.signal.hilbert (demod_am.py:29) restated from their definitions, on drawn
        lengths."""
        import scipy.signal as sps
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(n)
        want = sps.resample(x, num)
        got = R.resample(x, num)
        assert got.shape == want.shape
        scale = max(np.sqrt(np.mean(np.abs(want) ** 2)), 1e-30)
        assert np.max(np.abs(got - want)) <= 1e-10 * max(scale, 1.0)
        assert O.rel_rms(R.hilbert(x), sps.hilbert(x)) < 1e-12

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(order=st.integers(1, 8), n=st.integers(40, 400), seed=st.integers(0, 2**16),
           kind=st.sampled_from(["lp", "hp", "bp"]), split=st.integers(1, 39))
    def test_restated_recursion_carries_state_across_any_split(order, n, seed, kind, split):
        """filters.py:64-70: lfilter with carried zi over two chunks == one call, for drawn Butterworth
        designs (filters.py:262-269), restated recursion against scipy's."""
        import scipy.signal as sps
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        wn = {"lp": 0.2, "hp": 0.3, "bp": [0.1, 0.35]}[kind]
        b, a = sps.butter(order, wn, btype={"lp": "low", "hp": "high", "bp": "band"}[kind])
        zi = sps.lfilter_zi(b, a)
        assert np.allclose(R.lfilter_zi(b, a), zi, rtol=1e-6, atol=1e-9)
        want, zf = sps.lfilter(b, a, x, zi=zi.astype(complex))
        y1, z1 = R.lfilter(b, a, x[:split], zi)
        y2, z2 = R.lfilter(b, a, x[split:], z1)
        got = np.concatenate([y1, y2])
        tol = 1e-9 if order * (2 if kind == "bp" else 1) <= 8 else 1e-6      # high-order tf forms: scipy's own floor
        assert O.rel_rms(got, want) < tol
        assert np.allclose(z2, zf, rtol=1e-5, atol=1e-7 * max(1.0, float(np.max(np.abs(zf)))))

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(n=st.integers(1, 300), m=st.integers(1, 120), seed=st.integers(0, 2**16))
    def test_restated_same_mode_windows_on_drawn_lengths(n, m, seed):
        """'same'-mode alignment of correlate / convolve (decode_noaa.py:671-672) for even and odd needle
        lengths, haystack at least as long as the needle (the reference's case)."""
        import scipy.signal as sps
        if m > n:
            n, m = m, n
        rng = np.random.default_rng(seed)
        h, k = rng.standard_normal(n), rng.standard_normal(m)
        assert np.allclose(R.correlate_same(h, k), sps.correlate(h, k, mode="same"), rtol=1e-10, atol=1e-10)
        assert np.allclose(R.convolve_same(h, k), np.convolve(h, k, mode="same"), rtol=1e-10, atol=1e-10)

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(order=st.integers(1, 6), n=st.integers(60, 400), seed=st.integers(0, 2**16), fir=st.booleans())
    def test_restated_zero_phase_filter_on_drawn_designs(order, n, seed, fir):
        """filters.py:73 filtfilt: odd extension by 3*max(len a, len b), both passes seeded with zi * edge."""
        import scipy.signal as sps
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(n)
        if fir:
            b, a = sps.windows.hamming(2 * order + 3), np.array([1.0])
        else:
            b, a = sps.butter(order, 0.25)
        if n <= 3 * max(len(a), len(b)):
            return                                   # scipy refuses inputs shorter than the padding
        assert O.rel_rms(R.filtfilt(b, a, x), sps.filtfilt(b, a, x)) < 1e-9


def test_noaa_sync_oracle_matches_unmodified_reference_fixture(golden):
    """The oracle's crude and accurate sync (the checker of tests/test_sync_gpu.py) against what the
    UNMODIFIED reference produced for the same synthetic pass (oracle/gen_golden_noaa.py ->
    tests/golden/noaa_pass.npz: getCrudeSync, decode_noaa.py:769-806, and getAccurateSync, :808-880):
    sync positions equal sample for sample.  Windows: the first two, the middle one and the last one of
    each sync word (0.4 s of CPU each)."""
    from tests.util import apt_iq, oracle_accurate_window
    g = golden("noaa_pass")
    fs, f_off = int(g["fs"]), float(g["f_off"])
    x = apt_iq(int(g["seed"]), float(g["seconds"]), fs=fs, f_off=f_off)
    assert abs(float(np.abs(x[::1000]).sum()) - float(g["input_checksum"][0])) < 1e-3, "generator drifted"
    taps = O.taps_blackman_harris(151)[0]
    audio, rate = O.chain_stream(x, fs, f_off, taps, 60000)
    am = O.am_envelope_chunked(audio)
    crude = {}
    for name, bits in (("A", O.NOAA_SYNCA), ("B", O.NOAA_SYNCB)):
        crude[name], _ = O.find_syncs(am, rate, bits)
        assert np.array_equal(crude[name], g["sync" + name]), name
    width = int(3 * O.NOAA_T * len(O.NOAA_SYNCA) * fs)
    for name, bits in (("A", O.NOAA_SYNCA), ("B", O.NOAA_SYNCB)):
        want = g["async" + name]
        starts = [int(c) - width for c in crude[name] / rate * fs
                  if int(c) - width >= 0 and int(c) + width <= len(x)]
        assert len(starts) == len(want), name                      # the windows the reference accepted (:828-830)
        for k in sorted({0, 1, len(starts) // 2, len(starts) - 1}):
            a = starts[k]
            assert oracle_accurate_window((x[a:a + 2 * width], a, fs, bits)) == int(want[k]), (name, k)
