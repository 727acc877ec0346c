"""GPU parity of the drop-in Python API (directdemod_b200.comm / filters / demod_fm / chunker)
against the golden fixtures produced by the UNMODIFIED reference (oracle/gen_golden.py).  The
test bodies mirror the generator line by line with ``directdemod_b200`` in place of
``directdemod`` -- that is the drop-in claim."""

import numpy as np
import pytest

from oracle import ddoracle as O
from tests.util import TOL, wrap_rel_rms

pytestmark = pytest.mark.gpu


class _Src:
    def __init__(self, n):
        self.length = n


def _mods():
    from directdemod_b200 import chunker, comm, constants, demod_fm, filters
    return chunker, comm, constants, demod_fm, filters


@pytest.mark.parametrize("name", ["chain_noise_d34", "chain_fmtone_d34", "chain_fmtone_d68", "chain_noise_d50"])
def test_fluent_chain_matches_reference(golden, name):
    chunker, comm, constants, demod_fm, filters = _mods()
    from directdemod_b200 import _lib
    g = golden(name)
    x, fs, f_off, bw = g["x"], int(g["fs"]), float(g["f_off"]), int(g["bw"])
    for tag, csize in (("whole", len(x) + 1), ("c2500", 2500), ("c1111", 1111), ("c97", 97)):
        ck = chunker.chunker(_Src(len(x)), csize)
        bh = filters.blackmanHarris(151)
        fm = demod_fm.demod_fm()
        out = comm.commSignal(1)
        launches0 = _lib.launch_count()
        for a, b in ck.getChunks:
            s = comm.commSignal(fs, x[a:b], ck).offsetFreq(f_off).filter(bh).bwLim(bw, uniq="First")
            s.funcApply(fm.demod)
            out.extend(s)
        # the whole chain must have run as ONE fused launch per chunk
        assert _lib.launch_count() - launches0 == len(ck.getChunks), tag
        assert out.sampRate == int(g["rate"])
        assert out.signal.dtype == np.float64
        assert out.signal.shape == g["fm_" + tag].shape, tag
        assert wrap_rel_rms(out.signal, g["fm_" + tag]) <= TOL, tag


def test_chain_iq_then_separate_fm_hands_state_over(golden):
    """gen_golden reads .signal between bwLim and funcApply: the fused IQ chain must hand its
    state to the stand-alone FM kernel and the results must still match."""
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("chain_fmtone_d34")
    x, fs, f_off, bw = g["x"], int(g["fs"]), float(g["f_off"]), int(g["bw"])
    ck = chunker.chunker(_Src(len(x)), 1111)
    bh = filters.blackmanHarris(151)
    fm = demod_fm.demod_fm()
    out = comm.commSignal(1)
    iq = []
    for a, b in ck.getChunks:
        s = comm.commSignal(fs, x[a:b], ck).offsetFreq(f_off).filter(bh).bwLim(bw, uniq="First")
        iq.append(np.array(s.signal))
        s.funcApply(fm.demod)
        out.extend(s)
    iq = np.concatenate(iq)
    assert iq.dtype == np.complex128 and iq.shape == g["iq_c1111"].shape
    assert O.rel_rms(iq, g["iq_c1111"]) <= TOL
    assert wrap_rel_rms(out.signal, g["fm_c1111"]) <= TOL


def test_fused_to_unfused_state_migration(golden):
    """First chunks through the fused kernel, later chunks operator by operator (a plain
    applyOn on the same filter object): the delay line and FM sample must carry over."""
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("chain_noise_d34")
    x, fs, f_off, bw = g["x"], int(g["fs"]), float(g["f_off"]), int(g["bw"])
    ck = chunker.chunker(_Src(len(x)), 2500)
    bh = filters.blackmanHarris(151)
    fm = demod_fm.demod_fm()
    parts = []
    for i, (a, b) in enumerate(ck.getChunks):
        if i < 2:
            s = comm.commSignal(fs, x[a:b], ck).offsetFreq(f_off).filter(bh).bwLim(bw, uniq="First")
            s.funcApply(fm.demod)
            parts.append(s.signal)
        else:
            s = comm.commSignal(fs, x[a:b], ck).offsetFreq(f_off)
            y = bh.applyOn(s.signal)                       # stand-alone FIR, numpy in/out
            s = comm.commSignal(fs, y, ck).bwLim(bw, uniq="First")
            parts.append(fm.demod(s.signal))
    got = np.concatenate(parts)
    assert got.shape == g["fm_c2500"].shape
    assert wrap_rel_rms(got, g["fm_c2500"]) <= TOL


def test_mixer_at_huge_global_index(golden):
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("mixer")
    for tag in "abc":
        fsx, f, n0 = g["p_" + tag]
        ck = chunker.chunker(_Src(10), 10)
        ck.set(constants.CHUNK_FREQOFFSET, int(n0))
        s = comm.commSignal(int(fsx), g["x"], ck).offsetFreq(float(f))
        assert ck.get(constants.CHUNK_FREQOFFSET) == int(n0) + len(g["x"])
        assert O.rel_rms(s.signal, g["y_" + tag]) <= TOL, tag


def test_mixer_per_sample_frequency_array():
    chunker, comm, constants, demod_fm, filters = _mods()
    rng = np.random.default_rng(3)
    n, fs = 5000, 2048000
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 30).astype(np.complex64)
    f = 30000.0 + 500.0 * np.sin(np.arange(n) / 300.0)
    ck = chunker.chunker(_Src(10), 10)
    ck.set(constants.CHUNK_FREQOFFSET, 123456789)
    got = comm.commSignal(fs, x, ck).offsetFreq(f).signal
    want = x * np.exp(-1.0j * 2.0 * np.pi * f * np.arange(123456789, 123456789 + n) / fs)
    assert O.rel_rms(got, want) <= TOL


FILTERS = {
    "bh151": lambda F, C: F.blackmanHarris(151),
    "ham492": lambda F, C: F.hamming(492),
    "gauss51": lambda F, C: F.gaussian(51, 5),
    "roll7": lambda F, C: F.rollingAverage(7),
    "remez255": lambda F, C: F.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=255),
    "butlp8": lambda F, C: F.butter(2400000, 100000, n=8),
    "butbp6": lambda F, C: F.butter(60235, 400, 4400, n=6, typeFlt=C.FLT_BP),
    "buthp3": lambda F, C: F.butter(48000, 3000, n=3, typeFlt=C.FLT_HP),
}


@pytest.mark.parametrize("name", sorted(FILTERS))
def test_stateful_filters_match_reference(golden, name):
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("filters")
    cuts = g["cuts"].tolist()
    for tag, x in (("c", g["xc"]), ("r", g["xr"])):
        f = FILTERS[name](filters, constants)
        np.testing.assert_allclose(np.asarray(f.getB, dtype=np.float64), g[name + "_b"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(np.asarray(f.getA, dtype=np.float64), g[name + "_a"], rtol=1e-12, atol=1e-15)
        got = np.concatenate([f.applyOn(x[cuts[i]:cuts[i + 1]]) for i in range(len(cuts) - 1)])
        want = g["%s_%s" % (name, tag)]
        assert got.dtype == want.dtype and got.shape == want.shape
        assert O.rel_rms(got, want) <= TOL, (name, tag, O.rel_rms(got, want))


def test_iir_roundoff_floor_and_execution_modes(golden):
    """scipy's float64 tf-form recursion has a filter-dependent roundoff floor; the library
    measures it, replays the loop sequentially (bit-exact) when it would show at 1e-5, and its
    segment-parallel mode agrees with scipy to that floor -- the same distance two scipy runs
    started at different samples keep from each other."""
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("filters")
    x = g["xr"].astype(np.float64)
    f = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP)
    is_fir, warm, floor = f.info()
    assert not is_fir and warm > 0 and 1e-5 < floor < 5e-3          # the NOAA band-pass: ~3e-4
    exact = f.applyOn(x)                                           # AUTO -> sequential replay
    import scipy.signal as sps
    b, a = g["butbp6_b"], g["butbp6_a"]
    ref, _ = sps.lfilter(b, a, x, zi=sps.lfilter_zi(b, a))
    assert np.array_equal(exact.astype(np.float32), ref.astype(np.float32))   # bit-exact before narrowing
    par = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP).setIIRMode(1)
    rng = np.random.default_rng(2)
    xl = rng.standard_normal(400000).astype(np.float32).astype(np.float64)
    refl, _ = sps.lfilter(b, a, xl, zi=sps.lfilter_zi(b, a))
    err = O.rel_rms(par.applyOn(xl), refl)
    assert err <= 10 * floor, (err, floor)
    # the AUTO switch is a user-settable tolerance: raised above this filter's floor, AUTO runs it
    # segment-parallel (same result as forcing mode 1); at the default it stays the bit-exact replay
    tolr = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP).setIIRTolerance(1e-2)
    y_tol = tolr.applyOn(xl)
    assert O.rel_rms(y_tol, refl) <= 10 * floor
    assert np.array_equal(y_tol, filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP).setIIRMode(1).applyOn(xl))
    assert tolr.lookback() > 0                                      # ... and may then be time-sharded
    back = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP).setIIRTolerance(1e-2).setIIRTolerance(0)
    assert np.array_equal(back.applyOn(x).astype(np.float32), ref.astype(np.float32))
    # a well-conditioned filter: parallel by default and within the plain tolerance
    f8 = filters.butter(2400000, 100000, n=8)
    assert f8.info()[2] < 1e-7
    b8, a8 = O.taps_butter(2400000, 100000, n=8)
    ref8, _ = sps.lfilter(b8, a8, xl, zi=sps.lfilter_zi(b8, a8))
    assert O.rel_rms(f8.applyOn(xl), ref8) <= TOL
    # the segment-parallel mode contracts the recursion into DFMAs: far below the tolerance, at the
    # filter's own roundoff floor; mode 3 keeps scipy's separately rounded operations in the segments
    y_fma = filters.butter(2400000, 100000, n=8).setIIRMode(1).applyOn(xl)
    y_sep = filters.butter(2400000, 100000, n=8).setIIRMode(3).applyOn(xl)
    assert O.rel_rms(y_fma, ref8) <= 1e-6 and O.rel_rms(y_sep, ref8) <= 1e-6
    assert np.array_equal(y_sep[:64], ref8[:64].astype(np.float32).astype(np.float64))   # first segment: scipy's own bits


def test_stateless_and_zero_phase_filters_match_reference(golden):
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("filters")
    xr, xc = g["xr"], g["xc"]
    for k, cls, n in (("bh151", filters.blackmanHarris, 151), ("ham492", filters.hamming, 492)):
        assert O.rel_rms(cls(n, zeroPhase=True).applyOn(xr), g[k + "_zp_r"]) <= TOL, k
        assert O.rel_rms(cls(n, zeroPhase=True).applyOn(xc), g[k + "_zp_c"]) <= TOL, k
        assert O.rel_rms(cls(n, storeState=False).applyOn(xr), g[k + "_sl_r"]) <= TOL, k
    got = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP, zeroPhase=True).applyOn(xr)
    assert O.rel_rms(got, g["butbp6_zp_r"]) <= TOL
    got = filters.butter(2400000, 100000, n=8, storeState=False).applyOn(xc)
    assert O.rel_rms(got, g["butlp8_sl_c"]) <= TOL
    with pytest.raises(ValueError):
        filters.hamming(492, zeroPhase=True).applyOn(xr[:1000])       # scipy: len must exceed padlen


def test_filter_constructor_errors_match_reference():
    chunker, comm, constants, demod_fm, filters = _mods()
    with pytest.raises(ValueError):
        filters.butter(48000, 1000, typeFlt=constants.FLT_BP)
    with pytest.raises(ValueError):
        filters.butter(48000, 1000, typeFlt=7)
    with pytest.raises(ValueError):
        filters.remez(48000, [], [])
    with pytest.raises(ValueError):
        filters.remez(48000, [[0, 1000], [2000, 24000]], [1, 0])
    with pytest.raises(ValueError):
        filters.remez(48000, [[0, 1000], [2000, 23000]], [1])


def test_exp3_rolling_average_golden_vectors():
    """Experiment 3 cells 9/11/19: storeState=True chunked == whole except the first sample,
    which is 1.0 (unscaled lfilter_zi)."""
    chunker, comm, constants, demod_fm, filters = _mods()
    x = np.arange(1, 20, dtype=np.float64)
    whole = filters.rollingAverage(2, storeState=False).applyOn(x)
    np.testing.assert_allclose(whole, np.arange(1, 20) - 0.5, atol=1e-6)
    f = filters.rollingAverage(2)
    got = np.concatenate([f.applyOn(x[:9]), f.applyOn(x[9:14]), f.applyOn(x[14:])])
    want = np.arange(1, 20) - 0.5
    want[0] = 1.0
    np.testing.assert_allclose(got, want, atol=1e-6)


def test_init_out_lfiltic_first_call():
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    rng = np.random.default_rng(11)
    x = rng.standard_normal(4000)
    b, a = sps.butter(4, 0.2)
    f = filters.filter(b, a, initOut=[0.3, -0.1, 0.2, 0.05])
    got = np.concatenate([f.applyOn(x[:1500]), f.applyOn(x[1500:])])
    want, _ = O.filt_initout(b, a, x, [0.3, -0.1, 0.2, 0.05])
    assert O.rel_rms(got, want) <= TOL


def test_fm_demod_golden_and_exp5_vectors(golden):
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("demod")
    x, cuts = g["x"], g["cuts"].tolist()
    fm = demod_fm.demod_fm()
    got = np.concatenate([fm.demod(x[cuts[i]:cuts[i + 1]]) for i in range(4)])
    assert got.shape == g["fm_state"].shape and wrap_rel_rms(got, g["fm_state"]) <= TOL
    got = demod_fm.demod_fm(storeState=False).demod(x)
    assert got.shape == g["fm_whole"].shape and wrap_rel_rms(got, g["fm_whole"]) <= TOL
    cuts2 = [0, 3, 700, 5000]
    fmad = demod_fm.demod_fmAD()
    got = np.concatenate([fmad.demod(x[cuts2[i]:cuts2[i + 1]]) for i in range(3)])
    assert got.shape == g["fmad_state"].shape and wrap_rel_rms(got, g["fmad_state"]) <= TOL
    # Experiment 5 cells 6/8/10
    v = np.array([1 + 1j, 2 - 2j, 3 + 3j, 4 - 4j, 5 + 5j, 6 - 6j])
    hp = np.pi / 2
    np.testing.assert_allclose(demod_fm.demod_fm(storeState=False).demod(v), [-hp, hp, -hp, hp, -hp], atol=1e-6)
    fm = demod_fm.demod_fm()
    np.testing.assert_allclose(fm.demod(v[:3]), [-hp, hp], atol=1e-6)
    np.testing.assert_allclose(fm.demod(v[3:]), [-hp, hp, -hp], atol=1e-6)


def test_bwlim_decimation_with_chunker_carry(golden):
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("bwlim")
    x = g["x"]
    ck = chunker.chunker(_Src(len(x)), 1000)
    parts = []
    for a0, b0 in ck.getChunks:
        s = comm.commSignal(2048000, x[a0:b0], ck).bwLim(60000, uniq="q")
        assert s.sampRate == 60235
        parts.append(s.signal)
    got = np.concatenate(parts)
    assert np.array_equal(got.astype(np.float32), g["dec34"])
    # Experiment 6 cells 5/7
    sig = np.arange(100, dtype=np.float64)
    ck = chunker.chunker(_Src(100), 10)
    got = np.concatenate([comm.commSignal(40, sig[a:b], ck).bwLim(10).signal for a, b in ck.getChunks])
    assert np.array_equal(got, np.arange(0, 100, 4))
    assert np.array_equal(comm.commSignal(40, sig).bwLim(10).signal, np.arange(0, 100, 4))


def test_commsignal_errors_and_bookkeeping():
    chunker, comm, constants, demod_fm, filters = _mods()
    with pytest.raises(ValueError):
        comm.commSignal(0, np.zeros(4))
    with pytest.raises(TypeError):
        comm.commSignal(10, np.zeros((4, 4)))
    s = comm.commSignal(48000.7, np.zeros(10, dtype=np.complex64))
    assert s.sampRate == 48000 and s.length == 10
    with pytest.raises(ValueError):
        s.bwLim(96000)
    with pytest.raises(TypeError):
        s.updateSignal(np.zeros((2, 2)))
    a = comm.commSignal(100, np.arange(5.0))
    b = comm.commSignal(200, np.arange(3.0))
    with pytest.raises(TypeError):
        a.extend(b)
    e = comm.commSignal(1)
    e.extend(b)
    assert e.sampRate == 200 and e.length == 3
    ck = chunker.chunker(_Src(25), 10)
    assert ck.getChunks == [[0, 10], [10, 20], [20, 25]]
    with pytest.raises(KeyError):
        ck.get("nope")
    assert ck.get("x", 5) == 5 and ck.get("x") == 5
    # foreign callables and foreign filter objects still work (numpy in, numpy out)
    s = comm.commSignal(100, np.arange(6.0)).funcApply(lambda v: v[::-1] * 2)
    assert np.array_equal(s.signal, np.arange(6.0)[::-1] * 2)

    class Twice:
        def applyOn(self, v):
            return np.asarray(v) * 2
    assert np.array_equal(comm.commSignal(100, np.arange(4.0)).filter(Twice()).signal, np.arange(4.0) * 2)


def test_blackman_harris_conv_same():
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    rng = np.random.default_rng(5)
    for n in (40, 151, 1000):
        x = rng.standard_normal(n)
        want = sps.convolve(x, sps.windows.blackmanharris(151), mode="same")
        assert O.rel_rms(filters.blackmanHarrisConv(151).applyOn(x), want) <= TOL, n


def test_fir_direct_and_overlap_save_fft_paths_agree_with_scipy():
    """Both FIR kernels (register-tiled direct convolution, overlap-save 4096-point FFT) against
    scipy for complex and real input, stateful over uneven chunks, incl. chunks shorter than a block."""
    chunker, comm, constants, demod_fm, filters = _mods()
    rng = np.random.default_rng(19)
    n = 50000
    xc = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    xr = rng.standard_normal(n).astype(np.float32)
    cuts = [0, 3000, 3001, 20000, n]
    for mk in (lambda: filters.blackmanHarris(151), lambda: filters.hamming(492),
               lambda: filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023),
               lambda: filters.gaussian(2049, 300)):
        for x in (xc, xr):
            b = np.asarray(mk().getB, dtype=np.float64)
            want, _ = O.filt_stateful(b, [1], x.astype(np.complex128 if np.iscomplexobj(x) else np.float64), O.initial_zi(b))
            for mode in (1, 2):
                f = mk().setFIRMode(mode)
                got = np.concatenate([f.applyOn(x[a:c]) for a, c in zip(cuts[:-1], cuts[1:])])
                assert O.rel_rms(got, want) <= TOL, (len(b), mode, O.rel_rms(got, want))
    with pytest.raises(Exception):
        filters.gaussian(2050, 300).setFIRMode(2)        # above the FFT path's tap limit


def test_median_filter_matches_scipy():
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    rng = np.random.default_rng(8)
    x = rng.standard_normal(10007).astype(np.float32)
    for k in (1, 3, 5, 21):
        got = filters.medianFilter(k).applyOn(x)
        assert np.array_equal(got.astype(np.float32), sps.medfilt(x, k).astype(np.float32)), k
    with pytest.raises(ValueError):
        filters.medianFilter(4).applyOn(x)


def test_long_fir_and_iir_large_chunks():
    """C4-shaped operators at a size that crosses many FIR/IIR tiles: 1023-tap Remez and an
    8th-order Butterworth on complex noise, chunked unevenly."""
    chunker, comm, constants, demod_fm, filters = _mods()
    rng = np.random.default_rng(17)
    n = 300000
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    cuts = [0, 100001, 100002, 180000, n]
    fr = filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    fb = filters.butter(2400000, 100000, n=8)
    got_r = np.concatenate([fr.applyOn(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    got_b = np.concatenate([fb.applyOn(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    b, a = O.taps_remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], 1023)
    want_r, _ = O.filt_stateful(b, a, x.astype(np.complex128), O.initial_zi(b, a))
    b2, a2 = O.taps_butter(2400000, 100000, n=8)
    want_b, _ = O.filt_stateful(b2, a2, x.astype(np.complex128), O.initial_zi(b2, a2))
    assert O.rel_rms(got_r, want_r) <= TOL, O.rel_rms(got_r, want_r)
    assert O.rel_rms(got_b, want_b) <= TOL, O.rel_rms(got_b, want_b)
    # poles close to the unit circle on a real signal: warm-up lengths of ~1e5 samples
    xr = rng.standard_normal(n)
    for cutoff in (300.0, 100.0):
        flp = filters.butter(2048000, cutoff, n=2)
        got = np.concatenate([flp.applyOn(xr[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        b3, a3 = O.taps_butter(2048000, cutoff, n=2)
        want, _ = O.filt_stateful(b3, a3, xr, O.initial_zi(b3, a3))
        assert O.rel_rms(got, want) <= TOL, (cutoff, O.rel_rms(got, want))


# ---------------------------------------------------------------------------------------
# FFT-defined operators: Hilbert envelope and strict (FFT) resampling
# ---------------------------------------------------------------------------------------
def test_am_demod_golden(golden):
    from directdemod_b200 import demod_am
    g = golden("demod")
    a = g["am_x"]
    am = demod_am.demod_am()
    for got, want in ((am.demod(a), g["am_even"]), (am.demod(a[:2187]), g["am_odd"]),
                      (am.demod(a[:30]), g["am_small"])):
        assert got.dtype == np.float64 and got.shape == want.shape
        assert O.rel_rms(got, want) <= TOL, O.rel_rms(got, want)
    amf = demod_am.demod_amFLT(20800, 1200)
    got = np.concatenate([amf.demod(a[:1000]), amf.demod(a[1000:])])
    assert O.rel_rms(got, g["amflt"]) <= TOL


def test_am_chunked_like_getAM_and_power_of_two():
    """decode_noaa.__getAM: per-240 000-sample Hilbert, last chunk shorter; plus lengths that
    take the direct power-of-two path."""
    from directdemod_b200 import demod_am
    rng = np.random.default_rng(21)
    n = 2 * 240000 + 101234
    t = np.arange(n) / 60235.0
    x = ((1 + 0.4 * np.sin(2 * np.pi * 7 * t)) * np.sin(2 * np.pi * 2400 * t)
         + 0.05 * rng.standard_normal(n)).astype(np.float32)
    got = demod_am.demod_am().demodChunked(x, O.AM_CHUNK)
    want = O.am_envelope_chunked(x.astype(np.float64), O.AM_CHUNK)
    assert got.shape == want.shape and O.rel_rms(got, want) <= TOL
    # exact multiple: the reference's chunker still ends with a full-size chunk
    x2 = x[:480000]
    assert O.rel_rms(demod_am.demod_am().demodChunked(x2, O.AM_CHUNK),
                     O.am_envelope_chunked(x2.astype(np.float64), O.AM_CHUNK)) <= TOL
    for m in (1, 2, 64, 4096, 65536):
        assert O.rel_rms(demod_am.demod_am().demod(x[:m]), O.am_envelope(x[:m].astype(np.float64))) <= TOL, m


def test_strict_bwlim_golden(golden):
    chunker, comm, constants, demod_fm, filters = _mods()
    g = golden("bwlim")
    x = g["x"]
    s = comm.commSignal(60235, x).bwLim(20800, True)
    assert s.sampRate == int(g["strict_rate"][0]) and s.length == len(g["strict_even"])
    assert O.rel_rms(s.signal, g["strict_even"]) <= TOL
    assert O.rel_rms(comm.commSignal(60235, x[:5800]).bwLim(40960, True).signal, g["strict_b"]) <= TOL
    assert O.rel_rms(comm.commSignal(48000, x[:4801]).bwLim(12000, True).signal, g["strict_c"]) <= TOL


def test_resample_all_branches_match_scipy():
    """Down/up, even/odd lengths, real and complex: the spectrum bookkeeping of
    scipy.signal.resample (Nyquist split/join) restated on the device."""
    import scipy.signal as sps
    from directdemod_b200 import _dev, fftops
    rng = np.random.default_rng(33)
    for n, num in ((1000, 400), (1000, 401), (1001, 400), (1001, 333), (400, 1000), (401, 1000),
                   (400, 1001), (30117, 2080), (588235, 203127), (512, 512), (512, 128), (7, 3)):
        xr = rng.standard_normal(n).astype(np.float32)
        got = _dev.to_host(fftops.resample(_dev.to_device(xr), num))
        want = sps.resample(xr.astype(np.float64), num)
        assert got.shape == want.shape and O.rel_rms(got, want) <= TOL, ("real", n, num, O.rel_rms(got, want))
        if n > 40000:
            continue
        xc = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        got = _dev.to_host(fftops.resample(_dev.to_device(xc), num))
        want = sps.resample(xc.astype(np.complex128), num)
        assert got.shape == want.shape and O.rel_rms(got, want) <= TOL, ("cplx", n, num, O.rel_rms(got, want))


def test_raw_8bit_source_blocks_take_the_fused_u8_path(tmp_path):
    """source.readRaw -> commSignal: same chain, same results as the complex64 route, the
    unsigned 8-bit bytes converted inside the fused kernel (one launch per chunk)."""
    chunker, comm, constants, demod_fm, filters = _mods()
    from directdemod_b200 import _lib, decode_fm, source
    rng = np.random.default_rng(13)
    n, fs = 300000, 2048000
    t = np.arange(n) / fs
    z = 70 * np.exp(1j * (2 * np.pi * 30000 * t + 2.0 * np.sin(2 * np.pi * 1300 * t))) \
        + 5 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    pairs = np.stack([np.clip(np.round(z.real + 127.5), 0, 255), np.clip(np.round(z.imag + 127.5), 0, 255)], 1).astype(np.uint8)
    dat = tmp_path / "iq.dat"
    pairs.tofile(dat)
    src = source.IQdat(str(dat), fs)
    outs = []
    for raw_mode in (False, True):
        ck = chunker.chunker(src, 70001)
        bh, fm = filters.blackmanHarris(151), demod_fm.demod_fm()
        out = comm.commSignal(1)
        l0 = _lib.launch_count()
        for a, b in ck.getChunks:
            block = src.readRaw(a, b) if raw_mode else src.read(a, b)
            out.extend(comm.commSignal(fs, block, ck).offsetFreq(30000).filter(bh).bwLim(60000, uniq="First")
                       .funcApply(fm.demod))
        assert _lib.launch_count() - l0 <= len(ck.getChunks) + 3       # raw: + the fresh-stream general-kernel piece
        outs.append(out.signal)
    x = src.read(0, n)
    want, _ = O.chain_stream(x, fs, 30000.0, O.taps_blackman_harris(151)[0], 60000, chunk_size=70001)
    for got in outs:
        assert got.shape == want.shape and wrap_rel_rms(got, want) <= TOL
    # raw block used outside the fusable pattern: converted on the device, .signal == reference read()
    assert np.array_equal(comm.commSignal(fs, src.readRaw(5, 1000)).signal, src.read(5, 1000))
    got = comm.commSignal(fs, src.readRaw(0, 5000)).offsetFreq(1000.0).signal
    assert O.rel_rms(got, O.mix(src.read(0, 5000), 1000.0, fs, 0)[0]) <= TOL
    # decode_fm.getAudio (decode_fm.py:41-72) on the 8-bit source: chain + strict resample per chunk
    aud = decode_fm.decode_fm(src, 30000.0).getAudio
    fmw, rate = O.chain_stream(x, fs, 30000.0, O.taps_blackman_harris(151)[0], 30000)
    want_a, _ = O.resample_strict(fmw, rate, 15000)
    assert aud.sampRate == 15000 and aud.length == len(want_a)
    assert O.rel_rms(aud.signal, want_a) <= TOL


def test_empty_and_tiny_inputs_behave_like_the_reference():
    """Edge cases the reference defines: empty arrays flow through filters / mixer / decimator,
    the stateful FM discriminator rejects an empty chunk, one-sample chunks carry state."""
    chunker, comm, constants, demod_fm, filters = _mods()
    from directdemod_b200 import demod_am
    e = np.zeros(0, dtype=np.complex64)
    f = filters.blackmanHarris(151)
    assert f.applyOn(e).shape == (0,)
    assert filters.butter(48000, 1000).applyOn(np.zeros(0)).shape == (0,)
    s = comm.commSignal(2048000, e).offsetFreq(1000.0).filter(filters.blackmanHarris(151)).bwLim(60000)
    assert s.length == 0 and s.signal.shape == (0,) and s.sampRate == 60235
    with pytest.raises(IndexError):
        demod_fm.demod_fm().demod(e)
    assert demod_fm.demod_fm(storeState=False).demod(e).shape == (0,)
    out = comm.commSignal(1)
    out.extend(comm.commSignal(100, np.zeros(0)))
    assert out.length == 0 and out.sampRate == 100
    # one-sample chunks through a stateful FIR + IIR + FM: same as one call
    rng = np.random.default_rng(23)
    x = (rng.standard_normal(40) + 1j * rng.standard_normal(40)).astype(np.complex64)
    f1, f2 = filters.hamming(7), filters.hamming(7)
    g1, g2 = filters.butter(48000, 3000, n=3), filters.butter(48000, 3000, n=3)
    d1, d2 = demod_fm.demod_fm(), demod_fm.demod_fm()
    one = d1.demod(g1.applyOn(f1.applyOn(x)))
    parts = [d2.demod(g2.applyOn(f2.applyOn(x[i:i + 1]))) for i in range(40)]
    assert O.rel_rms(np.concatenate(parts), one) <= TOL
    # AM of a single sample / two samples (scipy.hilbert of tiny arrays)
    for m in (1, 2, 3):
        v = np.arange(1.0, m + 1)
        assert O.rel_rms(demod_am.demod_am().demod(v), O.am_envelope(v)) <= TOL
    # non-contiguous and integer inputs are accepted like numpy accepts them
    y = filters.rollingAverage(4).applyOn(np.arange(40)[::2])
    want, _ = O.filt_stateful([0.25] * 4, [1], np.arange(40)[::2].astype(float), O.initial_zi([0.25] * 4))
    assert O.rel_rms(y, want) <= TOL


@pytest.mark.parametrize("cplx", [False, True])
def test_parallel_iir_ragged_lengths_and_chunk_carry(cplx):
    """The segment-parallel IIR over awkward lengths (shorter than a block, not a multiple of 16,
    shorter than the warm-up, several warps with a ragged last segment), chunk after chunk with the
    carried state, and from a sample-misaligned device view (the kernel without shared-memory
    staging): all equal scipy's lfilter on the whole stream."""
    import scipy.signal as sps
    import torch
    chunker, comm, constants, demod_fm, filters = _mods()
    rng = np.random.default_rng(11)
    sizes = [1, 5, 15, 16, 17, 63, 64, 65, 1000, 4097, 70001, 300000, 1234567, 3, 2048]
    n = sum(sizes)
    x = rng.standard_normal(n).astype(np.float32)
    x[100:260] = 0.0                     # zeros and denormals leave the integer-pipe float->double path
    x[5000] = np.float32(1e-40)
    x[1300000:1300040] = 0.0
    if cplx:
        x = (x + 1j * rng.standard_normal(n).astype(np.float32)).astype(np.complex64)
        x[200:300] = 0.0
    for make in (lambda: filters.butter(2400000, 100000, n=8), lambda: filters.butter(48000, 3000, n=3, typeFlt=constants.FLT_HP)):
        f = make()
        b, a = np.asarray(f.getB, dtype=np.float64), np.asarray(f.getA, dtype=np.float64)
        want, _ = sps.lfilter(b, a, x.astype(np.complex128 if cplx else np.float64), zi=sps.lfilter_zi(b, a))
        got, pos = [], 0
        for m in sizes:
            got.append(f.applyOn(x[pos:pos + m]))
            pos += m
        got = np.concatenate(got)
        assert got.shape == want.shape
        assert O.rel_rms(got, want) <= TOL
        assert np.max(np.abs(got - want)) <= 1e-4 * np.max(np.abs(want))
        # misaligned device input (one sample into an aligned buffer)
        f2 = make()
        xd = torch.from_numpy(x).cuda()
        got2 = f2.applyOn(xd[1:200001])
        got2 = got2.cpu().numpy() if hasattr(got2, "cpu") else np.asarray(got2)
        want2, _ = sps.lfilter(b, a, x[1:200001].astype(np.complex128 if cplx else np.float64), zi=sps.lfilter_zi(b, a))
        assert O.rel_rms(got2, want2) <= TOL


def test_c4_filters_at_slab_scale_chunk_invariance_and_windows():
    """BASELINE config 4 at a size where the production code paths run (480 M cf32 samples: the
    persistent overlap-save FIR; the warp-staged IIR with one complex recursion per lane, which small
    inputs never reach).  Size-independent properties: the whole slab in one call equals the same slab
    in 20 M-sample chunks with carried state (those take the lane-split IIR kernel), and windows of the
    output equal scipy's lfilter started W samples early."""
    import scipy.signal as sps
    import torch
    chunker, comm, constants, demod_fm, filters = _mods()
    n, chunk = 480_000_000, 20_000_000        # well above the size where the IIR stops splitting re/im
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).normal_(0, 40, generator=g)
    t = torch.arange(n, device="cuda", dtype=torch.float32)
    x += 300 * torch.polar(torch.ones_like(t), t * (2 * np.pi * 0.01))      # a carrier inside the pass band
    del t
    for make, lookback in ((lambda: filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023), 1022),
                           (lambda: filters.butter(2400000, 100000, n=8), 4000)):
        f_whole, f_chunks = make(), make()
        y = f_whole._apply_dev(x)
        worst = 0.0
        for a in range(0, n, chunk):
            yc = f_chunks._apply_dev(x[a:a + chunk])
            ref = y[a:a + chunk]
            worst = max(worst, float((yc - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt()))
            del yc
        assert worst <= 2e-6, worst                                       # same arithmetic up to fp32 block layout
        b, a_ = np.asarray(f_whole.getB, dtype=np.float64), np.asarray(f_whole.getA, dtype=np.float64)
        for start in (0, 123_456_789, n - 300_000):
            lo = max(0, start - lookback)
            seg = x[lo:start + 300_000].cpu().numpy().astype(np.complex128)
            if lo == 0:
                want, _ = sps.lfilter(b, a_, seg, zi=sps.lfilter_zi(b, a_))
            else:
                want = sps.lfilter(b, a_, seg)
            want = want[start - lo:]
            got = y[start:start + 300_000].cpu().numpy()
            assert O.rel_rms(got, want) <= TOL, (start, O.rel_rms(got, want))
        del y


def test_constructor_snapshot_of_large_host_arrays():
    """commSignal copies its input (comm.py:38).  Large float32 / complex64 host arrays are
    snapshotted straight onto the device: later writes to the caller's array must not show, .signal
    must hand back the same values in the same dtype while no operator has run, and operators must
    see the snapshot."""
    chunker, comm, constants, demod_fm, filters = _mods()
    rng = np.random.default_rng(21)
    n = 300000
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    keep = x.copy()
    s = comm.commSignal(2048000, x)
    x[:] = 0                                             # the caller reuses its buffer
    assert s.length == n
    got = s.signal
    assert got.dtype == np.complex64 and np.array_equal(got, keep)
    s2 = comm.commSignal(2048000, keep.copy())
    y = s2.filter(filters.blackmanHarris(151)).bwLim(60000).signal
    bh = filters.blackmanHarris(151)
    want = comm.commSignal(2048000, [complex(v) for v in keep[:70000]]).filter(bh).bwLim(60000).signal   # small/list input: host path
    assert y.dtype == np.complex128
    assert O.rel_rms(y[:len(want)], want) <= TOL
    xr = rng.standard_normal(n).astype(np.float32)
    sr = comm.commSignal(48000, xr)
    assert sr.signal.dtype == np.float32 and np.array_equal(sr.signal, xr)
    with pytest.raises(TypeError):
        comm.commSignal(48000, np.zeros((70000, 2), dtype=np.float32))


def test_zero_phase_iir_on_complex_and_float64_inputs():
    """filtfilt through the segment-parallel IIR for the sample formats the golden fixtures do not
    cover: complex input (float2 -> double2 forward pass, double2 -> double2 backward pass) and a
    well-conditioned low-pass, against scipy.signal.filtfilt."""
    import scipy.signal as sps
    chunker, comm, constants, demod_fm, filters = _mods()
    rng = np.random.default_rng(33)
    n = 300000
    xc = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 10).astype(np.complex64)
    xr = rng.standard_normal(n).astype(np.float32)
    for make in (lambda **kw: filters.butter(2400000, 100000, n=8, **kw), lambda **kw: filters.butter(20800, 1200, **kw)):
        f = make(zeroPhase=True)
        b, a = np.asarray(f.getB, dtype=np.float64), np.asarray(f.getA, dtype=np.float64)
        for x in (xc, xr):
            want = sps.filtfilt(b, a, x.astype(np.complex128 if np.iscomplexobj(x) else np.float64))
            got = make(zeroPhase=True).applyOn(x)
            assert got.shape == want.shape and np.iscomplexobj(got) == np.iscomplexobj(want)
            assert O.rel_rms(got, want) <= TOL, O.rel_rms(got, want)
        # stateless lfilter on complex input
        want = sps.lfilter(b, a, xc.astype(np.complex128))
        assert O.rel_rms(make(storeState=False).applyOn(xc), want) <= TOL


def test_queued_filter_then_direct_use_keeps_call_order():
    """commSignal.filter() only queues a stateful filter (it may fuse with a following bwLim); using
    the same filter object directly before the signal is read must still consume the delay line in
    call order, as the eagerly executing reference does (comm.py:91, filters.py:69)."""
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    rng = np.random.default_rng(21)
    x = (rng.standard_normal(6000) + 1j * rng.standard_normal(6000)).astype(np.complex64)
    b = sps.windows.blackmanharris(151)
    zi = sps.lfilter_zi(b, [1.0])
    w1, z1 = sps.lfilter(b, [1.0], x[:2000].astype(np.complex128), zi=zi)
    w2, z2 = sps.lfilter(b, [1.0], x[2000:4000].astype(np.complex128), zi=z1)
    w3, z3 = sps.lfilter(b, [1.0], x[4000:].astype(np.complex128), zi=z2)
    bh = filters.blackmanHarris(151)
    s = comm.commSignal(2048000, x[:2000]).filter(bh)          # queued, not yet run
    got2 = bh.applyOn(x[2000:4000])                            # must see the state AFTER the queued chunk
    assert O.rel_rms(got2, w2) <= TOL
    assert O.rel_rms(s.signal, w1) <= TOL
    s3 = comm.commSignal(2048000, x[4000:]).filter(bh)         # queued again ...
    st = bh.getState()                                         # ... reading the state runs it first
    assert O.rel_rms(st, z3) <= TOL
    assert O.rel_rms(s3.signal, w3) <= TOL


def test_fmad_result_does_not_depend_on_the_chunking():
    """demod_fmAD carries the last angle (demod_fm.py:88-94); cut anywhere, the stream must give the
    same samples (the carried value is re-evaluated in float64 from the last sample)."""
    chunker, comm, constants, demod_fm, filters = _mods()
    rng = np.random.default_rng(22)
    x = (rng.standard_normal(9000) + 1j * rng.standard_normal(9000)).astype(np.complex64)
    whole = demod_fm.demod_fmAD().demod(x)
    want, _ = O.fm_angle_diff(x, None)
    assert wrap_rel_rms(whole, want) <= TOL
    for cuts in ([0, 1, 2, 4500, 9000], [0, 3333, 3334, 8999, 9000]):
        ad = demod_fm.demod_fmAD()
        got = np.concatenate([ad.demod(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        assert np.array_equal(got, whole)


def test_filter_follows_the_device_of_its_input():
    """The native handle is created on the device of the tensor being filtered, not torch's
    current device; a filter with carried state refuses tensors from another device."""
    chunker, comm, constants, demod_fm, filters = _mods()
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(23)
    x = rng.standard_normal(5000).astype(np.float32)
    want = filters.hamming(31, storeState=False).applyOn(torch.from_numpy(x).cuda(0)).cpu().numpy()
    f = filters.hamming(31, storeState=False)
    with torch.cuda.device(0):
        got = f.applyOn(torch.from_numpy(x).to("cuda:1"))
    assert got.device.index == 1 and np.array_equal(got.cpu().numpy(), want)
    g = filters.hamming(31)
    g.applyOn(torch.from_numpy(x).cuda(0))
    with pytest.raises(ValueError):
        g.applyOn(torch.from_numpy(x).to("cuda:1"))


def test_cascade_runs_as_one_filter_and_matches_the_stage_by_stage_reference():
    """filters.cascade on the device: C4's remez-1023 -> butter-8 as ONE overlap-save pass, chunked
    with carried state, against scipy run stage by stage like the reference (filters.py:64-70)."""
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    from directdemod_b200 import _lib
    fs = 2400000
    rng = np.random.default_rng(41)
    n = 2500000
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    fir = filters.remez(fs, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    iir = filters.butter(fs, 100000, n=8)
    want = x.astype(np.complex128)
    for f in (fir, iir):
        want, _ = sps.lfilter(f.getB, f.getA, want, zi=sps.lfilter_zi(f.getB, f.getA))
    cas = filters.cascade([fir, iir])
    l0 = _lib.launch_count()
    got = np.concatenate([cas.applyOn(x[a:b]) for a, b in ((0, 1200000), (1200000, 1200700), (1200700, n))])
    assert got.shape == want.shape and O.rel_rms(got, want) <= TOL
    assert O.rel_rms(got[:3000], want[:3000]) <= TOL          # the stages' initial conditions
    assert _lib.launch_count() - l0 <= 8                       # one filter launch + one state launch per chunk
    # real signals too
    xr = rng.standard_normal(1300000).astype(np.float32)
    wr = xr.astype(np.float64)
    for f in (fir, iir):
        wr, _ = sps.lfilter(f.getB, f.getA, wr, zi=sps.lfilter_zi(f.getB, f.getA))
    assert O.rel_rms(filters.cascade([fir, iir]).applyOn(xr), wr) <= TOL


def test_commsignal_runs_consecutive_filters_as_one_cascade_and_hands_state_back():
    """``sig.filter(fir).filter(iir)`` (comm.py:80-92 twice) on long chunks runs as ONE equivalent
    overlap-save filter; the stages' states move into it in mid-stream and come back when a later,
    short chunk takes the stage-by-stage kernels again.  Reference: scipy stage by stage, state
    carried across all chunks (filters.py:64-70)."""
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    from directdemod_b200 import _lib
    fs = 2400000
    rng = np.random.default_rng(43)
    cuts = [0, 3000, 1303000, 2603000, 2604500, 3904500]       # short, long, long, short, long
    n = cuts[-1]
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    fir = filters.remez(fs, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    iir = filters.butter(fs, 100000, n=8)
    want = x.astype(np.complex128)
    for f in (fir, iir):
        want, _ = sps.lfilter(f.getB, f.getA, want, zi=sps.lfilter_zi(f.getB, f.getA))
    parts, launches = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        l0 = _lib.launch_count()
        s = comm.commSignal(fs, x[a:b]).filter(fir).filter(iir)
        parts.append(s.signal)
        launches.append(_lib.launch_count() - l0)
    got = np.concatenate(parts)
    assert got.shape == want.shape
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert O.rel_rms(got[a:b], want[a:b]) <= TOL, (a, b, O.rel_rms(got[a:b], want[a:b]))
    assert launches[1] <= 2 and launches[2] <= 2                # one filter launch (+ its state update)
    assert launches[0] >= 3 and launches[3] >= 3                # short chunks: stage by stage
    # the stages can be read on their own afterwards (state handed back) and still agree with scipy
    _, z1 = sps.lfilter(fir.getB, [1.0], x.astype(np.complex128), zi=sps.lfilter_zi(fir.getB, [1.0]))
    assert O.rel_rms(fir.getState(), z1) <= TOL


def test_recursive_filter_takes_its_fir_form_for_complex_chunks_and_returns_to_the_recursion():
    """An 8th-order Butterworth on complex chunks of 2^20 .. 2^26 samples runs as its equivalent FIR
    through the overlap-save kernel; shorter chunks and state reads go back to the recursion.  The
    stream must match scipy's stateful lfilter across the switches (filters.py:64-70)."""
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    fs = 2400000
    rng = np.random.default_rng(47)
    cuts = [0, 5000, 1205000, 2405000, 2406000, 3606000]
    n = cuts[-1]
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    f = filters.butter(fs, 100000, n=8)
    want, zf = sps.lfilter(f.getB, f.getA, x.astype(np.complex128), zi=sps.lfilter_zi(f.getB, f.getA))
    parts, forms = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        parts.append(f.applyOn(x[a:b]))
        forms.append(f.__dict__.get("_cascade") is not None)
    got = np.concatenate(parts)
    assert forms == [False, True, True, False, True]
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert O.rel_rms(got[a:b], want[a:b]) <= TOL, (a, b, O.rel_rms(got[a:b], want[a:b]))
    assert O.rel_rms(f.getState(), zf) <= 1e-5
    # real signals keep the recursion
    g = filters.butter(fs, 100000, n=8)
    g.applyOn(rng.standard_normal(1300000).astype(np.float32))
    assert g.__dict__.get("_cascade") is None


def test_a_state_set_by_hand_survives_the_fused_paths():
    """setState on a filter that has not run yet: the fused chain and the equivalent-FIR form must
    start from THAT delay line, not from the reference's lfilter_zi."""
    chunker, comm, constants, demod_fm, filters = _mods()
    import scipy.signal as sps
    rng = np.random.default_rng(53)
    # stateful FIR in front of a decimator (the fusable pattern)
    x = (rng.standard_normal(40000) + 1j * rng.standard_normal(40000)).astype(np.complex64)
    b = sps.windows.blackmanharris(151)
    zi = (rng.standard_normal(150) + 1j * rng.standard_normal(150))
    bh = filters.blackmanHarris(151)
    bh.setState(zi)
    got = comm.commSignal(2048000, x).filter(bh).bwLim(60000).signal
    want, _ = sps.lfilter(b, [1.0], x.astype(np.complex128), zi=zi)
    assert O.rel_rms(got, want[::34]) <= TOL
    # recursive filter on a long complex chunk (equivalent-FIR form)
    fs, n = 2400000, 1200000
    xl = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    f = filters.butter(fs, 100000, n=8)
    z0 = (rng.standard_normal(8) + 1j * rng.standard_normal(8)) * 5
    f.setState(z0)
    w, _ = sps.lfilter(f.getB, f.getA, xl.astype(np.complex128), zi=z0)
    assert O.rel_rms(f.applyOn(xl), w) <= TOL
