"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the public
header declares, the ctypes table mirrors the header, and the pure host logic of the drop-in
API (chunker, constants, argument validation) behaves like the reference."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ddemod.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_a_sane_api():
    syms = declared_symbols()
    assert len(syms) >= 40
    for must in ("ddm_chain_apply_dev", "ddm_chain_apply_host", "ddm_filter_apply_dev", "ddm_am_hilbert",
                 "ddm_resample", "ddm_correlate", "ddm_fm_demod", "ddm_mix_cf32", "ddm_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from directdemod_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, missing


def test_ctypes_table_mirrors_the_header():
    from directdemod_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.lib()                      # binds restype/argtypes of every entry
    assert lib.ddm_version() >= 100
    n = ctypes.c_int(-1)
    assert lib.ddm_device_count(ctypes.byref(n)) == 0 and n.value >= 0


def test_lfilter_zi_restatement_matches_scipy():
    import scipy.signal as sps
    from directdemod_b200 import _lib
    lib = _lib.lib()
    # (the 12th-order band-pass is left out: its zi is a catastrophic cancellation, scipy's own
    # float64 solve is only good to ~3e-4 there, and the Python layer hands scipy's very bits to
    # the library through ddm_filter_set_zi_base for exactly that reason)
    for b, a in (sps.butter(8, 0.0833), sps.butter(4, 0.2),
                 (sps.windows.hamming(492), [1.0]), ([0.5, 0.5], [1.0]), sps.butter(3, 0.125, btype="highpass")):
        b = np.ascontiguousarray(b, dtype=np.float64)
        a = np.ascontiguousarray(a, dtype=np.float64)
        zi = np.zeros(max(len(a), len(b)) - 1)
        rc = lib.ddm_lfilter_zi(b.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(b),
                                a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(a),
                                zi.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert rc == 0
        np.testing.assert_allclose(zi, sps.lfilter_zi(b, a), rtol=1e-9, atol=1e-12)


def test_group_peaks_host_logic_matches_reference_scan():
    from directdemod_b200 import _lib
    from oracle import ddoracle as O
    rng = np.random.default_rng(4)
    fs, n = 1000, 30000
    cor = 0.05 * rng.standard_normal(n)
    for p in range(250, n, 500):
        cor[p - 2:p + 3] += np.array([0.4, 0.8, 1.0, 0.8, 0.4])
    cor[5250] = cor[5251] = 2.0
    want = O.pick_sync_peaks(cor, fs, 0)
    expected = int(2 * (n / fs)) + 2
    s = np.sort(cor)
    thr = s[-expected:].sum() / expected
    thr -= O.NOAA_PEAKHEIGHTWIGGLE * (thr - s[:expected].sum() / expected)
    idx = np.ascontiguousarray(np.argwhere(cor > thr).ravel().astype(np.int64))
    val = np.ascontiguousarray(cor[idx])
    out = np.zeros(len(idx), dtype=np.int64)
    cnt = ctypes.c_int64()
    rc = _lib.lib().ddm_group_peaks(idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                    val.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(idx),
                                    O.NOAA_MINPEAKDIST * fs, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                    len(out), ctypes.byref(cnt))
    assert rc == 0
    assert np.array_equal(np.sort(out[:cnt.value]), want)


def test_chunker_matches_reference_golden(golden):
    from directdemod_b200 import chunker

    class Src:
        def __init__(self, n):
            self.length = n
    g = golden("chunker")
    for key in g.files:
        ln, sz = (int(v[1:]) for v in key.split("_"))
        assert np.array_equal(np.array(chunker.chunker(Src(ln), sz).getChunks, dtype=np.int64), g[key]), key
    ck = chunker.chunker(Src(25), 10)
    with pytest.raises(KeyError):
        ck.get("missing")
    assert ck.get("v", 3) == 3
    ck.set("v", 9)
    assert ck.get("v", 3) == 9 and ck.get("v") == 9


def test_constants_match_reference_values():
    from directdemod_b200 import constants as c
    from oracle import ddoracle as O
    assert c.PROC_CHUNKSIZE == O.PROC_CHUNKSIZE == 20000000
    assert c.NOAA_SYNCA == O.NOAA_SYNCA and c.NOAA_SYNCB == O.NOAA_SYNCB
    assert c.NOAA_T == O.NOAA_T and c.NOAA_MINPEAKDIST == O.NOAA_MINPEAKDIST
    assert c.NOAA_PEAKHEIGHTWIGGLE == O.NOAA_PEAKHEIGHTWIGGLE
    assert (c.FLT_LP, c.FLT_HP, c.FLT_BP, c.FLT_BS) == (0, 1, 2, 3)
    assert c.CHUNK_FREQOFFSET == "freqoffset" and c.CHUNK_BWLIM == "bwlim"
    assert c.IQ_FREQOFFSET == 30000 and c.NOAA_FMBW == 60000 and c.NOAA_CRUDESYNCSAMPRATE == 40960


def test_argument_validation_needs_no_gpu():
    """Errors are raised by the Python layer before anything touches the device."""
    from directdemod_b200 import comm, constants, filters
    with pytest.raises(ValueError):
        comm.commSignal(-5, np.zeros(3))
    with pytest.raises(TypeError):
        comm.commSignal(10, np.zeros((2, 3)))
    s = comm.commSignal(1000, np.zeros(8, dtype=np.complex64))
    with pytest.raises(ValueError):
        s.bwLim(2000)
    with pytest.raises(TypeError):
        comm.commSignal(1000, np.zeros(8)).offsetFreq(10.0)        # real signal, complex rotator
    with pytest.raises(ValueError):
        filters.butter(48000, 1000, typeFlt=constants.FLT_BS)
    with pytest.raises(ValueError):
        filters.remez(48000, [[0, 1000], [2000, 30000]], [1, 0])
    f = filters.blackmanHarris(151)
    assert f.isFIR and len(f.getB) == 151 and list(f.getA) == [1]
    assert not filters.butter(48000, 1000).isFIR


def test_constructor_copy_semantics_without_a_device():
    """Large float32 / complex64 inputs take the straight-to-device snapshot only when a device is
    present; without one the constructor keeps the reference's host copy (comm.py:38) and .signal
    works, while any operator still refuses to run on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from directdemod_b200 import comm, filters
    x = np.arange(1 << 17, dtype=np.float32)
    s = comm.commSignal(48000, x)
    x[:] = -1
    assert s.length == 1 << 17 and s.signal.dtype == np.float32 and s.signal[5] == 5.0
    with pytest.raises(RuntimeError):
        s.filter(filters.rollingAverage(3)).signal


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from directdemod_b200 import filters
    from directdemod_b200.fused import FusedChain
    with pytest.raises(RuntimeError):
        FusedChain(np.ones(4), 2, 0.0, 1000.0)
    with pytest.raises(RuntimeError):
        filters.rollingAverage(3).applyOn(np.arange(10.0))


def _write_wav_u8(path, pairs, fs):
    """Two-channel unsigned 8-bit PCM WAV with the canonical 44-byte header (source.py:66)."""
    import struct
    data = np.ascontiguousarray(pairs, dtype=np.uint8).tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 2, fs, fs * 2, 2, 8) \
        + b"data" + struct.pack("<I", len(data))
    assert len(hdr) == 44
    with open(path, "wb") as fh:
        fh.write(hdr + data)


def test_sources_read_like_the_reference(tmp_path):
    from directdemod_b200 import constants, source
    rng = np.random.default_rng(12)
    pairs = rng.integers(0, 256, (5000, 2), dtype=np.uint8)
    want = (pairs[:, 0] + 1j * pairs[:, 1]).astype("complex64") - (127.5 + 1j * 127.5)
    wav = tmp_path / "iq.wav"
    dat = tmp_path / "iq.dat"
    _write_wav_u8(wav, pairs, 2048000)
    pairs.tofile(dat)
    for src in (source.IQwav(str(wav)), source.IQwavAlt(str(wav)), source.IQdat(str(dat), 2048000)):
        assert src.length == 5000 and src.sampFreq == 2048000
        got = src.read(10, 4000)
        assert got.dtype == np.complex64 and np.array_equal(got, want[10:4000])
        assert np.array_equal(src.read(7), want[7:8])
        raw = src.readRaw(10, 4000)
        assert isinstance(raw, source.RawIQ8) and len(raw) == 3990 and np.array_equal(raw.to_complex64(), want[10:4000])
        for bad in ((-1, 5), (0, 5001), (5000, 5000)):
            with pytest.raises(ValueError):
                src.read(*bad)
        src.limitData(100, 600)
        assert src.length == 500 and np.array_equal(src.read(0, 500), want[100:600])
        src.limitData()
        assert src.length == 5000
    assert source.IQwav(str(wav)).sourceType == constants.SOURCE_IQWAV
    assert source.IQdat(str(dat)).sourceType == constants.SOURCE_IQDAT
    assert source.IQdat(str(dat)).sampFreq == constants.IQ_SDRSAMPRATE
    assert source.IQwav(str(wav), 1000000).sampFreq == 1000000


def test_reference_arm_runs_the_staged_unmodified_reference(tmp_path):
    """bench.py's CPU arm must drive the reference itself (oracle/stage_ref.py -> oracle/_ref, the copy
    that travels to the GPU box), chunk loop as decode_noaa.py:613-627, and both arms must print the
    same `config` dict."""
    import filecmp
    import bench
    from oracle import ref_shim, stage_ref
    if not ref_shim.available():
        pytest.skip("no reference checkout and nothing staged")
    if os.path.isdir(os.path.join(stage_ref.SRC_ROOT, "directdemod")):
        staged = stage_ref.stage()
        assert staged, "nothing staged"
        for path in staged:                                   # byte-for-byte copies, no edits
            assert filecmp.cmp(path, os.path.join(stage_ref.SRC_ROOT, "directdemod", os.path.basename(path)),
                               shallow=False)
    assert bench._cpu_kind() == "reference"
    assert bench.workload_config(4) == bench.workload_config(4) and "mode" not in bench.workload_config(1)
    # one small pass through the reference arm's worker against the oracle port of the same chain
    from oracle import ddoracle as O
    n = 300000
    bench._CPU.clear()
    bench._CPU.update(x=bench._cpu_chunk_input(5, n), n=n, kind="reference")
    ref_shim.load()
    from directdemod import chunker, comm, demod_fm, filters
    src = bench._LoopSource(bench._CPU["x"], 2)
    ck = chunker.chunker(src, n)
    bh, fm = filters.blackmanHarris(bench.NTAPS), demod_fm.demod_fm()
    got = comm.commSignal(bench.CRUDE_RATE)
    for i in ck.getChunks:
        got.extend(comm.commSignal(src.sampFreq, src.read(*i), ck).offsetFreq(bench.F_OFF).filter(bh)
                   .bwLim(bench.BW, uniq="First").funcApply(fm.demod).bwLim(bench.CRUDE_RATE, False))
    st = O.ChainState(bench.taps_bh151())
    want = np.concatenate([O.chain_chunk(bench._CPU["x"], bench.FS, bench.F_OFF, bench.taps_bh151(), bench.BW, st)[0]
                           for _ in range(2)])
    assert got.signal.shape == want.shape and np.array_equal(got.signal, want)
    assert bench._cpu_step(1) > 0
    bench._CPU.clear()


def test_cascade_equivalent_filter_reproduces_the_stage_by_stage_result():
    """filters.cascade (host part): the equivalent FIR taps and initial delay line of the C4 cascade
    (remez-1023 -> butter-8, BASELINE configs[3]) reproduce scipy's stage-by-stage stateful result
    (each stage from its own lfilter_zi, filters.py:45,69) far below the 1e-5 tolerance."""
    import scipy.signal as sps
    from directdemod_b200 import constants, filters
    fs = 2400000
    fir = filters.remez(fs, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    iir = filters.butter(fs, 100000, n=8)
    cas = filters.cascade([fir, iir])
    assert cas.isFIR and 1023 < len(cas.getB) < 1400 and cas.lookback() == len(cas.getB) - 1
    rng = np.random.default_rng(3)
    x = ((rng.standard_normal(60000) + 1j * rng.standard_normal(60000)) * 40).astype(np.complex64).astype(np.complex128)
    want = x
    for f in (fir, iir):
        want, _ = sps.lfilter(f.getB, f.getA, want, zi=sps.lfilter_zi(f.getB, f.getA))
    got, _ = sps.lfilter(cas.getB, [1.0], x, zi=cas._zir)
    err = np.sqrt(np.mean(np.abs(got - want) ** 2) / np.mean(np.abs(want) ** 2))
    assert err <= 1e-8, err
    # not a cascade: zero-phase stages, filters that never die out
    with pytest.raises(ValueError):
        filters.cascade([fir, filters.butter(fs, 100000, n=8, zeroPhase=True)])
    with pytest.raises(ValueError):
        filters.cascade([filters.filter([1.0], [1.0, -1.0])])               # integrator: pole on the unit circle
    with pytest.raises(ValueError):
        filters.cascade([filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP)], max_taps=64)


def test_chain_position_mirror_follows_the_reference_carry_rules():
    """fused.FusedChain keeps (sample counter, decimation offset, has-previous) on the host so the
    decoders' 93-chunk loops size their outputs without asking the library.  The mirror is pure
    integer arithmetic: pin it to the oracle's chunk loop (comm.py:76, comm.py:124, demod_fm.py:44-48)
    on ragged chunk sizes, including chunks shorter than the carried offset (an empty chunk, which
    scipy's lfilter refuses, must leave the mirror where it was)."""
    from directdemod_b200 import fused
    from oracle import ddoracle as od

    rng = np.random.default_rng(11)
    for decim, demod in [(1, True), (2, False), (7, True), (34, True), (50, False), (341, True)]:
        ch = object.__new__(fused.FusedChain)        # no device: only the host arithmetic is exercised
        ch.decim, ch.demod, ch._pos = decim, demod, (0, 0, False)
        taps = np.array([0.5, 0.5])
        st = od.ChainState(taps)
        fs = 1000 * decim
        sizes = [1, 0, max(decim - 1, 1), decim, decim + 1, 3, 0] + list(rng.integers(1, 5 * decim + 3, 40))
        for n in sizes:
            n = int(n)
            if n == 0:
                before = ch.position_cached
                assert ch.out_count(0) == 0
                ch._advance(0)
                assert ch.position_cached == before
                continue
            x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
            kept, _ = od.chain_chunk(x, fs, 0.0, taps, fs // decim, st, demod=False)
            if demod and len(kept):          # demod_fm.py:44-51 (it raises on an empty chunk; here: 0 outputs)
                want, st.fm_last = od.fm_discriminator(kept, st.fm_last)
            else:
                want = kept
            assert ch.out_count(n) == len(want), (decim, demod, n)
            assert ch.count_for(n, ch._pos[1], ch._pos[2]) == len(want)
            ch._advance(n)
            assert ch.position_cached[0] == st.n0
            assert ch.position_cached[1] == st.dec_off
            assert ch.position_cached[2] == (st.fm_last is not None) or not demod


def test_fill_sync_matches_the_unmodified_reference_on_ragged_sync_lists():
    """decode_noaa.__fillSync (decode_noaa.py:467-508) is pure host logic inside getImage: keep the syncs
    spaced by the modal distance, extrapolate back to the start, fill every gap up to the end.  The drop-in
    restates it with a set beside the list (the reference's `x not in list` is quadratic); pin it to the
    reference's own method on passes with missed syncs, false detections, late starts and ties in the modal
    spacing -- same list, element for element."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference checkout and nothing staged")
    ref = ref_shim.load()
    import importlib
    ref_noaa = importlib.import_module("directdemod.decode_noaa")
    from directdemod_b200 import decode_noaa as ours
    their = object.__new__(ref_noaa.decode_noaa)
    fill_ref = getattr(their, "_decode_noaa__fillSync")
    rng = np.random.default_rng(20)
    spacing = 20480                                   # half a second at 40960 Hz
    for case in range(40):
        n_lines = int(rng.integers(6, 120))
        start = int(rng.integers(0, 3 * spacing))
        syncs = start + spacing * np.arange(n_lines) + rng.integers(-3, 4, n_lines)
        keep = rng.random(n_lines) > (0.0 if case % 4 == 0 else 0.25)          # missed syncs
        keep[:2] = True                                # (an all-gaps pass has no modal spacing to speak of)
        syncs = syncs[keep]
        if case % 3 == 0:                              # false detections between two lines
            extra = rng.choice(syncs[:-1], size=min(3, len(syncs) - 1), replace=False) + rng.integers(500, 9000, min(3, len(syncs) - 1))
            syncs = np.sort(np.concatenate([syncs, extra]))
        syncs = np.unique(syncs).astype(np.float64)    # getImage hands over floats (sync / crudeRate * fs)
        max_len = int(syncs[-1] + rng.integers(0, 5 * spacing))
        want = fill_ref(list(syncs), max_len)
        got = ours.decode_noaa._fillSync(list(syncs), max_len)
        assert len(got) == len(want), case
        assert all(a == b for a, b in zip(got, want)), case


def test_image_rows_quantised_in_one_pass_equal_the_line_by_line_form():
    """getImage records every line's pixel row with the two numbers its formula uses and quantises all rows
    at once (decode_noaa._quantise_rows); the reference rounds, clips and casts line by line with the
    levels in force at that line (decode_noaa.py:432-452).  Same bytes, including levels that change from
    line to line, out-of-range pixels on both sides and the pass that never saw a telemetry frame."""
    from directdemod_b200 import decode_noaa as ours

    def line_by_line(v):
        v = np.round(v)
        v[v < 0] = 0
        v[v > 255] = 255
        return v.astype(np.uint8)
    rng = np.random.default_rng(7)
    rows = [rng.normal(0.35, 0.3, 2080) for _ in range(60)]
    slope = [float(rng.normal(420, 60)) for _ in rows]
    icpt = [float(rng.normal(-25, 30)) for _ in rows]
    low = [float(rng.normal(0.1, 0.02)) for _ in rows]
    high = [float(rng.normal(0.8, 0.02)) for _ in rows]
    got = ours._quantise_rows(list(zip(rows, slope, icpt)), list(zip(rows[:5], low, high)))
    want = np.array([line_by_line(r * a + b) for r, a, b in zip(rows, slope, icpt)])
    assert got.dtype == np.uint8 and got.shape == want.shape and np.array_equal(got, want)
    assert got.min() == 0 and got.max() == 255                      # both clips were exercised
    got = ours._quantise_rows([], list(zip(rows, low, high)))       # no telemetry frame: first-guess levels
    want = np.array([line_by_line(255 * (r - lo) / (hi - lo)) for r, lo, hi in zip(rows, low, high)])
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    with pytest.raises(ValueError):                                 # the reference's max() of an empty sequence
        ours._quantise_rows([], [])


def test_usefulness_measure_equals_the_reference_expression():
    """getCrudeSync's "was a NOAA signal found" measure (decode_noaa.py:794-799): the smallest, over all
    runs of NOAA_DETECTCONSSYNCSNUM consecutive sync spacings, of the run's largest deviation from half a
    second.  The drop-in takes one windowed maximum instead of a list comprehension of np.max calls: same
    number, and the same error for a pass with too few syncs."""
    from directdemod_b200 import constants, decode_noaa as ours
    n = constants.NOAA_DETECTCONSSYNCSNUM
    rng = np.random.default_rng(3)
    rate = 60235
    for count in (n + 1, n + 2, 40, 300):
        sync = np.cumsum(rng.integers(rate // 2 - 40, rate // 2 + 40, count))
        sync[rng.integers(0, count)] += 5000                      # a false detection in the middle
        dev = np.abs(np.diff(sync) - (rate * 0.5))
        want = np.min([np.max(dev[i:i + n]) for i in range(len(dev) - n + 1)])
        assert ours._steadiest(sync, rate, n) == want
    for count in (2, n):                                          # too few syncs: np.min of an empty list
        with pytest.raises(ValueError):
            ours._steadiest(np.arange(count) * (rate // 2), rate, n)


def test_iir_analysis_runs_on_the_host_and_means_what_the_header_says():
    """ddm_iir_analyse (include/ddemod.h) is host arithmetic: the warm-up after which a segment started from
    zero state agrees with the running filter, and the roundoff floor of the float64 transposed direct form
    II recursion (scipy's lfilter, filters.py:69) against extended precision on white noise -- the number
    DDM_IIR_AUTO compares with its tolerance.  Checked against an independent measurement (numpy longdouble
    recursion in this test vs scipy.signal.lfilter) for the three recursive filters the decoders build, and
    for the two edge cases (FIR: nothing to analyse; unstable: no finite warm-up)."""
    import scipy.signal as sps
    from directdemod_b200 import _lib
    lib = _lib.lib()

    def analyse(b, a):
        b = np.ascontiguousarray(b, dtype=np.float64)
        a = np.ascontiguousarray(a, dtype=np.float64)
        w, nf = ctypes.c_int64(), ctypes.c_double()
        _lib.check(lib.ddm_iir_analyse(b.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(b),
                                       a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(a),
                                       ctypes.byref(w), ctypes.byref(nf)), "ddm_iir_analyse")
        return int(w.value), float(nf.value)

    def measured_floor(b, a, n=30000):
        rng = np.random.default_rng(0)
        x = rng.standard_normal(n)
        y64 = sps.lfilter(b, a, x)
        bl, al = np.asarray(b, dtype=np.longdouble), np.asarray(a, dtype=np.longdouble)
        bl, al = bl / al[0], al / al[0]
        k = max(len(bl), len(al))
        bl, al = np.pad(bl, (0, k - len(bl))), np.pad(al, (0, k - len(al)))
        z = np.zeros(k, dtype=np.longdouble)                       # transposed direct form II, extended precision
        y = np.empty(n, dtype=np.longdouble)
        for i, xi in enumerate(x.astype(np.longdouble)):
            yi = bl[0] * xi + z[0]
            z[:-1] = bl[1:] * xi + z[1:] - al[1:] * yi
            y[i] = yi
        err = (y64.astype(np.longdouble) - y)[n // 2:]
        return float(np.sqrt(np.mean(err ** 2)) / np.sqrt(np.mean(y[n // 2:] ** 2)))

    cases = {
        "C4 low-pass, butter(8) at 100 kHz of 2.4 Msps (parallel under AUTO)": sps.butter(8, 100e3 / 1.2e6),
        "NOAA band-pass, butter(6) 400-4400 Hz of 60235 (decode_noaa.py:274)": sps.butter(6, [400 / 30117.5, 4400 / 30117.5], btype="band"),
        "AFSK band-pass, butter(6) 700-2700 Hz of 48000 (decode_afsk1200.py:98)": sps.butter(6, [700 / 24000, 2700 / 24000], btype="band"),
    }
    floors = {}
    for name, (b, a) in cases.items():
        warm, nf = analyse(b, a)
        floors[name] = nf
        r = np.max(np.abs(np.roots(a)))
        assert 0 < r < 1 and warm > 0
        assert r ** warm < 1e-6, (name, warm, r)                  # the slowest mode has died out after the warm-up
        assert warm < 40 * np.log(1e-16) / np.log(r), (name, warm, r)     # ... and it is not absurdly long
        want = measured_floor(b, a)
        assert want / 4 <= nf <= want * 4, (name, nf, want)       # measured here: within 25 % for all three
    lp, noaa, afsk = floors.values()
    assert lp < 1e-7 < afsk < noaa                                # which side of the AUTO switch each one falls
    assert analyse(sps.windows.hamming(31), [1.0]) == (0, 0.0)    # FIR: no recursion
    assert analyse([1.0], [1.0, -1.01])[0] == -1                  # unstable: never segment-parallel


def test_c_abi_reports_errors_as_status_plus_message_without_a_device():
    """The boundary's error convention (include/ddemod.h): every entry point returns an int
    status, 0 or negative, with the reason in ddm_last_error(); nothing is thrown across the ABI and
    argument checks come before any device work -- so they can be exercised here, with no GPU."""
    from directdemod_b200 import _lib
    lib = _lib.lib()
    pd = ctypes.POINTER(ctypes.c_double)
    ones, zi = np.ones(3), np.zeros(4)
    handle, count = ctypes.c_void_p(), ctypes.c_int64()
    assert lib.ddm_version() >= 100
    checks = [
        (lambda: lib.ddm_lfilter_zi(None, 3, ones.ctypes.data_as(pd), 3, zi.ctypes.data_as(pd)), b"ddm_lfilter_zi"),
        (lambda: lib.ddm_lfilter_zi(ones.ctypes.data_as(pd), 3, np.array([0.0, 1.0, 1.0]).ctypes.data_as(pd), 3,
                                    zi.ctypes.data_as(pd)), b"a[0]"),
        (lambda: lib.ddm_chain_create(0, ones.ctypes.data_as(pd), 3, 0, 0.0, 1000.0, 0, 0, ctypes.byref(handle)),
         b"decimation"),
        (lambda: lib.ddm_chain_create(0, ones.ctypes.data_as(pd), 3, 2, 0.0, -1.0, 0, 0, ctypes.byref(handle)),
         b"sampling rate"),
        (lambda: lib.ddm_chain_create(0, ones.ctypes.data_as(pd), 3, 2, 0.0, 1000.0, 0, 0, None), b"NULL"),
        (lambda: lib.ddm_chain_create(0, ones.ctypes.data_as(pd), 3, 2, 0.0, 1000.0, 7, 0, ctypes.byref(handle)),
         b"out_mode"),
        (lambda: lib.ddm_chain_apply_dev(None, None, 0, None, 0, ctypes.byref(count), None), b"NULL handle"),
    ]
    for call, needle in checks:
        rc = call()
        assert rc == _lib.ERR_INVALID, (rc, needle)
        assert needle in lib.ddm_last_error(), (needle, lib.ddm_last_error())
        assert not handle.value                              # no half-built handle is handed out
    with pytest.raises(_lib.DdmError) as info:               # the Python layer turns the status into an exception
        _lib.check(checks[2][0](), "ddm_chain_create")
    assert "decimation" in str(info.value)
