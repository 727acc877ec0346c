"""GPU parity of the sync path: normalised correlation, threshold selection, peak picking,
decode_noaa.getCrudeSync / getAccurateSync (positions bit-exact), and the AFSK correlator bank."""

import os

import numpy as np
import pytest

from oracle import ddoracle as O
from tests.util import TOL, ArraySource, apt_iq, oracle_chain

pytestmark = pytest.mark.gpu


def test_correlate_matches_oracle_runs_and_direct_paths():
    from directdemod_b200 import sync
    rng = np.random.default_rng(1)
    n = 200000
    hay = np.abs(rng.standard_normal(n)) + 0.1
    for needle in (sync.sync_needle(O.NOAA_SYNCA, 60235), sync.sync_needle(O.NOAA_SYNCB, 60235, False),
                   rng.standard_normal(561), rng.standard_normal(64), np.ones(7)):
        for normalised in (True, False):
            got = sync.correlate(hay, needle, normalised).cpu().numpy()
            want = O.ncc(hay, needle) if normalised else __import__("scipy.signal").signal.correlate(hay, needle, "same")
            assert got.shape == want.shape
            assert O.rel_rms(got, want) <= 1e-9, (len(needle), normalised, O.rel_rms(got, want))
    # float32 device input (what the AM kernel hands over)
    import torch
    h32 = torch.from_numpy(hay.astype(np.float32)).cuda()
    needle = sync.sync_needle(O.NOAA_SYNCA, 60235)
    got = sync.correlate(h32, needle).cpu().numpy()
    assert O.rel_rms(got, O.ncc(hay.astype(np.float32).astype(np.float64), needle)) <= 1e-9
    assert np.array_equal(sync.sync_needle(O.NOAA_SYNCA, 60235), O.sync_needle(O.NOAA_SYNCA, 60235))


def test_topk_sums_compaction_and_group_scan():
    import ctypes as C
    import torch
    from directdemod_b200 import _dev, _lib
    rng = np.random.default_rng(2)
    x = rng.standard_normal(1000003)
    x[1234] = x[99999] = 7.5            # ties at the top
    x[5] = -9.0
    xd = torch.from_numpy(x).cuda()
    for k in (1, 2, 17, 4000, len(x)):
        top, bot = C.c_double(), C.c_double()
        _lib.check(_lib.lib().ddm_topk_sums(0, _dev.ptr(xd), len(x), k, C.byref(top), C.byref(bot), _dev.stream_ptr(0)), "topk")
        s = np.sort(x)
        assert abs(top.value - s[-k:].sum()) <= 1e-9 * max(1.0, abs(s[-k:].sum())), k
        assert abs(bot.value - s[:k].sum()) <= 1e-9 * max(1.0, abs(s[:k].sum())), k
    thr = 2.5
    want = np.argwhere(x > thr).ravel()
    idx = torch.empty(len(want) + 10, dtype=torch.int64, device="cuda")
    val = torch.empty(len(want) + 10, dtype=torch.float64, device="cuda")
    cnt = C.c_int64()
    _lib.check(_lib.lib().ddm_compact_above(0, _dev.ptr(xd), len(x), thr, _dev.ptr(idx), _dev.ptr(val), idx.numel(),
                                            C.byref(cnt), _dev.stream_ptr(0)), "compact")
    assert cnt.value == len(want)
    assert np.array_equal(idx[:cnt.value].cpu().numpy(), want)
    assert np.array_equal(val[:cnt.value].cpu().numpy(), x[want])


def test_pick_peaks_matches_reference_logic_including_ties():
    import torch
    from directdemod_b200 import sync
    rng = np.random.default_rng(3)
    fs = 1000
    n = 20000
    cor = 0.05 * rng.standard_normal(n)
    for p in range(250, n, 500):
        cor[p - 3:p + 4] += np.array([0.2, 0.5, 0.8, 1.0, 0.8, 0.5, 0.2])
    cor[2750] = cor[2751] = 1.5         # exact tie: the first one must win (strict '<')
    got, _ = sync.pick_peaks(torch.from_numpy(cor).cuda(), fs, 40)
    want = O.pick_sync_peaks(cor, fs, 40)
    assert np.array_equal(got, want)


@pytest.fixture(scope="module")
def apt_pass():
    fs = 2048000
    x = apt_iq(7, 11.0, fs=fs)          # 22.5 M samples -> two PROC_CHUNKSIZE chunks
    return x, fs


def test_crude_sync_positions_bit_exact(apt_pass):
    from directdemod_b200 import decode_noaa
    x, fs = apt_pass
    dec = decode_noaa.decode_noaa(ArraySource(x, fs), 30000.0)
    syncA, syncB = dec.getCrudeSync()
    assert dec.useful == 1
    # oracle: the reference's algorithm on the same input (float64 scipy path)
    taps = O.taps_blackman_harris(151)[0]
    audio, rate = O.chain_stream(x, fs, 30000.0, taps, 60000)
    assert rate == 60235
    am = O.am_envelope_chunked(audio)
    wantA, _ = O.find_syncs(am, rate, O.NOAA_SYNCA)
    wantB, _ = O.find_syncs(am, rate, O.NOAA_SYNCB)
    assert np.array_equal(np.asarray(syncA), wantA)
    assert np.array_equal(np.asarray(syncB), wantB)
    assert len(wantA) >= 20 and np.all(np.abs(np.diff(wantA)[:-1] - rate / 2) <= 2)   # last one: cut-off line
    # intermediate signals within the stated tolerance
    from tests.util import wrap_rel_rms
    assert wrap_rel_rms(dec._audOut.signal, audio) <= TOL


def test_accurate_sync_positions_bit_exact(apt_pass):
    from directdemod_b200 import decode_noaa
    x, fs = apt_pass
    n_use = int(6.2 * fs)               # >= 11 syncs, or the usefulness test has nothing to look at
    xs = x[:n_use]
    dec = decode_noaa.decode_noaa(ArraySource(xs, fs), 30000.0)
    res = dec.getAccurateSync()
    asyncA, asyncB = res[0], res[4]
    # oracle windows (decode_noaa.py:826-856 restated with the oracle primitives)
    taps = O.taps_blackman_harris(151)[0]
    audio, rate = O.chain_stream(xs, fs, 30000.0, taps, 60000)
    am = O.am_envelope_chunked(audio)
    width = int(3 * O.NOAA_T * 40 * fs)
    ham = O.taps_hamming(492)[0]
    for bits, got in ((O.NOAA_SYNCA, asyncA), (O.NOAA_SYNCB, asyncB)):
        crude, _ = O.find_syncs(am, rate, bits)
        want = []
        for c in crude / rate * fs:
            a, b = int(c) - width, int(c) + width
            if a < 0 or b > len(xs):
                continue
            if len(want) == 4:          # a handful of windows keeps the CPU oracle in seconds
                break
            w, _ = O.mix(xs[a:b], 30000.0, fs, 0)
            w = O.filt_zero_phase(taps, [1], w)
            w, _ = O.fm_discriminator(w, None, store_state=True)
            w = O.am_envelope(w)
            pk, _ = O.find_syncs(w, fs, bits, prefilter_taps=ham)
            want.append(pk[0] + a)
        assert len(want) >= 3
        assert list(map(int, got[:len(want)])) == list(map(int, want))


def test_accurate_sync_32_windows_per_sync_word_bit_exact(tmp_path):
    """getAccurateSync on a 17.5 s pass: the first 32 windows of EACH sync word against the oracle
    (decode_noaa.py:826-856 restated), the CPU windows computed by a pool of processes
    (tests/oracle_pool.py, its own interpreter)."""
    import subprocess
    import sys
    from directdemod_b200 import decode_noaa
    fs = 2048000
    x = apt_iq(17, 17.5, fs=fs)
    dec = decode_noaa.decode_noaa(ArraySource(x, fs), 30000.0)
    res = dec.getAccurateSync()
    asyncA, asyncB = res[0], res[4]
    taps = O.taps_blackman_harris(151)[0]
    audio, rate = oracle_chain(x, fs, 30000.0, taps, 60000, list(range(0, len(x), 20000000)) + [len(x)])
    am = O.am_envelope_chunked(audio)
    width = int(3 * O.NOAA_T * 40 * fs)
    windows, starts, which = [], [], []
    for w, bits in enumerate((O.NOAA_SYNCA, O.NOAA_SYNCB)):
        crude, _ = O.find_syncs(am, rate, bits)
        cnt = 0
        for c in crude / rate * fs:
            a, b = int(c) - width, int(c) + width
            if a < 0 or b > len(x):
                continue
            if cnt == 32:
                break
            windows.append(x[a:b])
            starts.append(a)
            which.append(w)
            cnt += 1
        assert cnt == 32, (w, cnt)
    jobs, out = str(tmp_path / "jobs.npz"), str(tmp_path / "out.npy")
    np.savez(jobs, windows=np.stack(windows), starts=np.asarray(starts), which=np.asarray(which), fs=fs,
             bits=np.asarray([O.NOAA_SYNCA, O.NOAA_SYNCB]))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # one BLAS thread per worker: numpy's direct convolution calls a threaded ddot per output sample,
    # and a pool of processes each spinning up a thread team per call does not finish in hours
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
    subprocess.run([sys.executable, "-m", "tests.oracle_pool", jobs, out, str(min(16, os.cpu_count() or 1))],
                   cwd=root, check=True, timeout=900, env=env)
    want = np.load(out)
    assert list(map(int, asyncA[:32])) == [int(v) for v in want[:32]]
    assert list(map(int, asyncB[:32])) == [int(v) for v in want[32:]]


@pytest.mark.slow
def test_crude_sync_of_a_120_second_pass_bit_exact():
    """A long pass (120 s = 245.76 M samples, 13 PROC_CHUNKSIZE chunks, 240 lines): crude sync
    positions of both sync words against the oracle run chunk by chunk like the reference."""
    from directdemod_b200 import decode_noaa
    from tests.util import apt_iq_long
    fs = 2048000
    x = apt_iq_long(23, 120.0, fs=fs)
    dec = decode_noaa.decode_noaa(ArraySource(x, fs), 30000.0)
    syncA, syncB = dec.getCrudeSync()
    assert dec.useful == 1
    taps = O.taps_blackman_harris(151)[0]
    audio, rate = oracle_chain(x, fs, 30000.0, taps, 60000, list(range(0, len(x), 20000000)) + [len(x)])
    assert rate == 60235
    from tests.util import wrap_rel_rms
    assert wrap_rel_rms(dec._audOut.signal, audio) <= TOL
    am = O.am_envelope_chunked(audio)
    wantA, _ = O.find_syncs(am, rate, O.NOAA_SYNCA)
    wantB, _ = O.find_syncs(am, rate, O.NOAA_SYNCB)
    assert len(wantA) >= 238 and len(wantB) >= 238
    assert np.array_equal(np.asarray(syncA), wantA)
    assert np.array_equal(np.asarray(syncB), wantB)


def test_batched_accurate_sync_equals_per_window_path(apt_pass):
    """The row-batched accurate sync (one launch sequence per 256 windows) against the per-window
    operator-by-operator path: identical positions, peak heights and time-sync means."""
    from directdemod_b200 import constants, decode_noaa
    x, fs = apt_pass
    dec = decode_noaa.decode_noaa(ArraySource(x[:int(7.2 * fs)], fs), 30000.0)
    dec.getCrudeSync()
    width = int(3 * constants.NOAA_T * 40 * fs)
    csync = dec._syncA / dec._syncCrudeSampRate * fs
    for norm in (True, False):
        a = dec._accurate(csync, constants.NOAA_SYNCA, width, norm, batch=5)       # several batches
        b = dec._accurate_loop(csync, constants.NOAA_SYNCA, width, norm)
        assert len(a[0]) == len(b[0]) >= 8
        assert list(map(int, a[0])) == list(map(int, b[0]))
        np.testing.assert_allclose(a[1], b[1], rtol=1e-5)
        assert [v is None for v in a[2]] == [v is None for v in b[2]]
        np.testing.assert_allclose([v for v in a[2] if v is not None], [v for v in b[2] if v is not None], rtol=1e-5)


def test_afsk_bank_matches_oracle():
    from directdemod_b200 import afsk
    rng = np.random.default_rng(5)
    for bw, n in ((48000, 30000), (22050, 5000), (48000, 30)):
        t = np.arange(n) / bw
        x = np.sin(2 * np.pi * np.where((t * 1200).astype(int) % 2 == 0, 1200, 2200) * t) + 0.1 * rng.standard_normal(n)
        got = afsk.mark_space_bank(x, bw)
        want = O.afsk_bank(x, bw)
        assert got.shape == want.shape
        nbuf = int(np.round(bw / 1200))
        assert np.all(got[max(0, n - nbuf):] == 0)
        assert O.rel_rms(got, want) <= TOL


def test_noaa_pass_matches_unmodified_reference(golden):
    """The whole NOAA APT decode of a synthetic 14 s pass against what the UNMODIFIED reference
    produced for the same input (oracle/gen_golden_noaa.py): sync positions bit-exact, image
    pixels within +-1 LSB on >= 99.9 % of pixels (BASELINE.json north_star)."""
    from directdemod_b200 import decode_noaa
    g = golden("noaa_pass")
    x = apt_iq(int(g["seed"]), float(g["seconds"]), fs=int(g["fs"]), f_off=float(g["f_off"]))
    assert abs(float(np.abs(x[::1000]).sum()) - float(g["input_checksum"][0])) < 1e-3, "generator drifted"
    dec = decode_noaa.decode_noaa(ArraySource(x, int(g["fs"])), float(g["f_off"]))
    syncA, syncB = dec.getCrudeSync()
    assert dec.useful == int(g["useful"]) == 1
    assert np.array_equal(np.asarray(syncA), g["syncA"])
    assert np.array_equal(np.asarray(syncB), g["syncB"])
    img = dec.getImage
    want = g["image"]
    assert img.dtype == np.uint8 and img.shape == want.shape
    diff = np.abs(img.astype(np.int32) - want.astype(np.int32))
    frac = float(np.mean(diff <= 1))
    assert frac >= 0.999, (frac, int(diff.max()))


@pytest.mark.parametrize("exact_iir", [False, True], ids=["default-segment-parallel", "exact-replay"])
def test_afsk_front_end_matches_oracle(exact_iir):
    """decode_afsk1200.getMsg up to the bit-edge correlation (decode_afsk1200.py:62-158) on a
    synthetic FM-modulated AFSK stream at 960 kHz IQ with bw = 48000 (at a literal
    48 kHz IQ rate the reference's own 151-tap filter wipes the packet out).

    Both execution modes of the 12th-order band-pass are tested -- the DEFAULT (segment-parallel DFMA,
    the one the benchmarks time) and the bit-exact replay.  The bound is the same for both and is the
    reference's own: scipy's float64 tf-form recursion for this filter has a measured roundoff floor
    of ~4e-5 relative (ddm_iir_analyse: float64 vs long double on white noise), i.e. two float64 runs
    of scipy itself whose inputs differ in the last bit stay that far apart; the filter's input here
    already differs from the float64 oracle's by the fp32 rounding of the FM stage.  What the decoder
    consumes are the SIGNS of the bank output and the edge positions, asserted exactly below."""
    import scipy.signal as sps
    from directdemod_b200 import afsk
    rng = np.random.default_rng(6)
    fs, bw, n = 960000, 48000, 1500000
    t = np.arange(n) / fs
    bits = rng.integers(0, 2, int(n / fs * 1200) + 2)
    tone = np.where(bits[(t * 1200).astype(int)] == 1, 1200.0, 2200.0)
    audio = np.sin(2 * np.pi * np.cumsum(tone) / fs)
    x = (50 * np.exp(1j * 2 * np.pi * 3000 * np.cumsum(audio) / fs)
         + 1.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    sig, bf, changes = afsk.front_end(ArraySource(x, fs), 0.0, bw, exact_iir=exact_iir)
    # oracle
    taps = O.taps_blackman_harris(151)[0]
    iq, rate = O.chain_stream(x, fs, 0.0, taps, bw, demod=False)
    assert rate == sig.sampRate == 48000
    fm, _ = O.fm_discriminator(iq, None)
    b, a = O.taps_butter(rate, 700, 2700, n=6, kind=O.FLT_BP)
    aud, _ = O.filt_stateful(b, a, fm, O.initial_zi(b, a))
    want_bf = O.afsk_bank(aud, bw)
    spb = bw // 1200
    kernel = np.ones(spb)
    kernel[:spb // 2] = -1
    want_ch = np.correlate(np.sign(want_bf), kernel, mode="same") / spb
    # the band-pass output agrees to the filter's own roundoff floor (its input differs from the
    # float64 oracle's by the fp32 rounding of the FM stage)
    floor = 10 * b_floor(b, a)
    assert O.rel_rms(sig.signal, aud) <= max(TOL, floor)
    got_bf = bf.cpu().numpy()
    assert O.rel_rms(got_bf, want_bf) <= max(10 * TOL, 2 * floor)
    got_ch = changes.cpu().numpy()
    assert got_ch.shape == want_ch.shape
    # bit decisions: the edge correlation is built from signs, so it is exact wherever no sample of
    # its window sat within rounding distance of zero
    assert np.mean(np.abs(got_ch - want_ch) < 1e-9) >= 0.999
    strong = np.abs(want_ch) > 0.5
    assert np.array_equal(np.sign(got_ch[strong]), np.sign(want_ch[strong]))
    # downstream decisions (decode_afsk1200.py:157-170): the sign of the bank output wherever it is not
    # within the filter's floor of zero, and the bit-edge positions = local extrema of `changes`
    clear = np.abs(want_bf) > 20 * floor * np.sqrt(np.mean(want_bf ** 2))
    assert np.mean(clear) > 0.95
    assert np.array_equal(np.sign(got_bf[clear]), np.sign(want_bf[clear]))

    def edges(c):
        mid = c[1:-1]
        return np.nonzero((np.abs(mid) > 0.9) & (np.abs(mid) >= np.abs(c[:-2])) & (np.abs(mid) > np.abs(c[2:])))[0]
    assert np.array_equal(edges(got_ch), edges(want_ch))


def b_floor(b, a):
    import ctypes as C
    from directdemod_b200 import _lib
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    w, nf = C.c_int64(), C.c_double()
    _lib.check(_lib.lib().ddm_iir_analyse(b.ctypes.data_as(C.POINTER(C.c_double)), len(b),
                                          a.ctypes.data_as(C.POINTER(C.c_double)), len(a), C.byref(w), C.byref(nf)), "analyse")
    return nf.value


def test_device_group_scan_equals_sequential_scan():
    """ddm_pick_peaks (dominance formulation on the device) against the sequential scan of
    decode_noaa.py:731-746 (ddm_compact_above + ddm_group_peaks, and the oracle's Python loop) on
    random data: dense candidates, plateaus and exact ties, various window lengths."""
    import ctypes as C
    import torch
    from directdemod_b200 import _dev, _lib, sync
    rng = np.random.default_rng(31)
    l = _lib.lib()
    for trial, (n, dist, thr_q) in enumerate([(50000, 449.5, 0.3), (50000, 450.0, 0.6), (200003, 2710.75, 0.1),
                                             (10000, 37.2, 0.5), (5000, 6000.0, 0.2), (300000, 1024.0, 0.45)]):
        x = rng.standard_normal(n)
        if trial % 2 == 0:
            x = np.round(x * 4) / 4            # many exact ties and plateaus
        x[rng.integers(0, n, 20)] = x.max()    # repeated global maxima
        thr = float(np.quantile(x, thr_q))
        xd = torch.from_numpy(x).cuda()
        got = np.empty(n, dtype=np.int64)
        cnt = C.c_int64()
        _lib.check(l.ddm_pick_peaks(0, _dev.ptr(xd), n, thr, dist, got.ctypes.data_as(C.POINTER(C.c_int64)), n,
                                    C.byref(cnt), _dev.stream_ptr(0)), "ddm_pick_peaks")
        # the same through the table walk kept for inputs with more than 2^20 dominant candidates
        got2 = np.empty(n, dtype=np.int64)
        cnt2 = C.c_int64()
        os.environ["DDM_PEAKS_DENSE"] = "1"
        try:
            _lib.check(l.ddm_pick_peaks(0, _dev.ptr(xd), n, thr, dist, got2.ctypes.data_as(C.POINTER(C.c_int64)), n,
                                        C.byref(cnt2), _dev.stream_ptr(0)), "ddm_pick_peaks")
        finally:
            del os.environ["DDM_PEAKS_DENSE"]
        assert cnt2.value == cnt.value and np.array_equal(got[:cnt.value], got2[:cnt.value]), trial
        # sequential reference scan over the candidate list
        cand = np.argwhere(x > thr).ravel().astype(np.int64)
        want = np.empty(len(cand), dtype=np.int64)
        wc = C.c_int64()
        _lib.check(l.ddm_group_peaks(cand.ctypes.data_as(C.POINTER(C.c_int64)),
                                     np.ascontiguousarray(x[cand]).ctypes.data_as(C.POINTER(C.c_double)), len(cand),
                                     dist, want.ctypes.data_as(C.POINTER(C.c_int64)), len(want), C.byref(wc)), "group")
        assert cnt.value == wc.value, (trial, cnt.value, wc.value)
        assert np.array_equal(got[:cnt.value], want[:wc.value]), trial
    # and through pick_peaks against the oracle's restatement of the reference loop
    fs = 1000
    cor = rng.standard_normal(40000) * 0.2
    cor[::500] += 3.0
    got, _ = sync.pick_peaks(torch.from_numpy(cor).cuda(), fs, 40)
    assert np.array_equal(got, O.pick_sync_peaks(cor, fs, 40))


def test_fifo_medians_equal_the_appending_loop():
    """getImage reads two 10 000-sample FIFOs through their medians after every line
    (decode_noaa.py:358-366); the batched form must equal the append-truncate-median loop."""
    from directdemod_b200 import decode_noaa as dn
    rng = np.random.default_rng(17)
    fifo_len = 1000
    chunks = [rng.standard_normal(int(k)).astype(np.float32).astype(np.float64) for k in rng.integers(1, 140, 60)]
    chunks[7] = np.repeat(chunks[7][:1], 50)                       # ties
    got = dn._fifo_medians(chunks, fifo_len, "cuda")
    fifo = np.zeros(0)
    for j, c in enumerate(chunks):
        fifo = np.concatenate([fifo, c])[-fifo_len:]
        assert got[j] == np.median(fifo), j
    assert dn._fifo_medians([], fifo_len, "cuda").size == 0
