"""The oracle port (oracle/ddoracle.py) against the UNMODIFIED reference, imported through
oracle/ref_shim.py from the reference checkout or the staged copy (oracle/_ref), on drawn inputs.

The committed fixtures (tests/golden) pin fixed cases; here the two run side by side on shapes a
property-based generator draws -- chunk sizes that split blocks, decimation factors that do not divide the
chunk, offsets of either sign, filters of several kinds -- which is what "the oracle restates the
reference" has to mean for every input, not for nine of them.  CPU only; skipped when no reference is
reachable.  No GPU code is involved: this validates the checker the GPU tests rely on.  The last tests do
the same for the drop-in's own host-only pieces (the chunker, the filter designers and their errors), which
need no device either.
"""

import numpy as np
import pytest

from oracle import ddoracle as O
from oracle import ref_shim

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st      # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="no reference checkout and nothing staged")


class _Src:
    """What chunker.chunker wants from a source (source.py: length) -- nothing else is touched."""

    def __init__(self, n):
        self.length = n


def _ref():
    ref_shim.load()
    from directdemod import chunker, comm, constants, demod_am, demod_fm, filters
    return chunker, comm, constants, demod_am, demod_fm, filters


def _signal(seed, n, complex_=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(n) * 40
    if complex_:
        x = x + 1j * rng.standard_normal(n) * 40
        return x.astype(np.complex64)
    return x


@settings(max_examples=25, deadline=None, derandomize=True)
@given(n=st.integers(400, 6000), chunk=st.integers(97, 3000), decim=st.integers(1, 40),
       f_off=st.floats(-3e5, 3e5, allow_nan=False), ntaps=st.sampled_from([1, 7, 31, 151]),
       demod=st.booleans(), seed=st.integers(0, 2**16))
def test_fused_chain_port_equals_reference_chunk_loop(n, chunk, decim, f_off, ntaps, demod, seed):
    """decode_noaa.py:613-627 / decode_afsk1200.py:79-91: the chunk loop through commSignal with the
    chunker's carried variables, against the port's explicit state."""
    chunker, comm, constants, demod_am, demod_fm, filters = _ref()
    fs = 2048000
    target = fs / decim * 1.0001 if decim > 1 else fs          # int(fs / target) == decim
    if int(fs / target) != decim:
        target = fs / decim
    x = _signal(seed, n)
    ck = chunker.chunker(_Src(n), chunk)
    bh = filters.blackmanHarris(ntaps)
    fm = demod_fm.demod_fm()
    out = comm.commSignal(int(fs / int(fs / target)))
    for a, b in ck.getChunks:
        sig = comm.commSignal(fs, x[a:b], ck).offsetFreq(f_off).filter(bh).bwLim(target, uniq="First")
        if demod:
            if sig.length == 0:
                return                      # the reference's discriminator raises on an empty chunk
            sig = sig.funcApply(fm.demod)
        out.extend(sig)
    st_ = O.ChainState(O.taps_blackman_harris(ntaps)[0])
    parts = []
    for a, b in O.chunk_bounds(n, chunk):
        y, _ = O.chain_chunk(x[a:b], fs, f_off, O.taps_blackman_harris(ntaps)[0], target, st_, demod=demod)
        parts.append(y)
    want = np.concatenate(parts)
    got = np.asarray(out.signal)
    assert got.shape == want.shape
    assert np.array_equal(got, want)          # the same scipy calls in the same order: the same bits


@settings(max_examples=20, deadline=None, derandomize=True)
@given(n=st.integers(200, 3000), split=st.integers(1, 199), seed=st.integers(0, 2**16),
       kind=st.sampled_from(["bh", "hamming", "gauss", "roll", "butlp", "buthp", "butbp", "butbs", "remez"]),
       cplx=st.booleans())
def test_stateful_filters_port_equals_reference(n, split, seed, kind, cplx):
    """filters.py:21-75: every designer, stateful over two chunks, stateless and zero phase."""
    chunker, comm, constants, demod_am, demod_fm, filters = _ref()
    fs = 48000
    make = {
        "bh": (lambda **kw: filters.blackmanHarris(31, **kw), lambda: O.taps_blackman_harris(31)),
        "hamming": (lambda **kw: filters.hamming(41, **kw), lambda: O.taps_hamming(41)),
        "gauss": (lambda **kw: filters.gaussian(25, 3.0, **kw), lambda: O.taps_gaussian(25, 3.0)),
        "roll": (lambda **kw: filters.rollingAverage(5, **kw), lambda: O.taps_rolling_average(5)),
        "butlp": (lambda **kw: filters.butter(fs, 3000, n=5, **kw), lambda: O.taps_butter(fs, 3000, n=5)),
        "buthp": (lambda **kw: filters.butter(fs, 3000, n=3, typeFlt=constants.FLT_HP, **kw),
                  lambda: O.taps_butter(fs, 3000, n=3, kind=O.FLT_HP)),
        "butbp": (lambda **kw: filters.butter(fs, 1000, 4000, n=3, typeFlt=constants.FLT_BP, **kw),
                  lambda: O.taps_butter(fs, 1000, 4000, n=3, kind=O.FLT_BP)),
        "butbs": (lambda **kw: filters.butter(fs, 1000, 4000, n=2, typeFlt=constants.FLT_BS, **kw),
                  lambda: O.taps_butter(fs, 1000, 4000, n=2, kind=O.FLT_BS)),
        "remez": (lambda **kw: filters.remez(fs, [[0, 4000], [6000, 24000]], [1, 0], ntaps=33, **kw),
                  lambda: O.taps_remez(fs, [[0, 4000], [6000, 24000]], [1, 0], ntaps=33)),
    }[kind]
    x = _signal(seed, n, cplx)
    b, a = make[1]()
    f = make[0]()
    assert np.array_equal(np.asarray(f.getB), np.asarray(b)) and np.array_equal(np.asarray(f.getA), np.asarray(a))
    y1 = f.applyOn(x[:split])
    y2 = f.applyOn(x[split:])
    w1, z = O.filt_stateful(b, a, x[:split], O.initial_zi(b, a))
    w2, _ = O.filt_stateful(b, a, x[split:], z)
    assert np.array_equal(np.concatenate([y1, y2]), np.concatenate([w1, w2]))
    assert np.array_equal(make[0](storeState=False).applyOn(x), O.filt_stateless(b, a, x))
    if n > 3 * max(len(a), len(b)):
        assert np.array_equal(make[0](zeroPhase=True).applyOn(x), O.filt_zero_phase(b, a, x))


@settings(max_examples=20, deadline=None, derandomize=True)
@given(n=st.integers(64, 4000), seed=st.integers(0, 2**16), target=st.integers(2000, 40000),
       split=st.integers(1, 63))
def test_demodulators_and_strict_resample_port_equals_reference(n, seed, target, split):
    """demod_fm / demod_fmAD with carried state (demod_fm.py:29-51, :74-96), demod_am (demod_am.py:18-29)
    and the strict bwLim (comm.py:110-116)."""
    chunker, comm, constants, demod_am, demod_fm, filters = _ref()
    x = _signal(seed, n)
    fm = demod_fm.demod_fm()
    got = np.concatenate([fm.demod(x[:split].astype(np.complex128)), fm.demod(x[split:].astype(np.complex128))])
    w1, last = O.fm_discriminator(x[:split].astype(np.complex128), None)
    w2, _ = O.fm_discriminator(x[split:].astype(np.complex128), last)
    assert np.array_equal(got, np.concatenate([w1, w2]))
    ad = demod_fm.demod_fmAD()
    got = np.concatenate([ad.demod(x[:split].astype(np.complex128)), ad.demod(x[split:].astype(np.complex128))])
    w1, last = O.fm_angle_diff(x[:split].astype(np.complex128), None)
    w2, _ = O.fm_angle_diff(x[split:].astype(np.complex128), last)
    assert np.array_equal(got, np.concatenate([w1, w2]))
    r = x.real.astype(np.float64)
    assert np.array_equal(demod_am.demod_am().demod(r), O.am_envelope(r))
    fs = 60235
    if target <= fs and int(target * n / fs) >= 1:
        sig = comm.commSignal(fs, r).bwLim(target, True)
        want, rate = O.resample_strict(r, fs, target)
        assert sig.sampRate == rate and np.array_equal(np.asarray(sig.signal), want)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(lines=st.integers(3, 14), k=st.integers(1, 4), seed=st.integers(0, 2**16), sync_b=st.booleans(),
       missing=st.integers(0, 3), flat=st.booleans(), variant=st.sampled_from(["norm", "filter", "neg"]))
def test_sync_search_port_equals_reference(lines, k, seed, sync_b, missing, flat, variant):
    """decode_noaa.__correlateAndFindPeaks (decode_noaa.py:677-767): needle, normalised correlation,
    threshold from the K largest / smallest values, sequential group scan -- on envelopes with syncs
    removed, with plateaus (ties in the strict-< maximum), and through the three call variants the
    decoder uses (plain, hamming(492) zero-phase prefilter, +-0.5 needle)."""
    chunker, comm, constants, demod_am, demod_fm, filters = _ref()
    import importlib
    ref_noaa = importlib.import_module("directdemod.decode_noaa")
    their = object.__new__(ref_noaa.decode_noaa)
    search = getattr(their, "_decode_noaa__correlateAndFindPeaks")
    fs = 4160 * k
    rng = np.random.default_rng(seed)
    bits = O.NOAA_SYNCB if sync_b else O.NOAA_SYNCA
    needle = O.sync_needle(bits, fs)
    n = lines * (fs // 2) + int(rng.integers(0, fs // 4))
    sig = 0.35 + 0.03 * rng.standard_normal(n)
    if flat:
        sig = np.round(sig, 2)                       # plateaus: equal correlation values inside a group
    starts = [int(rng.integers(50, 400)) + j * (fs // 2) for j in range(lines)]
    for j in sorted(rng.choice(lines, size=min(missing, lines - 2), replace=False)) if missing else []:
        starts[j] = None
    for s in starts:
        if s is not None and s + len(needle) <= n:
            sig[s:s + len(needle)] = needle
    kw = {"norm": {}, "filter": {"useFilter": True}, "neg": {"usePosNeedle": False}}[variant]
    if variant == "filter" and n <= 3 * 492:
        return                                       # filtfilt refuses inputs shorter than its padding
    want_ref = np.asarray(search(comm.commSignal(fs, sig), bits, **kw))
    got, _ = O.find_syncs(sig, fs, bits, prefilter_taps=O.taps_hamming(492)[0] if variant == "filter" else None,
                          positive=variant != "neg")
    assert got.dtype.kind == "i" and np.array_equal(got, want_ref)


def test_afsk_front_end_port_equals_reference_intermediates():
    """decode_afsk1200.getMsg (decode_afsk1200.py:62-158) is one monolithic property, so the arrays the
    oracle restates -- the band-passed FM audio, the mark/space bank output of the pure-Python double loop
    (:106-142) and the bit-edge correlation (:145-158) -- are read out of the UNMODIFIED reference's own
    frame when getMsg returns (sys.setprofile; nothing of the reference is edited or copied).  This pins
    the vectorised restatement of the bank, which no fixture of the reference covers."""
    import importlib
    import sys
    chunker, comm, constants, demod_am, demod_fm, filters = _ref()
    ref_afsk = importlib.import_module("directdemod.decode_afsk1200")
    fs, bw, n = 96000, 48000, 16000
    rng = np.random.default_rng(12)
    t = np.arange(n) / fs
    bits = rng.integers(0, 2, int(n / fs * 1200) + 2)
    tone = np.where(bits[(t * 1200).astype(int)] == 1, 1200.0, 2200.0)
    audio = np.sin(2 * np.pi * np.cumsum(tone) / fs)
    x = (50 * np.exp(1j * 2 * np.pi * 3000 * np.cumsum(audio) / fs)
         + 1.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)

    class Src:
        sampFreq, length = fs, n

        def read(self, a, b=None):
            return x[a:b]

    seen = {}

    def grab(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "getMsg":
            loc = frame.f_locals
            if "sig" in loc:
                seen["audio"] = np.array(loc["sig"].signal)
            for k in ("binary_filter", "changes"):
                if k in loc:
                    seen[k] = np.array(loc[k])
    dec = ref_afsk.decode_afsk1200(Src(), 0.0, bw)
    sys.setprofile(grab)
    try:
        dec.getMsg
    except Exception:
        pass                                  # whatever the bit logic makes of 0.17 s of signal is not the point
    finally:
        sys.setprofile(None)
    assert {"audio", "binary_filter", "changes"} <= set(seen), sorted(seen)

    taps = O.taps_blackman_harris(151)[0]
    iq, rate = O.chain_stream(x, fs, 0.0, taps, bw, demod=False)
    assert rate == 48000
    fm, _ = O.fm_discriminator(iq, None)
    b, a = O.taps_butter(rate, 700, 2700, n=6, kind=O.FLT_BP)
    aud, _ = O.filt_stateful(b, a, fm, O.initial_zi(b, a))
    assert np.array_equal(aud, seen["audio"])                         # the same scipy calls: the same bits
    want_bf = O.afsk_bank(aud, bw)
    assert want_bf.shape == seen["binary_filter"].shape
    # the loop adds 40 products one by one, the restatement uses np.correlate: same numbers up to the order of the sums
    scale = np.max(np.abs(seen["binary_filter"]))
    assert np.max(np.abs(want_bf - seen["binary_filter"])) <= 1e-12 * scale
    assert np.all(want_bf[-40:] == 0) and np.all(seen["binary_filter"][-40:] == 0)
    spb = bw // 1200
    kernel = np.ones(spb)
    kernel[:spb // 2] = -1
    want_ch = np.correlate(np.sign(want_bf), kernel, mode="same") / spb
    near_zero = np.abs(seen["binary_filter"]) <= 1e-9 * scale           # a sign there may legitimately differ
    if not near_zero.any():
        assert np.array_equal(want_ch, seen["changes"])
    else:
        assert np.mean(want_ch == seen["changes"]) > 0.99


@settings(max_examples=60, deadline=None, derandomize=True)
@given(length=st.integers(0, 5000), size=st.integers(1, 1200), names=st.lists(st.sampled_from(["freqoffset", "bwlimFirst", "bwlimabcd", "x"]), max_size=6))
def test_product_chunker_equals_reference_chunker(length, size, names):
    """The drop-in's own chunker (product code, pure host logic) next to the reference's (chunker.py:21-84):
    the chunk list for drawn lengths and chunk sizes -- exact multiples still end with a full-size chunk, an
    empty source gives [[0, 0]] -- and the variable store: get with / without an initial value, set, KeyError."""
    ref_chunker = _ref()[0]
    from directdemod_b200 import chunker as ours
    a, b = ours.chunker(_Src(length), size), ref_chunker.chunker(_Src(length), size)
    assert [list(c) for c in a.getChunks] == [list(c) for c in b.getChunks]
    for k, name in enumerate(names):
        for obj in (a, b):
            if k % 3 == 0:
                obj.set(name, k)
        outcomes = []
        for obj in (a, b):
            try:
                outcomes.append(("ok", obj.get(name) if k % 2 else obj.get(name, 7 * k)))
            except KeyError:
                outcomes.append(("KeyError", None))
        assert outcomes[0] == outcomes[1], (name, k)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(kind=st.sampled_from(["bh", "hamming", "gauss", "roll", "butter", "remez"]), n=st.integers(1, 200),
       fs=st.integers(8000, 3000000), ca=st.floats(0.01, 0.49), cb=st.floats(0.01, 0.49),
       typ=st.integers(-1, 5), sigma=st.floats(0.5, 30.0), with_b=st.booleans())
def test_product_filter_designers_equal_reference_designers(kind, n, fs, ca, cb, typ, sigma, with_b):
    """The drop-in's designers (host side of filters.py:95-314; no device involved) next to the reference's:
    same coefficients bit for bit, same exception types for the same bad arguments (butter without cutoffB
    for band filters, unknown filter types, remez band / gain mismatches)."""
    chunker, comm, constants, demod_am, demod_fm, filters = _ref()
    from directdemod_b200 import filters as ours

    def both(make):
        out = []
        for mod in (ours, filters):
            try:
                f = make(mod)
                out.append(("ok", np.asarray(f.getB, dtype=np.float64), np.asarray(f.getA, dtype=np.float64)))
            except Exception as exc:                         # noqa: BLE001 -- the type is what is compared
                out.append((type(exc).__name__, None, None))
        assert out[0][0] == out[1][0], out
        if out[0][0] == "ok":
            assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])

    if kind == "bh":
        both(lambda m: m.blackmanHarris(n))
    elif kind == "hamming":
        both(lambda m: m.hamming(n))
    elif kind == "gauss":
        both(lambda m: m.gaussian(n, sigma))
    elif kind == "roll":
        both(lambda m: m.rollingAverage(n))
    elif kind == "butter":
        order = 1 + n % 8
        lo, hi = sorted((ca * fs, cb * fs))
        if hi - lo < 1e-3 * fs:
            hi = lo + 1e-3 * fs
        both(lambda m: m.butter(fs, lo, hi if with_b else None, n=order, typeFlt=typ))
    else:
        edge = ca * fs
        bands = [[0, edge], [min(edge * 1.3 + 1, fs / 2 - 1), fs / 2 - 1]]
        gains = [1, 0] if with_b else [1]                      # wrong number of gains -> ValueError in both
        both(lambda m: m.remez(fs, bands if typ != 5 else [[0, edge, 2 * edge]], gains, ntaps=8 + n % 120))


@settings(max_examples=25, deadline=None, derandomize=True)
@given(n=st.integers(2, 3000), seed=st.integers(0, 2**16),
       reads=st.lists(st.tuples(st.integers(-3, 3100), st.one_of(st.none(), st.integers(-3, 3100))), min_size=1, max_size=8),
       limit=st.one_of(st.none(), st.tuples(st.integers(0, 1500), st.integers(0, 3000))),
       given_fs=st.one_of(st.none(), st.integers(1000, 4000000)))
def test_product_sources_equal_reference_sources(tmp_path_factory, n, seed, reads, limit, given_fs):
    """The drop-in's file sources (product host code, source.py:40-230: 44-byte WAV header, raw .dat, the
    -127.5 conversion, read / limitData bounds and their ValueErrors) next to the reference's, on drawn
    files, index pairs and limits."""
    ref_shim.load()
    import importlib
    ref_source = importlib.import_module("directdemod.source")
    from directdemod_b200 import source as ours
    from tests.test_host import _write_wav_u8
    rng = np.random.default_rng(seed)
    pairs = rng.integers(0, 256, (n, 2), dtype=np.uint8)
    d = tmp_path_factory.mktemp("src")
    wav, dat = str(d / "iq.wav"), str(d / "iq.dat")
    _write_wav_u8(wav, pairs, 2048000)
    pairs.tofile(dat)

    def outcome(fn):
        try:
            return ("ok", fn())
        except Exception as exc:                             # noqa: BLE001 -- the type is what is compared
            return (type(exc).__name__, None)
    for cls, path in (("IQwav", wav), ("IQdat", dat)):
        a = getattr(ours, cls)(path, given_fs) if given_fs else getattr(ours, cls)(path)
        b = getattr(ref_source, cls)(path, given_fs) if given_fs else getattr(ref_source, cls)(path)
        assert (a.sampFreq, a.length, a.sourceType) == (b.sampFreq, b.length, b.sourceType)
        for phase in ("whole", "limited"):
            if phase == "limited":
                if limit is None:
                    break
                ra, rb = outcome(lambda: a.limitData(*limit)), outcome(lambda: b.limitData(*limit))
                assert ra[0] == rb[0], (cls, limit)
                if ra[0] != "ok":
                    break
                assert a.length == b.length
            for lo, hi in reads:
                ra = outcome(lambda: a.read(lo, hi) if hi is not None else a.read(lo))
                rb = outcome(lambda: b.read(lo, hi) if hi is not None else b.read(lo))
                assert ra[0] == rb[0], (cls, phase, lo, hi)
                if ra[0] == "ok":
                    assert ra[1].dtype == rb[1].dtype and np.array_equal(ra[1], rb[1]), (cls, phase, lo, hi)
