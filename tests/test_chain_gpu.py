"""GPU parity of the fused chain (ddm_chain_* through the C ABI) against the oracle and the
golden fixtures generated from the unmodified reference."""

import numpy as np
import pytest

from oracle import ddoracle as O
from tests.util import TOL, fm_tone_c64, noise_c64, oracle_chain, random_cuts, wrap_rel_rms

pytestmark = pytest.mark.gpu


def run_chain(x, fs, f, taps, decim, cuts, demod=True, host=True):
    import torch
    from directdemod_b200.fused import FusedChain
    ch = FusedChain(taps, decim, f, fs, demod=demod)
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        if host:
            parts.append(np.array(ch.apply_host(x[a:b])))
        else:
            xd = torch.from_numpy(np.ascontiguousarray(x[a:b])).cuda()
            parts.append(ch.apply(xd).cpu().numpy())
    ch.close()
    return np.concatenate(parts)


@pytest.mark.parametrize("name", ["chain_noise_d34", "chain_fmtone_d34", "chain_fmtone_d68", "chain_noise_d50"])
def test_chain_matches_reference_golden(golden, name):
    g = golden(name)
    x, fs, f, bw = g["x"], int(g["fs"]), float(g["f_off"]), int(g["bw"])
    taps = O.taps_blackman_harris(151)[0]
    decim = int(fs / bw)
    for tag, cs in (("whole", len(x) + 1), ("c2500", 2500), ("c1111", 1111), ("c97", 97)):
        cuts = [a for a, _ in O.chunk_bounds(len(x), cs)] + [len(x)]
        fm = run_chain(x, fs, f, taps, decim, cuts, demod=True)
        assert fm.shape == g["fm_" + tag].shape, tag
        assert wrap_rel_rms(fm, g["fm_" + tag]) <= TOL, (tag, wrap_rel_rms(fm, g["fm_" + tag]))
        iq = run_chain(x, fs, f, taps, decim, cuts, demod=False)
        assert iq.shape == g["iq_" + tag].shape, tag
        assert O.rel_rms(iq, g["iq_" + tag]) <= TOL, (tag, O.rel_rms(iq, g["iq_" + tag]))


CASES = [
    # ntaps, decim, fs, f_off, n, n_cuts
    (151, 34, 2048000, 30000.0, 300000, 7),
    (151, 50, 10000000, -125000.0, 200000, 5),
    (151, 68, 2048000, 12345.678, 200000, 9),
    (151, 34, 2048000, 0.0, 100000, 3),        # no mixer
    (31, 2, 48000, 1000.0, 50000, 6),
    (7, 4, 40, 3.0, 5000, 12),
    (255, 32, 2400000, 100000.0, 150000, 4),
    (492, 64, 2048000, -30000.0, 150000, 4),
    (1, 2, 1000, 10.0, 4000, 3),               # single tap
    (151, 33, 2048000, 30000.0, 60000, 5),     # odd decimation: element-aligned blocks, even tiles
    (151, 17, 1024000, 30000.0, 90000, 6),     # 1.024 Msps recordings: D = 17, Q = 9
    (151, 35, 2100000, -20000.0, 90000, 4),
    (101, 25, 1500000, 12000.0, 90000, 7),
    (151, 1, 60235, 500.0, 20000, 4),          # no decimation -> general path (sliding-window tile)
    (151, 1, 60235, 0.0, 5000, 3),             # ... without mixer
    (1, 1, 1000, 10.0, 3000, 3),               # ... single tap
    (40, 3, 48000, 1000.0, 30000, 5),          # general path, odd D, tile kernel
    (600, 8, 2048000, 30000.0, 30000, 3),      # Q > 10 -> general path
]


@pytest.mark.parametrize("ntaps,decim,fs,f,n,ncuts", CASES)
@pytest.mark.parametrize("demod", [True, False])
def test_chain_matches_oracle_random_chunking(ntaps, decim, fs, f, n, ncuts, demod):
    taps = O.taps_blackman_harris(ntaps)[0] if ntaps > 1 else np.array([0.7])
    # keep the FM tone's sidebands inside the FIR passband (~fs/ntaps) so the demodulated
    # signal -- the denominator of the relative error -- is the tone, not residual noise
    f_mod = min(fs / decim / 40.0, 0.1 * fs / ntaps)
    x = fm_tone_c64(ntaps * 1000 + decim, n, fs, f, f_mod, 2.0) if demod \
        else noise_c64(ntaps * 1000 + decim, n)
    cuts = random_cuts(decim, n, ncuts, small=2)
    want, _ = oracle_chain(x, fs, f, taps, fs / decim, cuts, demod)
    got = run_chain(x, fs, f, taps, decim, cuts, demod, host=(ncuts % 2 == 0))
    assert got.shape == want.shape
    err = wrap_rel_rms(got, want) if demod else O.rel_rms(got, want)
    assert err <= TOL, err


def test_tiny_chunks_shorter_than_decimation_and_halo():
    # chunks of 1..40 samples with D = 34: many calls produce no output at all
    fs, f, decim = 2048000, 30000.0, 34
    taps = O.taps_blackman_harris(151)[0]
    x = fm_tone_c64(5, 3000, fs, f, 2000.0, 2.0)
    rng = np.random.default_rng(9)
    cuts = [0]
    while cuts[-1] < len(x):
        cuts.append(min(len(x), cuts[-1] + int(rng.integers(1, 41))))
    # the reference itself raises on an empty decimated chunk (demod_fm.py:44 indexes sig[-1]),
    # so compare the un-demodulated IQ, which it does define, plus FM against the whole-run
    want_iq, _ = oracle_chain(x, fs, f, taps, fs / decim, cuts, demod=False)
    got_iq = run_chain(x, fs, f, taps, decim, cuts, demod=False)
    assert got_iq.shape == want_iq.shape and O.rel_rms(got_iq, want_iq) <= TOL
    want_fm, _ = oracle_chain(x, fs, f, taps, fs / decim, [0, len(x)], demod=True)
    got_fm = run_chain(x, fs, f, taps, decim, cuts, demod=True)
    assert got_fm.shape == want_fm.shape and wrap_rel_rms(got_fm, want_fm) <= TOL


def test_empty_chunk_and_capacity_error():
    import torch
    from directdemod_b200 import _lib
    from directdemod_b200.fused import FusedChain
    ch = FusedChain(O.taps_blackman_harris(151)[0], 34, 30000.0, 2048000)
    assert ch.apply_host(np.zeros(0, dtype=np.complex64)).size == 0
    assert ch.position == (0, 0, False)
    x = torch.zeros(34 * 10, dtype=torch.complex64, device="cuda")
    small = torch.empty(3, dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError):
        ch.apply(x, out=small)
    out = ch.apply(x)
    assert out.numel() == 9 and ch.position == (340, 0, True)
    with pytest.raises(_lib.DdmError):
        FusedChain([1.0], 0, 0.0, 48000)
    ch.close()


def test_full_chunk_tone_property_and_chunk_invariance():
    """Size-independent checks at the reference's real chunk size (20 M samples): a pure tone
    at f_off + df must demodulate to the constant 2*pi*df*D/fs everywhere (including at
    global indices ~1.8e9 where fp32 phase would be useless), and splitting the stream at
    arbitrary points must not change the result."""
    import torch
    from directdemod_b200.fused import FusedChain
    fs, f, decim, df = 2048000, 30000.0, 34, 1500.0
    taps = O.taps_blackman_harris(151)[0]
    n = 20000000
    n_start = 1800000000                      # near the end of a 15-minute pass
    idx = torch.arange(n_start, n_start + n, device="cuda", dtype=torch.float64)
    ph = torch.remainder(idx * ((f + df) / fs), 1.0) * (2 * np.pi)
    x = torch.polar(torch.full_like(ph, 50.0), ph).to(torch.complex64)
    del idx, ph
    ch = FusedChain(taps, decim, f, fs)
    ch.set_position(n_start, 0, False)
    y = ch.apply(x)
    want = 2 * np.pi * df * decim / fs
    steady = y[10:]
    assert steady.numel() == (n + decim - 1) // decim - 1 - 10
    assert float((steady - want).abs().max()) < 2e-5
    # same stream in three uneven pieces
    ch2 = FusedChain(taps, decim, f, fs)
    ch2.set_position(n_start, 0, False)
    parts = [ch2.apply(x[:7000001]), ch2.apply(x[7000001:7000050]), ch2.apply(x[7000050:])]
    y2 = torch.cat(parts)
    assert y2.shape == y.shape
    assert float((y2 - y).abs().max()) < 1e-5
    ch.close()
    ch2.close()


def _u8_iq(seed, n, fs, f, f_mod, beta):
    """RTL-SDR style recording: unsigned 8-bit interleaved I/Q of an FM tone plus noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    ph = 2 * np.pi * f * t + beta * np.sin(2 * np.pi * f_mod * t)
    z = 70 * np.exp(1j * ph) + 6 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    iq = np.empty((n, 2), dtype=np.uint8)
    iq[:, 0] = np.clip(np.round(z.real + 127.5), 0, 255)
    iq[:, 1] = np.clip(np.round(z.imag + 127.5), 0, 255)
    return iq


@pytest.mark.parametrize("ntaps,decim,fs,f,n,ncuts", [
    (151, 34, 2048000, 30000.0, 400000, 6),
    (151, 50, 10000000, -125000.0, 300000, 5),
    (151, 34, 2048000, 0.0, 100000, 3),        # no mixer
    (492, 64, 2048000, -30000.0, 200000, 4),
    (151, 33, 2048000, 30000.0, 60000, 4),     # odd decimation
    (151, 35, 2100000, -20000.0, 150000, 5),
])
@pytest.mark.parametrize("demod", [True, False])
def test_u8_ingest_matches_oracle(ntaps, decim, fs, f, n, ncuts, demod):
    """The chain fed with raw unsigned 8-bit I/Q (source.py:117-118: u8 - 127.5, fused into the
    kernel) equals the oracle run on the converted complex64 samples, for random chunkings that
    include the fresh-stream hand-over from the general to the fused kernel."""
    import torch
    from directdemod_b200.fused import FusedChain
    iq = _u8_iq(ntaps + decim, n, fs, f, min(fs / decim / 40.0, 0.1 * fs / ntaps), 2.0)
    x = (iq[:, 0].astype(np.float32) - 127.5) + 1j * (iq[:, 1].astype(np.float32) - 127.5)
    x = x.astype(np.complex64)
    taps = O.taps_blackman_harris(ntaps)[0]
    cuts = random_cuts(decim + 1, n, ncuts, small=2)
    want, _ = oracle_chain(x, fs, f, taps, fs / decim, cuts, demod)
    ch = FusedChain(taps, decim, f, fs, demod=demod, in_format="cu8")
    parts = []
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        if i % 2:
            parts.append(np.array(ch.apply_host(iq[a:b])))
        else:
            parts.append(ch.apply(torch.from_numpy(iq[a:b].copy()).cuda()).cpu().numpy())
    got = np.concatenate(parts)
    assert got.shape == want.shape
    err = wrap_rel_rms(got, want) if demod else O.rel_rms(got, want)
    assert err <= TOL, err
    # state hand-over to the stand-alone operators works from a u8 halo too
    zi, last = ch.export_state()
    st = O.ChainState(taps)
    for a, b in zip(cuts[:-1], cuts[1:]):
        O.chain_chunk(x[a:b], fs, f, taps, fs / decim, st, demod=False)
    assert O.rel_rms(zi, st.zi) <= 1e-5
    ch.close()


def test_batch_of_independent_captures_in_one_launch():
    """BASELINE config 5 in miniature: every row is its own fresh stream; one launch."""
    import torch
    from directdemod_b200 import _lib
    from directdemod_b200.fused import FusedChain
    fs, f, decim, n, ncap = 10000000, -125000.0, 50, 120002, 5      # rows stay 16-byte aligned
    taps = O.taps_blackman_harris(151)[0]
    caps = np.stack([fm_tone_c64(100 + k, n, fs, f, 4000.0 + 300 * k, 2.0) for k in range(ncap)])
    ch = FusedChain(taps, decim, f, fs)
    l0 = _lib.launch_count()
    got = ch.apply_batch(torch.from_numpy(caps).cuda()).cpu().numpy()
    assert _lib.launch_count() - l0 == 1
    for k in range(ncap):
        want, _ = O.chain_stream(caps[k], fs, f, taps, fs / decim)
        assert got[k].shape == want.shape
        assert wrap_rel_rms(got[k], want) <= TOL, k
    # rows that are not 16-byte aligned, and odd decimation -> capture by capture, same results
    got1 = ch.apply_batch(torch.from_numpy(np.ascontiguousarray(caps[:2, :120001])).cuda()).cpu().numpy()
    want, _ = O.chain_stream(caps[1, :120001], fs, f, taps, fs / decim)
    assert got1[1].shape == want.shape and wrap_rel_rms(got1[1], want) <= TOL
    ch2 = FusedChain(taps, 33, f, fs)
    got2 = ch2.apply_batch(torch.from_numpy(caps[:2]).cuda()).cpu().numpy()
    want, _ = O.chain_stream(caps[1], fs, f, taps, fs / 33)
    assert wrap_rel_rms(got2[1], want) <= TOL


def test_batch_of_short_captures_crosses_capture_boundaries_inside_a_warp_range():
    """Many short captures at the NOAA decimation (warp-autonomous kernel): a warp's contiguous range of
    tiles spans several captures, each of which must start from the reference's fresh state."""
    import torch
    from directdemod_b200.fused import FusedChain
    fs, f, decim, n, ncap = 2048000, 30000.0, 34, 5002, 40
    taps = O.taps_blackman_harris(151)[0]
    caps = np.stack([fm_tone_c64(300 + k, n, fs, f, 900.0 + 70 * k, 2.0) for k in range(ncap)])
    got = FusedChain(taps, decim, f, fs).apply_batch(torch.from_numpy(caps).cuda()).cpu().numpy()
    for k in (0, 1, 7, 20, 39):
        want, _ = O.chain_stream(caps[k], fs, f, taps, fs / decim)
        assert got[k].shape == want.shape
        assert wrap_rel_rms(got[k], want) <= TOL, k
    iq = FusedChain(taps, decim, f, fs, demod=False).apply_batch(torch.from_numpy(caps).cuda()).cpu().numpy()
    want, _ = O.chain_stream(caps[13], fs, f, taps, fs / decim, demod=False)
    assert iq[13].shape == want.shape and O.rel_rms(iq[13], want) <= TOL


def test_host_entry_point_pipelines_long_chunks_in_pieces():
    """ddm_chain_apply_host splits chunks of >= 2^22 samples into four pieces whose copies run ahead of
    the kernels: same samples as the device entry point on the same chunk, cf32 and u8, carried state
    included (second chunk)."""
    import torch
    from directdemod_b200.fused import FusedChain
    fs, f, decim = 2048000, 30000.0, 34
    taps = O.taps_blackman_harris(151)[0]
    n = (1 << 22) + 12345
    x = fm_tone_c64(77, 2 * n, fs, f, 1700.0, 2.0)
    for fmt in ("cf32", "cu8"):
        if fmt == "cu8":
            data = np.clip(np.rint(np.stack([x.real, x.imag], axis=1) + 127.5), 0, 255).astype(np.uint8)
            dev_in = lambda a, b: torch.from_numpy(data[a:b]).cuda()
            host_in = lambda a, b: data[a:b]
        else:
            dev_in = lambda a, b: torch.from_numpy(x[a:b]).cuda()
            host_in = lambda a, b: x[a:b]
        cd, chost = FusedChain(taps, decim, f, fs, in_format=fmt), FusedChain(taps, decim, f, fs, in_format=fmt)
        for a, b in ((0, n), (n, 2 * n)):
            want = cd.apply(dev_in(a, b)).cpu().numpy()
            got = chost.apply_host(host_in(a, b))
            assert got.shape == want.shape and np.array_equal(got, want), (fmt, a)
        assert cd.position == chost.position


@pytest.mark.parametrize("fmt", ["cf32", "cu8"])
@pytest.mark.parametrize("decim,fs", [(34, 2048000), (33, 2048000), (20, 960000), (50, 10000000)])
def test_chunk_loop_mechanics_are_bit_neutral(monkeypatch, fmt, decim, fs):
    """The chunk loops' launch mechanics -- the delay line carried by the kernel itself (save_halo) or by
    a copy node, chunks shorter than the halo in between (which always take the copy), the trimmed
    warm-up tile of a warp's range, programmatic dependent launch on or off -- must not move a single
    bit: every variant of a ragged chunk loop gives the same bits, and the one-launch result of the same
    stream within the tolerance (the chunk invariance of the reference's stateful chain, filters.py:69 + comm.py:76,124 +
    demod_fm.py:44-48, which the oracle tests pin separately)."""
    import torch
    from directdemod_b200.fused import FusedChain
    rng = np.random.default_rng(decim)
    n = 3_000_000
    taps = O.taps_blackman_harris(151)[0]
    if fmt == "cu8":
        xd = torch.from_numpy(rng.integers(0, 256, 2 * n, dtype=np.uint8)).cuda()
        piece = lambda a, b: xd[2 * a:2 * b]
    else:
        xd = torch.from_numpy(noise_c64(rng, n)).cuda()
        piece = lambda a, b: xd[a:b]
    # even cut points (the fused kernels want 16-byte aligned chunk starts; others take the general path,
    # covered elsewhere), long chunks, chunks of a few halos, chunks shorter than the halo
    cuts = [0, 1_000_000, 1_000_100, 1_000_150, 1_700_000, 1_700_008, 2_400_000, 2_400_600, n]
    if fmt == "cu8":
        cuts = [c // 8 * 8 for c in cuts]

    def loop(env):
        for k in ("DDM_CHAIN_HALO_MEMCPY", "DDM_CHAIN_NO_PDL"):
            monkeypatch.delenv(k, raising=False)
        for k in env:
            monkeypatch.setenv(k, "1")
        ch = FusedChain(taps, decim, 30000.0, fs, in_format=fmt)          # the switches are read at creation
        parts = [ch.apply(piece(a, b)) for a, b in zip(cuts[:-1], cuts[1:])]
        ch.close()
        return torch.cat(parts)

    ch = FusedChain(taps, decim, 30000.0, fs, in_format=fmt)
    whole = ch.apply(piece(0, n))
    ch.close()
    base = loop(())
    assert base.shape == whole.shape
    # against the one-launch result: the same samples up to float32 rounding (odd D shifts the block
    # grid by one sample with the parity of the carried offset, so bit equality is not promised here)
    assert wrap_rel_rms(base.cpu().numpy(), whole.cpu().numpy()) <= TOL
    for env in (("DDM_CHAIN_HALO_MEMCPY",), ("DDM_CHAIN_NO_PDL",), ("DDM_CHAIN_HALO_MEMCPY", "DDM_CHAIN_NO_PDL")):
        got = loop(env)
        assert torch.equal(got, base), (env, float((got - base).abs().max()))
