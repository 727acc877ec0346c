"""BASELINE configs[0] (C1): the reference's tutorial scripts (tutorial/1_fm.py:21-37,
2_filter.py:21-48, 3_chunking.py:30-38) run with ``directdemod_b200`` imported in place of
``directdemod`` on a synthetic two-channel unsigned 8-bit WAV, against
 (a) tests/golden/c1_tutorials.npz -- what the UNMODIFIED reference produced for the same file
     (oracle/gen_golden_c1.py), and
 (b) the unmodified reference run live, wherever it is present (the build container's
     /root/reference, or the copy oracle/stage_ref.py staged for the GPU box)."""

import hashlib
import os

import numpy as np
import pytest

from oracle import ddoracle as O
from tests.util import TOL, afsk_wav_u8, tutorial_sequences, wrap_rel_rms

pytestmark = pytest.mark.gpu


def _ours():
    from directdemod_b200 import chunker, comm, constants, demod_fm, filters, source
    return source, comm, chunker, constants, filters, demod_fm


def _bp_floor(rate):
    """Measured float64 roundoff floor of the tutorial's 12th-order band-pass (ddm_iir_analyse)."""
    from directdemod_b200 import constants, filters
    return filters.butter(rate, 200, 3200, typeFlt=constants.FLT_BP).analysis()[1]


def _compare(got, want):
    for k in ("t1_rate", "t2_rate", "t3_rate", "t3_chunks"):
        assert int(got[k]) == int(want[k]), k
    assert int(got["t1_rate"]) == 30117                      # int(2048000 / 68)
    for k in ("t1", "t2_fm", "t3"):                          # FM output: phases, modulo 2 pi
        assert got[k].dtype == np.float64 and got[k].shape == want[k].shape, k
        assert wrap_rel_rms(got[k], want[k]) <= TOL, (k, wrap_rel_rms(got[k], want[k]))
    # tutorial 2 ends in a stateful 12th-order Butterworth band-pass in tf form: scipy's own float64
    # recursion has a roundoff floor for it (two scipy runs whose inputs differ in the last bit stay
    # that far apart), and its input here carries the fp32 rounding of the FM stage
    floor = _bp_floor(int(want["t2_rate"]))
    assert got["t2"].shape == want["t2"].shape
    assert O.rel_rms(got["t2"], want["t2"]) <= max(TOL, 10 * floor), (O.rel_rms(got["t2"], want["t2"]), floor)


def test_tutorial_sequences_match_reference_fixture(golden, tmp_path):
    g = golden("c1_tutorials")
    wav = str(tmp_path / "c1.wav")
    n = afsk_wav_u8(wav, int(g["seed"]), float(g["seconds"]))
    assert n == int(g["samples"]) and os.path.getsize(wav) == 44 + 2 * n
    assert hashlib.sha256(open(wav, "rb").read()).hexdigest() == str(g["wav_sha256"]), \
        "the synthetic WAV differs from the one the fixture was generated from"
    got = tutorial_sequences(_ours(), wav, int(g["chunk"]))
    assert int(got["t3_chunks"]) == 3
    _compare(got, g)


def test_tutorial_sequences_match_live_reference(tmp_path):
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("unmodified reference neither checked out nor staged (oracle/stage_ref.py)")
    ref_shim.load()
    from directdemod import chunker, comm, constants, demod_fm, filters, source
    wav = str(tmp_path / "c1b.wav")
    afsk_wav_u8(wav, seed=12, seconds=0.8)
    want = tutorial_sequences((source, comm, chunker, constants, filters, demod_fm), wav, 500001)
    got = tutorial_sequences(_ours(), wav, 500001)
    _compare(got, want)


def test_tutorial_3_with_raw_8bit_blocks_equals_the_complex_path(tmp_path):
    """The same chunk loop fed with source.readRaw (bytes straight into the fused kernel)."""
    source, comm, chunker, constants, filters, demod_fm = _ours()
    wav = str(tmp_path / "c1c.wav")
    afsk_wav_u8(wav, seed=13, seconds=0.6)
    outs = []
    for raw in (False, True):
        sigsrc = source.IQwav(wav)
        sigOut = comm.commSignal(sigsrc.sampFreq)
        bh, fm = filters.blackmanHarris(151), demod_fm.demod_fm()
        ck = chunker.chunker(sigsrc, 400001)
        for i in ck.getChunks:
            blk = sigsrc.readRaw(*i) if raw else sigsrc.read(*i)
            sig = comm.commSignal(sigsrc.sampFreq, blk, ck)
            sig.filter(bh)
            sig.bwLim(30000)
            sig.funcApply(fm.demod)
            sigOut.extend(sig)
        outs.append(sigOut.signal)
    assert outs[0].shape == outs[1].shape
    assert wrap_rel_rms(outs[1], outs[0]) <= TOL
