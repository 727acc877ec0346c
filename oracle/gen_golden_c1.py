"""Generate tests/golden/c1_tutorials.npz: the operator sequences of the reference's tutorial/1_fm.py,
2_filter.py and 3_chunking.py (BASELINE configs[0]) run on the UNMODIFIED reference over a synthetic
two-channel unsigned 8-bit WAV (tests/util.afsk_wav_u8 -- the SDRSharp sample recording is not part of
the reference checkout).  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_c1.py

The WAV is regenerated from its seed by the test (3 MB per second is too much to commit); the fixture
stores the reference's outputs plus a checksum of the WAV bytes so that a drifting generator is caught.
"""

from __future__ import annotations

import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from tests.util import afsk_wav_u8, tutorial_sequences  # noqa: E402

SEED, SECONDS, CHUNK = 2, 1.5, 1100001


def main():
    ref_shim.load()
    from directdemod import chunker, comm, constants, demod_fm, filters, source
    with tempfile.TemporaryDirectory() as tmp:
        wav = os.path.join(tmp, "c1.wav")
        n = afsk_wav_u8(wav, SEED, SECONDS)
        digest = hashlib.sha256(open(wav, "rb").read()).hexdigest()
        out = tutorial_sequences((source, comm, chunker, constants, filters, demod_fm), wav, CHUNK)
    path = os.path.join(ROOT, "tests", "golden", "c1_tutorials.npz")
    np.savez_compressed(path, seed=SEED, seconds=SECONDS, chunk=CHUNK, samples=n, wav_sha256=np.array(digest),
                        **{k: np.asarray(v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes;", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
