"""Import the UNMODIFIED reference (/root/reference) under modern numpy/scipy.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py and by the differential tests
that run in the build container; /root/reference does not exist on the GPU box, so
everything that imports this must skip when ``available()`` is False.  No reference code
is copied: this only patches renamed third-party symbols before importing it.
"""

from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("DDM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "directdemod"))


def _stub(name):
    mod = types.ModuleType(name)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def load():
    """Return the reference's ``directdemod`` package (importing it on first use)."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    if "directdemod" in sys.modules and getattr(sys.modules["directdemod"], "__ddm_ref__", False):
        return sys.modules["directdemod"]

    import numpy as np
    import scipy
    import scipy.fft
    import scipy.signal as sps

    sys.dont_write_bytecode = True          # the reference tree is read-only
    # window functions moved to scipy.signal.windows (filters.py:139,161,199,226)
    for name in ("blackmanharris", "hamming", "gaussian"):
        if not hasattr(sps, name):
            setattr(sps, name, getattr(sps.windows, name))
    # remez(Hz=) became remez(fs=) (filters.py:314)
    if not getattr(sps.remez, "__ddm_wrapped__", False):
        _remez = sps.remez

        def remez(*args, **kw):
            if "Hz" in kw:
                kw["fs"] = kw.pop("Hz")
            return _remez(*args, **kw)

        remez.__ddm_wrapped__ = True
        sps.remez = remez
    if not hasattr(np, "int"):
        np.int = int                        # decode_afsk1200.py:365
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                     # peakdetect.py:196
    if not hasattr(scipy, "ifft"):
        scipy.ifft = scipy.fft.ifft         # peakdetect.py:21
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.pylab", "pylab", "scipy.misc"):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(scipy, "misc"):
        scipy.misc = sys.modules.get("scipy.misc")

    sys.path.insert(0, REF_ROOT)
    try:
        import directdemod  # noqa: F401  (the reference package)
        from directdemod import chunker, comm, constants, demod_am, demod_fm, filters  # noqa: F401
    finally:
        sys.path.remove(REF_ROOT)
    sys.modules["directdemod"].__ddm_ref__ = True
    return sys.modules["directdemod"]
