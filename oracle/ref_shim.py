"""Import the UNMODIFIED reference under modern numpy/scipy.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py, by the differential tests and
by the CPU arm of bench.py.  The reference is looked for at /root/reference (the build container)
and, failing that, at oracle/_ref/ -- unmodified module files staged there by oracle/stage_ref.py
(git-ignored, shipped to the GPU box by gpurun like a built binary).  Everything that imports this
must cope with ``available()`` being False.  This file only patches renamed third-party symbols
before importing the reference.
"""

from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("DDM_REFERENCE_ROOT", "/root/reference"), os.path.join(_HERE, "_ref")]


def _root():
    for cand in _CANDIDATES:
        if os.path.isfile(os.path.join(cand, "directdemod", "comm.py")):
            return cand
    return None


REF_ROOT = _root() or _CANDIDATES[0]


def available() -> bool:
    return _root() is not None


def where() -> str:
    """"checkout" (/root/reference), "staged" (oracle/_ref) or "absent"."""
    r = _root()
    return "absent" if r is None else ("staged" if os.path.abspath(r) == os.path.join(_HERE, "_ref") else "checkout")


def _stub(name):
    mod = types.ModuleType(name)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def load():
    """Return the reference's ``directdemod`` package (importing it on first use)."""
    ref_root = _root()
    if ref_root is None:
        raise RuntimeError("reference not present (neither %s nor %s)" % tuple(_CANDIDATES))
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

    sys.path.insert(0, ref_root)
    try:
        import directdemod  # noqa: F401  (the reference package)
        from directdemod import chunker, comm, constants, demod_am, demod_fm, filters  # noqa: F401
    finally:
        sys.path.remove(ref_root)
    sys.modules["directdemod"].__ddm_ref__ = True
    return sys.modules["directdemod"]
