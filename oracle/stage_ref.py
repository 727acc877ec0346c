"""Stage the UNMODIFIED reference modules of the hot path into oracle/_ref/ so that they travel to
the GPU box (TEST / BASELINE INFRASTRUCTURE ONLY).

/root/reference exists only in the build container.  ``oracle/_ref/`` is git-ignored (no reference
source ever enters the history) but not gpurun-ignored, so what this script puts there is shipped with
the snapshot exactly like the built ``libddemod.so``.  The reference is pure Python without a build
step (its setup.py's find_packages() finds nothing: the package has no __init__.py), so "building"
it is a byte-for-byte copy of the module files; oracle/ref_shim.py imports them from there when
/root/reference is absent.  Called from __graft_entry__.build(); also runnable by hand:

    python oracle/stage_ref.py
"""

from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("DDM_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
# the modules of the path plus what decode_noaa imports at module level
MODULES = ["comm", "filters", "chunker", "demod_fm", "demod_am", "constants", "decode_noaa", "decode_afsk1200",
           "source", "sink", "log", "peakdetect", "framechecksequence"]


def stage(verbose=False):
    """Copy the reference modules; returns the list of staged files ([] when the reference checkout
    is not present -- then whatever an earlier build staged is left alone)."""
    src = os.path.join(SRC_ROOT, "directdemod")
    if not os.path.isdir(src):
        return []
    dst = os.path.join(DST_ROOT, "directdemod")
    os.makedirs(dst, exist_ok=True)
    done = []
    for name in MODULES:
        a, b = os.path.join(src, name + ".py"), os.path.join(dst, name + ".py")
        if not os.path.exists(a):
            continue
        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
        done.append(b)
    lic = os.path.join(SRC_ROOT, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(DST_ROOT, "LICENSE"))
    with open(os.path.join(DST_ROOT, "README"), "w") as fh:
        fh.write("Unmodified copies of aerospaceresearch/DirectDemod modules, staged by oracle/stage_ref.py\n"
                 "for the CPU baseline on the GPU box.  Git-ignored; not part of this repository's source.\n")
    if verbose:
        print("staged %d reference modules into %s" % (len(done), dst))
    return done


if __name__ == "__main__":
    sys.exit(0 if stage(verbose=True) else 1)
