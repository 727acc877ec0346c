"""Generate tests/golden/noaa_pass.npz by running the UNMODIFIED reference decode_noaa
(crude sync, accurate sync subset, image) on the synthetic APT pass of tests/util.apt_iq.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference):
    python -m oracle.gen_golden_noaa
The input is not stored (229 MB): the tests regenerate it from the same seeded generator.
"""
import logging
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from tests.util import ArraySource, apt_iq  # noqa: E402

SEED, SECONDS, FS, F_OFF = 11, 14.0, 2048000, 30000.0


def main():
    ref_shim.load()
    sys.path.insert(0, ref_shim.REF_ROOT)
    from directdemod import decode_noaa
    logging.disable(logging.CRITICAL)
    x = apt_iq(SEED, SECONDS, fs=FS, f_off=F_OFF)
    t0 = time.time()
    dec = decode_noaa.decode_noaa(ArraySource(x, FS), F_OFF)
    syncA, syncB = dec.getCrudeSync()
    useful = dec.useful
    img = dec.getImage
    print("reference: useful=%d, %d syncA, image %s, %.1f s" % (useful, len(syncA), np.asarray(img).shape, time.time() - t0))
    # accurate sync of every window the reference accepts (decode_noaa.py:808-880), ~0.4 s per window
    t0 = time.time()
    acc = dec.getAccurateSync()
    print("reference: %d + %d accurate syncs, %.1f s" % (len(acc[0]), len(acc[4]), time.time() - t0))
    out = os.path.join(ROOT, "tests", "golden", "noaa_pass.npz")
    np.savez_compressed(out, seed=SEED, seconds=SECONDS, fs=FS, f_off=F_OFF, useful=useful,
                        syncA=np.asarray(syncA), syncB=np.asarray(syncB), image=np.asarray(img, dtype=np.uint8),
                        input_checksum=np.array([float(np.abs(x[::1000]).sum())]),
                        asyncA=np.asarray(acc[0], dtype=np.int64), asyncApk=np.asarray(acc[2], dtype=np.float64),
                        asyncAtime=np.asarray(acc[3], dtype=np.float64),
                        asyncB=np.asarray(acc[4], dtype=np.int64), asyncBpk=np.asarray(acc[6], dtype=np.float64),
                        asyncBtime=np.asarray(acc[7], dtype=np.float64))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
