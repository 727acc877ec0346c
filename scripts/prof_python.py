#!/usr/bin/env python
"""Where the host time of the drop-in API goes: cProfile of decode_noaa._audio (93 chunks through
commSignal), getCrudeSync and the line assembly of getImage on a device-resident 15-minute pass.
    python scripts/prof_python.py [audio,crude,image]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch

import bench_configs as B
from directdemod_b200 import constants, decode_noaa

torch.cuda.set_device(0)
fs = 2048000
x = B.apt_iq_device(900, fs)
src = B.DeviceSource(x, fs)
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["audio", "crude"]
for what in which:
    dec = decode_noaa.decode_noaa(src, 30000.0)
    if what == "image":
        dec.getImage                       # warm run; the syncs stay cached, the image is assembled again below
        dec._image = None
        dec._audOut = None                 # (the band-pass of :274 works on the audio in place)
        dec._audio(constants.NOAA_CRUDESYNCSAMPRATE, False)
        fn = lambda: dec.getImage
    else:
        fn = (lambda: dec._audio(constants.NOAA_CRUDESYNCSAMPRATE, False)) if what == "audio" else dec.getCrudeSync
        decode_noaa.decode_noaa(src, 30000.0).getCrudeSync()          # warm
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    fn()
    torch.cuda.synchronize()
    pr.disable()
    print("==== %s: %.2f ms wall" % (what, (time.perf_counter() - t0) * 1e3))
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
    print("\n".join(s.getvalue().splitlines()[:55]))
