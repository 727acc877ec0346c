#!/usr/bin/env python
"""Where the host time of the drop-in API goes: cProfile of decode_noaa._audio (93 chunks through
commSignal) and getCrudeSync on a device-resident 15-minute pass.   python scripts/prof_python.py"""
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
for what in ("audio", "crude"):
    dec = decode_noaa.decode_noaa(src, 30000.0)
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
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print("\n".join(s.getvalue().splitlines()[:45]))
