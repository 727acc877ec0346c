#!/usr/bin/env python
"""IIR (butter-8 low-pass, cf32) at the reference's call granularity: device time per chunk for the
segment-parallel kernel as a function of warps per SM sub-partition (DDM_IIR_WARPS, read per launch).
    python scripts/iir_sweep.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from directdemod_b200 import filters

torch.cuda.set_device(0)
for n in (20000000, 100000000):
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).normal_(0, 40)
    for wps in ("auto", "1", "2", "3", "4", "6", "8"):
        os.environ.pop("DDM_IIR_WARPS", None)
        if wps != "auto":
            os.environ["DDM_IIR_WARPS"] = wps
        f = filters.butter(2400000, 100000, n=8)
        for _ in range(3):
            f._apply_dev(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f._apply_dev(x)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(json.dumps({"op": "iir8 cf32", "n": n, "warps_per_subpartition": wps, "ms_min": round(ts[0], 4),
                          "ms_med": round(ts[len(ts) // 2], 4), "gsps": round(n / ts[len(ts) // 2] / 1e6, 1),
                          "warmup": f.info()[1]}), flush=True)
    del x
