#!/usr/bin/env python
"""Per-operator throughput on one GPU (CUDA events, device-resident inputs larger than L2 where
the operator streams).  Prints one JSON object per line; used for the table in DESIGN.md and to
pick the next kernel to optimise.  Not the driver's bench (that is bench.py)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from directdemod_b200 import _dev, _lib, afsk, comm, constants, demod_am, demod_fm, fftops, filters, sync
from directdemod_b200.fused import FusedChain

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def report(name, n, ms_min, ms_avg, bytes_per_sample=None, flops_per_sample=None, note=""):
    rec = {"op": name, "samples": n, "ms_min": round(ms_min, 4), "ms_avg": round(ms_avg, 4),
           "Msps": round(n / ms_min / 1e3, 1)}
    if bytes_per_sample:
        gbs = n * bytes_per_sample / ms_min / 1e6
        rec["GBps"] = round(gbs, 1)
        rec["hbm_frac"] = round(gbs / HBM, 4)
    if flops_per_sample:
        rec["TFLOPs"] = round(n * flops_per_sample / ms_min / 1e9, 2)
    if note:
        rec["note"] = note
    print(json.dumps(rec), flush=True)


def noise(n, cplx=True):
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    if cplx:
        x = torch.empty(n, dtype=torch.complex64, device="cuda")
        torch.view_as_real(x).normal_(0, 40, generator=g)
        return x
    return torch.empty(n, dtype=torch.float32, device="cuda").normal_(0, 1, generator=g)


def main():
    torch.cuda.set_device(0)
    n = 200_000_000
    x = noise(n)
    import scipy.signal as sps
    bh = sps.windows.blackmanharris(151)
    # fused chains
    for tag, fs, f, d in (("fused chain D=34 (C2)", 2048000, 30000.0, 34), ("fused chain D=50 (C5)", 10000000, 125000.0, 50)):
        ch = FusedChain(bh, d, f, fs)
        out = torch.empty(ch.out_count(n) + 1, dtype=torch.float32, device="cuda")
        def run():
            ch.set_position(0, 0, False)
            ch.apply(x, out=out)
        report(tag, n, *timeit(run), bytes_per_sample=8 + 4 / d)
    # unfused operators
    xm = x.clone()
    def mix():
        _lib.check(_lib.lib().ddm_mix_cf32(0, _dev.ptr(xm), n, 30000.0, 2048000.0, 0, _dev.stream_ptr(0)), "mix")
    report("mixer (in place)", n, *timeit(mix), bytes_per_sample=16)
    fm = demod_fm.demod_fm(storeState=False)
    report("fm_demod", n, *timeit(lambda: fm._demod_dev(x)), bytes_per_sample=12)
    nf = 100_000_000
    xf = x[:nf]
    for tag, flt, flops in (("FIR bh151 cf32", filters.blackmanHarris(151), 4 * 151),
                            ("FIR remez1023 cf32 (C4)", filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023), 4 * 1023),
                            ("IIR butter8 LP cf32 parallel (C4)", filters.butter(2400000, 100000, n=8), 2 * 34 * 2),
                            ("IIR butter6 LP f32 parallel", filters.butter(20800, 1200), 2 * 26)):
        xin = xf if " cf32" in tag else noise(nf, False)
        flt._apply_dev(xin[:1000000])
        report(tag, nf, *timeit(lambda: flt._apply_dev(xin), reps=3, warm=1), bytes_per_sample=16 if xin.is_complex() else 8,
               flops_per_sample=flops, note=str(flt.info()))
    fbp = filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP)
    xr = noise(4_000_000, False)
    report("IIR butter BP order12 f32 sequential (auto)", xr.numel(), *timeit(lambda: fbp._apply_dev(xr), reps=2, warm=1),
           bytes_per_sample=8, note=str(fbp.info()))
    fbp.setIIRMode(1)
    xr2 = noise(54_000_000, False)
    report("IIR butter BP order12 f32 parallel (forced)", xr2.numel(), *timeit(lambda: fbp._apply_dev(xr2), reps=2, warm=1),
           bytes_per_sample=8)
    # AM / resample / sync at the C2 crude-rate size
    na = 54_211_765
    aud = noise(na, False).abs_()
    report("AM hilbert chunked 240000 (C2)", na, *timeit(lambda: fftops.hilbert_envelope(aud, 240000), reps=3, warm=1), bytes_per_sample=8)
    xs = aud[:588235].contiguous()
    report("resample 588235->203127", 588235, *timeit(lambda: fftops.resample(xs, 203127), reps=5, warm=2))
    needle = sync.sync_needle(constants.NOAA_SYNCA, 60235)
    report("ncc 560-tap needle (C2 crude)", na, *timeit(lambda: sync.correlate(aud, needle), reps=3, warm=1), bytes_per_sample=12)
    cor = sync.correlate(aud, needle)
    t0 = time.perf_counter()
    try:
        pk, thr = sync.pick_peaks(cor, 60235, len(needle))
        npk = len(pk)
    except Exception as e:  # noise has no peaks structure; still times the selection
        npk = str(e)
    report("pick_peaks (radix select + compaction + host scan)", na, (time.perf_counter() - t0) * 1e3, (time.perf_counter() - t0) * 1e3, note=str(npk))
    win = noise(118152, False).abs_()
    needle2 = sync.sync_needle(constants.NOAA_SYNCA, 2048000)
    report("ncc 19680-tap needle, one accurate-sync window", 118152, *timeit(lambda: sync.correlate(win, needle2), reps=5, warm=2))
    xa = noise(28_800_000, False)
    report("afsk bank 40 taps (C3)", xa.numel(), *timeit(lambda: afsk.mark_space_bank(xa, 48000), reps=3, warm=1), bytes_per_sample=8)


if __name__ == "__main__":
    main()
