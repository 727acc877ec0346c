#!/usr/bin/env python
"""Where the time of decode_noaa._audio goes on a device-resident C2 pass (93 chunks of 20 M samples).

Three ways through the same 1 843 200 000 samples, each timed as (a) the host's time to queue the work and
(b) the time until the GPU is done:

    api     the reference-shaped loop: commSignal(...).offsetFreq().filter().bwLim().funcApply().bwLim() + extend
    chain   FusedChain.apply per chunk (torch.empty + one C call)
    raw     ddm_chain_apply_dev per chunk into one preallocated output (the C-ABI floor for 93 launches)
    single  the whole pass as one launch (what bench.py's headline times)

    python scripts/audio_breakdown.py [--seconds 900] [--reps 9]
    DDM_CHAIN_HALO_MEMCPY=1 python scripts/audio_breakdown.py      # halo carried by a copy node per chunk (A/B)
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from directdemod_b200 import chunker, constants, decode_noaa, filters
from directdemod_b200.fused import FusedChain


class DeviceSource:
    def __init__(self, x, fs):
        self._x, self.sampFreq, self.length = x, fs, x.numel()

    def read(self, a, b=None):
        return self._x[a:b]


def timed(fn, reps):
    q, tot = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        q.append((t1 - t0) * 1e3)
        tot.append((t2 - t0) * 1e3)
    return {"queue_ms": round(statistics.median(q), 3), "total_ms": round(statistics.median(tot), 3),
            "total_ms_min": round(min(tot), 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=int, default=900)
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--profile", action="store_true", help="cProfile of one pass through the API loop (stderr)")
    a = ap.parse_args()
    fs = 2048000
    n = a.seconds * fs
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    xr = torch.view_as_real(x)
    step = 1 << 26
    for i in range(0, n, step):
        xr[i:i + step].normal_(0, 40)
    src = DeviceSource(x, fs)
    bounds = chunker.chunker(src).getChunks
    taps = filters.blackmanHarris(151).getB
    decim = int(fs / 60000)

    def api():
        return decode_noaa.decode_noaa(src, 30000.0)._audio(constants.NOAA_CRUDESYNCSAMPRATE, False)

    def chain():
        ch = FusedChain(taps, decim, 30000.0, fs)
        return [ch.apply(x[lo:hi]) for lo, hi in bounds]

    out = torch.empty(n // decim + 8, dtype=torch.float32, device="cuda")
    ch_raw = FusedChain(taps, decim, 30000.0, fs)
    stream = torch.cuda.current_stream().cuda_stream
    got = C.c_int64()
    base, es = x.data_ptr(), 8

    def raw():
        ch_raw.reset()
        o = out.data_ptr()
        for lo, hi in bounds:
            ch_raw._fn_apply(ch_raw._h, base + lo * es, hi - lo, o, 1 << 40, C.byref(got), stream)
            o += got.value * 4

    def single():
        ch_raw.reset()
        ch_raw._fn_apply(ch_raw._h, base, n, out.data_ptr(), out.numel(), C.byref(got), stream)

    slices = [x[lo:hi] for lo, hi in bounds]
    m_chunk = (bounds[0][1] - bounds[0][0]) // decim + 1
    outs = [torch.empty(m_chunk, dtype=torch.float32, device="cuda") for _ in bounds]
    ch_pre = FusedChain(taps, decim, 30000.0, fs)

    def chain_prealloc():             # FusedChain.apply without the slicing and without torch.empty
        ch_pre.reset()
        for xs, o in zip(slices, outs):
            ch_pre.apply(xs, out=o)

    def only_slice():
        return [x[lo:hi] for lo, hi in bounds]

    def only_empty():
        return [torch.empty(m_chunk, dtype=torch.float32, device="cuda") for _ in bounds]

    res = {"samples": n, "chunks": len(bounds), "halo_carry": "copy node" if os.environ.get("DDM_CHAIN_HALO_MEMCPY") else "in kernel"}
    for name, fn in (("api", api), ("chain", chain), ("chain_prealloc", chain_prealloc), ("raw", raw),
                     ("single", single), ("only_slice", only_slice), ("only_empty", only_empty)):
        for _ in range(2):
            fn()
        res[name] = timed(fn, a.reps)
        res[name]["gsps"] = round(n / res[name]["total_ms"] / 1e6, 1)
    # the chunked result must not depend on how the halo travels: compare against the single launch
    ref = torch.empty_like(out)
    ch_raw.reset()
    ch_raw._fn_apply(ch_raw._h, base, n, ref.data_ptr(), ref.numel(), C.byref(got), stream)
    m = got.value
    raw()
    torch.cuda.synchronize()
    d = torch.remainder(out[:m] - ref[:m] + torch.pi, 2 * torch.pi) - torch.pi       # a wrap at +-pi is not an error
    res["chunked_vs_single_max_abs_rad"] = float(d.abs().max())
    print(json.dumps(res), flush=True)
    if a.profile:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        torch.cuda.synchronize()
        pr.enable()
        api()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
