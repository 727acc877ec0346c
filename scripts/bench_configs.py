#!/usr/bin/env python
"""BASELINE.json configs C1..C5 on one GPU.  One JSON object per line; this is the per-config table
of DESIGN.md (the driver's bench is bench.py).

    python scripts/bench_configs.py [--quick] [--only c2,c4]      GPU legs only
    python bench.py --configs [--quick]                           plus the CPU legs

The CPU legs time the oracle port of the reference on a bounded slice of the same workload
(1 core).  Only bench.py may execute oracle/ outside the tests, so the oracle module is injected
by bench.py's cpu_baseline leg (``set_cpu_oracle``); run directly, this script skips the CPU legs.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from directdemod_b200 import afsk, comm, constants, decode_noaa, demod_fm, filters, shard
from directdemod_b200.fused import FusedChain
O = None          # the oracle module, injected by bench.py --configs (see the module docstring)


def set_cpu_oracle(module):
    global O
    O = module


def sync():
    torch.cuda.synchronize()


def wall(fn, reps=1):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    sync()
    return (time.perf_counter() - t0) / reps, r


def emit(**kw):
    print(json.dumps(kw), flush=True)


class DeviceSource:
    """IQ source whose samples already live on the GPU (source.py interface: sampFreq, length, read)."""

    def __init__(self, x, fs):
        self._x, self.sampFreq, self.length = x, fs, x.numel()

    def read(self, a, b=None):
        return self._x[a:b]


def apt_iq_device(seconds, fs=2048000, f_off=30000.0, dev_hz=17000.0, amp=60.0, noise=3.0, seed=3):
    """tests/util.apt_iq on the device, chunked (a 15-minute pass is 1.84e9 samples)."""
    rng = np.random.default_rng(seed)
    n_lines = int(np.ceil(seconds * 2)) + 1
    sync_a = np.array(constants.NOAA_SYNCA[:39]) * 233 + 11
    sync_b = np.array(constants.NOAA_SYNCB[:39]) * 233 + 11
    lines = []
    for ln in range(n_lines):
        img_a = (128 + 100 * np.sin(np.arange(909) / 30.0 + ln / 5.0) + rng.integers(-10, 10, 909)).clip(0, 255)
        img_b = (100 + 80 * np.cos(np.arange(909) / 50.0 - ln / 7.0) + rng.integers(-10, 10, 909)).clip(0, 255)
        tel = np.full(45, 30 + 25 * ((ln // 8) % 8))
        lines.append(np.concatenate([sync_a, np.full(47, 11), img_a, tel, sync_b, np.full(47, 244), img_b, tel]))
    words = torch.from_numpy(np.concatenate(lines).astype(np.float64) / 255.0).cuda()
    n = int(seconds * fs)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    acc = 0.0
    step = 1 << 24
    for a in range(0, n, step):
        b = min(n, a + step)
        idx = torch.arange(a, b, device="cuda", dtype=torch.float64)
        t = idx / fs
        widx = torch.clamp((t * 4160).to(torch.int64), max=words.numel() - 1)
        audio = words[widx] * torch.cos(2 * np.pi * 2400 * t)
        cs = torch.cumsum(audio, 0) + acc
        acc = float(cs[-1])
        ph = torch.remainder(f_off * t + dev_hz * cs / fs, 1.0) * (2 * np.pi)
        x[a:b] = torch.polar(torch.full_like(ph, amp), ph).to(torch.complex64)
        torch.view_as_real(x[a:b]).add_(torch.empty((b - a, 2), device="cuda").normal_(0, noise, generator=g))
    return x


def c1(quick):
    """tutorial/1_fm.py + 2_filter.py: whole-file chain through the drop-in API, host arrays in and out."""
    fs, n = 2048000, 2048000 * (2 if quick else 10)
    rng = np.random.default_rng(2)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)

    def ours():
        s = comm.commSignal(fs, x).filter(filters.blackmanHarris(151)).bwLim(30000) \
            .funcApply(demod_fm.demod_fm().demod).filter(filters.butter(30117, 200, 3200, typeFlt=constants.FLT_BP))
        return s.signal
    ours()
    t, y = wall(ours, 3)
    if O is None:
        emit(config="C1 tutorial chain (bh151 -> bwLim 30k -> FM -> butter BP), host in/out", samples=n,
             gpu_ms=round(t * 1e3, 2), gpu_msps=round(n / t / 1e6, 1))
        return
    ncpu = min(n, 4096000)
    t0 = time.perf_counter()
    b = O.taps_blackman_harris(151)[0]
    yy, _ = O.filt_stateful(b, [1], x[:ncpu].astype(np.complex128), O.initial_zi(b))
    yy, rate, _ = O.decimate(yy, fs, 30000)
    yy, _ = O.fm_discriminator(yy, None)
    bb, aa = O.taps_butter(rate, 200, 3200, n=6, kind=O.FLT_BP)
    O.filt_stateful(bb, aa, yy, O.initial_zi(bb, aa))
    tc = time.perf_counter() - t0
    emit(config="C1 tutorial chain (bh151 -> bwLim 30k -> FM -> butter BP), host in/out", samples=n,
         gpu_ms=round(t * 1e3, 2), gpu_msps=round(n / t / 1e6, 1), cpu_msps_1core=round(ncpu / tc / 1e6, 2),
         cpu_sample=ncpu)


def c2(quick):
    """NOAA APT decode of a synthetic pass at 2.048 Msps, device-resident capture."""
    fs, seconds = 2048000, (60 if quick else 900)
    t_gen, x = wall(lambda: apt_iq_device(seconds, fs))
    n = x.numel()
    warm = decode_noaa.decode_noaa(DeviceSource(x, fs), 30000.0)      # plans, workspaces, kernel attributes
    warm._correlateAndFindPeaks(warm._getAM(warm._audio(constants.NOAA_CRUDESYNCSAMPRATE, False)), constants.NOAA_SYNCA)
    del warm
    dec = decode_noaa.decode_noaa(DeviceSource(x, fs), 30000.0)
    audio = []
    for _ in range(5):
        dec = decode_noaa.decode_noaa(DeviceSource(x, fs), 30000.0)
        t, aud = wall(lambda: dec._audio(constants.NOAA_CRUDESYNCSAMPRATE, False))
        audio.append(t)
    t_audio = sorted(audio)[len(audio) // 2]
    t_am, am = wall(lambda: dec._getAM(aud))
    t_sa, sa = wall(lambda: dec._correlateAndFindPeaks(am, constants.NOAA_SYNCA))
    crude = []                       # whole crude sync, a fresh decoder each time (one run is at the mercy of the box)
    for _ in range(5):
        dec2 = decode_noaa.decode_noaa(DeviceSource(x, fs), 30000.0)
        crude.append(wall(lambda: dec2.getCrudeSync())[0])
    t_crude = sorted(crude)[len(crude) // 2]
    imgs = []
    for _ in range(3):
        dec2._image = None               # the syncs stay cached: only the line assembly is timed
        t, img = wall(lambda: dec2.getImage)
        imgs.append(t)
    t_img = sorted(imgs)[1]
    spacing = np.diff(np.asarray(sa))
    emit(config="C2 NOAA APT pass %d s @ 2.048 Msps (device-resident)" % seconds, samples=n,
         audio_ms=round(t_audio * 1e3, 2), audio_runs_ms=[round(t * 1e3, 2) for t in audio], audio_msps=round(n / t_audio / 1e6, 1), am_ms=round(t_am * 1e3, 2),
         syncA_ms=round(t_sa * 1e3, 2), crude_sync_total_ms=round(t_crude * 1e3, 2),
         crude_sync_runs_ms=[round(t * 1e3, 2) for t in crude],
         crude_sync_msps=round(n / t_crude / 1e6, 1), image_ms=round(t_img * 1e3, 2), image_runs_ms=[round(t * 1e3, 2) for t in imgs],
         image_shape=list(np.asarray(img).shape), useful=int(dec2.useful), n_syncA=int(len(sa)),
         syncA_spacing_ok=bool(np.all(np.abs(spacing[:-1] - 60235 / 2) <= 2)))
    nw = 24 if quick else 200
    dec2._syncA, dec2._syncB = dec2._syncA[:nw], dec2._syncB[:nw]
    accs = []
    for _ in range(3):
        dec2._asyncA = None              # forces the windows to be searched again
        t, res = wall(lambda: dec2.getAccurateSync())
        accs.append(t)
    t_acc = sorted(accs)[1]
    emit(config="C2 accurate sync, %d windows of 118152 samples (of ~%d per pass)" % (2 * nw, 4 * seconds),
         windows=2 * nw, total_ms=round(t_acc * 1e3, 1), runs_ms=[round(t * 1e3, 1) for t in accs], ms_per_window=round(t_acc * 1e3 / (2 * nw), 3),
         spacing=[int(v) for v in np.unique(res[1])][:6])
    if O is None:
        return
    # CPU: the oracle's crude-sync path on a 14 s slice (1 core)
    xs = x[:int(14 * fs)].cpu().numpy()
    t0 = time.perf_counter()
    audio, rate = O.chain_stream(xs, fs, 30000.0, O.taps_blackman_harris(151)[0], 60000)
    env = O.am_envelope_chunked(audio)
    O.find_syncs(env, rate, O.NOAA_SYNCA)
    O.find_syncs(env, rate, O.NOAA_SYNCB)
    tc = time.perf_counter() - t0
    emit(config="C2 CPU oracle: crude sync path on a 14 s slice, 1 core", samples=len(xs),
         cpu_msps_1core=round(len(xs) / tc / 1e6, 2))
    del x


def c3(quick):
    """AFSK1200 front end: 960 kHz IQ with bw = 48000, synthetic FM-modulated AFSK."""
    fs, bw, seconds = 960000, 48000, (30 if quick else 600)
    n = fs * seconds
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    acc_a = acc_p = 0.0
    bits = torch.randint(0, 2, (int(seconds * 1200) + 2,), device="cuda", generator=g)
    step = 1 << 24
    for a in range(0, n, step):
        b = min(n, a + step)
        t = torch.arange(a, b, device="cuda", dtype=torch.float64) / fs
        tone = torch.where(bits[(t * 1200).to(torch.int64)] == 1, 1200.0, 2200.0)
        pa = torch.cumsum(tone, 0) / fs + acc_a
        acc_a = float(pa[-1])
        audio = torch.sin(2 * np.pi * torch.remainder(pa, 1.0))
        pp = torch.cumsum(audio, 0) * (3000.0 / fs) + acc_p
        acc_p = float(pp[-1])
        x[a:b] = torch.polar(torch.full_like(pp, 50.0), torch.remainder(pp, 1.0) * (2 * np.pi)).to(torch.complex64)
        torch.view_as_real(x[a:b]).add_(torch.empty((b - a, 2), device="cuda").normal_(0, 1.0, generator=g))
    afsk.front_end(DeviceSource(x[:40000000], fs), 0.0, bw)      # first use: lazy kernel loading, filter analysis
    runs = []
    for _ in range(5):
        t, (sig, bf, ch) = wall(lambda: afsk.front_end(DeviceSource(x, fs), 0.0, bw))
        runs.append(t)
    t_fe = sorted(runs)[len(runs) // 2]
    emit(config="C3 AFSK1200 front end, %d s @ 960 kHz IQ -> 48 kHz (chain, FM, BP, bank, edges)" % seconds,
         samples=n, audio_samples=int(bf.numel()), total_ms=round(t_fe * 1e3, 2),
         runs_ms=[round(t * 1e3, 2) for t in runs], msps_iq=round(n / t_fe / 1e6, 1))
    if O is None:
        return
    ns = 48000 * 5
    aud = sig.signal[:ns]
    t0 = time.perf_counter()
    O.afsk_bank(aud, bw)
    tc = time.perf_counter() - t0
    emit(config="C3 CPU: mark/space bank, vectorised numpy restatement, 5 s of audio, 1 core "
                "(the reference's own pure-Python double loop is ~1000x slower)",
         samples=ns, cpu_msps_1core=round(ns / tc / 1e6, 3))
    del x


def c4(quick):
    """1 h @ 2.4 Msps, 1023-tap Remez + 8th-order Butterworth: one GPU's slab of an 8-way time split."""
    fs = 2400000
    n_total = fs * 3600
    n = (n_total // 8) if not quick else 100_000_000
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).normal_(0, 40)
    fir = filters.remez(fs, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    iir = filters.butter(fs, 100000, n=8)
    y = fir._apply_dev(x)              # warm: plans, attributes, allocator
    z = iir._apply_dev(y)
    del y, z
    t_fir, y = wall(lambda: fir._apply_dev(x))
    t_iir, z = wall(lambda: iir._apply_dev(y))
    # the same cascade as ONE equivalent filter (filters.cascade): a single overlap-save pass
    cas = filters.cascade([filters.remez(fs, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023),
                           filters.butter(fs, 100000, n=8)])
    w = cas._apply_dev(x)
    del w
    t_cas, w = wall(lambda: cas._apply_dev(x), 3)
    # ... and through the drop-in API: sig.filter(fir).filter(iir) (comm.py:80-92) fuses the same way
    f1 = filters.remez(fs, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    f2 = filters.butter(fs, 100000, n=8)
    comm.commSignal(fs, x).filter(f1).filter(f2).deviceSignal
    t_api, _ = wall(lambda: comm.commSignal(fs, x).filter(f1).filter(f2).deviceSignal, 3)
    emit(config="C4 slab of %d samples (1/8 of 1 h @ 2.4 Msps): remez1023 + butter8" % n, samples=n,
         api_filter_filter_ms=round(t_api * 1e3, 2), api_filter_filter_msps=round(n / t_api / 1e6, 1),
         fir_ms=round(t_fir * 1e3, 2), fir_msps=round(n / t_fir / 1e6, 1),
         fir_direct_form_equivalent_tflops=round(n * 4 * 1023 / t_fir / 1e12, 1), iir_ms=round(t_iir * 1e3, 2),
         iir_msps=round(n / t_iir / 1e6, 1), stage_by_stage_msps=round(n / (t_fir + t_iir) / 1e6, 1),
         cascade_taps=int(len(cas.getB)), cascade_ms=round(t_cas * 1e3, 2), cascade_msps=round(n / t_cas / 1e6, 1),
         cascade_hbm_gbs=round(n * 16 / t_cas / 1e9, 1),
         halo_samples_stage_by_stage=fir.lookback() + iir.lookback(), halo_samples_cascade=cas.lookback())
    del w
    if O is None:
        return
    nc = 8_000_000
    xs = x[:nc].cpu().numpy().astype(np.complex128)
    b1, a1 = np.asarray(fir.getB), [1.0]
    t0 = time.perf_counter()
    y1, _ = O.filt_stateful(b1, a1, xs, O.initial_zi(b1, a1))
    O.filt_stateful(iir.getB, iir.getA, y1, O.initial_zi(iir.getB, iir.getA))
    tc = time.perf_counter() - t0
    emit(config="C4 CPU oracle: same cascade on 8 M samples, 1 core", samples=nc, cpu_msps_1core=round(nc / tc / 1e6, 2))
    del x, y, z


def c5(quick):
    """256 independent 1 s captures @ 10 Msps, FM demod + decimate (D = 50); one GPU's share of 8."""
    fs, ncap = 10000000, (4 if quick else 32)
    n = fs
    caps = torch.empty((ncap, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(caps).normal_(0, 40)
    import scipy.signal as sps
    taps = sps.windows.blackmanharris(151)
    ch = FusedChain(taps, 50, 125000.0, fs)
    out = torch.empty(ch.out_count(n) + 1, dtype=torch.float32, device="cuda")

    out2 = torch.empty((ncap, ch.out_count(n)), dtype=torch.float32, device="cuda")

    def run():
        ch.apply_batch(caps, out=out2)
    run()
    t, _ = wall(run, 5)
    emit(config="C5 %d of 256 captures (1 s @ 10 Msps each), bh151 -> D=50 -> FM, one batched launch" % ncap,
         samples=ncap * n, total_ms=round(t * 1e3, 3), msps=round(ncap * n / t / 1e6, 1),
         hbm_gbs=round(ncap * n * (8 + 4 / 50) / t / 1e9, 1))
    if O is None:
        return
    xs = caps[0, :8_000_000].cpu().numpy()
    t0 = time.perf_counter()
    O.chain_stream(xs, fs, 125000.0, taps, fs / 50)
    tc = time.perf_counter() - t0
    emit(config="C5 CPU oracle: same chain on 8 M samples, 1 core", samples=len(xs), cpu_msps_1core=round(len(xs) / tc / 1e6, 2))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args(argv)
    torch.cuda.set_device(0)
    for name, fn in (("c1", c1), ("c2", c2), ("c3", c3), ("c4", c4), ("c5", c5)):
        if args.only and name not in args.only.split(","):
            continue
        try:
            fn(args.quick)
        except Exception as exc:
            emit(config=name, error="%s: %s" % (type(exc).__name__, exc))
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
