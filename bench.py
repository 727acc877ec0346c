#!/usr/bin/env python
"""Benchmark of the hot path: IQ Msamples/s through shift -> FIR -> decimate -> FM demod.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1] -- one NOAA APT pass, 15 min at
2.048 Msps = 1 843 200 000 cf32 samples (14.7 GB, far larger than the 126 MB L2, so no L2
flush is needed between steps), chain offsetFreq(30 kHz) -> blackmanHarris(151) ->
bwLim(60000) [decimate by 34] -> demod_fm, i.e. decode_noaa.__audio (decode_noaa.py:600-629).
A step is one pass of the chain over the whole capture.  With N GPUs every rank holds its own
independent capture (weak scaling, no data-path collective); `value` is the total samples of
all ranks divided by the slowest rank's device time.

Prints ONE JSON line (rank 0).  `value` is timed with the capture resident in HBM; `e2e`
is the same metric through the C-ABI host entry point (ddm_chain_apply_host) with pinned
HOST input, chunked at the reference's PROC_CHUNKSIZE = 20 M samples, H2D and D2H inside the
timed region.  Sub-records of the same line (every N, incl. 1): `timeshard_chain` (ONE stream of
8 x the pass split in time over the N ranks, NCCL halo exchange inside the timed region),
`c4_cascade` (BASELINE configs[3]: remez-1023 -> butter-8 on a 1.08 G-sample slab per rank, one
stream split in time) -- both with a seam self-check against a single-rank recomputation -- and
`c5_captures` (BASELINE configs[4]: 256 independent captures dealt over the ranks).

`--impl reference` times the UNMODIFIED reference (the modules oracle/stage_ref.py staged into
oracle/_ref/, imported through oracle/ref_shim.py) running decode_noaa.__audio's chunk loop
(decode_noaa.py:613-627) on the host cores, one independent stream per process; only when
nothing is staged does it fall back to the oracle port and says so (cpu_baseline.kind).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 2048000
F_OFF = 30000.0
BW = 60000
NTAPS = 151
DECIM = int(FS / BW)                       # 34  (comm.py:119)
PASS_SECONDS = 900
N_PASS = FS * PASS_SECONDS                 # 1 843 200 000
CHUNK = 20000000                           # constants.PROC_CHUNKSIZE
CRUDE_RATE = 40960                         # constants.NOAA_CRUDESYNCSAMPRATE (second, non-strict bwLim: j = 1)
ALG_BYTES_PER_SAMPLE = 8.0 + 4.0 / DECIM   # cf32 in + f32 out every D samples
METRIC = "IQ Msps (shift->FIR->decim->FM demod)"
UNIT = "Msamples/s"
E2E_SAMPLES = N_PASS // 4                  # end-to-end leg: a quarter pass per rank at EVERY N (PCIe bound;
                                           # 8 ranks x 3.7 GB of pinned host memory stays affordable)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # CUDA-core FP32 FMA peak of a B200 at 1965 MHz

# C4 (BASELINE configs[3]) and C5 (configs[4])
C4_FS = 2400000
C4_SLAB = C4_FS * 3600 // 8                # 1.08e9 samples: one rank's share of the hour on 8 GPUs
C4_FLOPS_DIRECT = 1023 * 4 + (9 + 8) * 4   # direct-form flops per complex sample
C5_CAPTURES, C5_LEN, C5_FS, C5_BW = 256, 10000000, 10000000, 200000


def _traffic(n):
    """DRAM bytes per launch of the fused kernel from the committed `ncu --set full` capture of this
    round's kernel (profiles/r02_traffic.json) when it was taken at this launch size; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            t = json.load(fh)
        if int(t["samples_per_launch"]) == int(n):
            return t
    except Exception:
        pass
    return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.max_sm = None
        self.reasons = set()
        self.ok = False

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            self.ok = True
            while not self.stop_flag.is_set():
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = get_reasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.002)
        except Exception:
            self.ok = False

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["unavailable"]}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(s)}


def taps_bh151():
    import scipy.signal.windows as w
    return w.blackmanharris(NTAPS)          # filters.py:139 (coefficient design is host side)


def workload_config(n_gpus):
    """The same dict for both arms (the driver compares them key by key)."""
    return {
        "workload": "NOAA APT pass (BASELINE configs[1]): %d s @ %d sps = %d cf32 samples per GPU; "
                    "offsetFreq(%g) -> blackmanHarris(%d) -> bwLim(%d) [D=%d] -> demod_fm"
                    % (PASS_SECONDS, FS, N_PASS, F_OFF, NTAPS, BW, DECIM),
        "samples_per_gpu": N_PASS, "decim": DECIM, "ntaps": NTAPS, "chunk": CHUNK,
        "l2": "input per step is 14.7 GB (>> 126 MB L2); no flush needed",
        "sharding": "independent capture per GPU, no data-path collective" if n_gpus > 1 else "single GPU",
    }


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference on the host cores
# --------------------------------------------------------------------------------------
def _cpu_chunk_input(seed, n):
    import numpy as np
    rng = np.random.default_rng(seed)
    x = np.empty(n, dtype=np.complex64)
    step = 1 << 22
    for a in range(0, n, step):
        b = min(n, a + step)
        x.real[a:b] = rng.standard_normal(b - a, dtype=np.float32) * 40
        x.imag[a:b] = rng.standard_normal(b - a, dtype=np.float32) * 40
    return x


_CPU = {}


class _LoopSource:
    """Signal source with the reference's interface (source.py: sampFreq, length, read(a, b)) that
    serves every chunk from one resident buffer: file reading is not part of the measured path."""

    def __init__(self, x, chunks):
        self._x = x
        self.sampFreq = FS
        self.length = len(x) * chunks

    def read(self, a, b=None):
        return self._x[:b - a]


def _cpu_kind():
    from oracle import ref_shim
    return "reference" if ref_shim.available() else "port"


def _cpu_init(n):
    """Per-process setup: synthesise one chunk of input, import the reference (or the port)."""
    seed = os.getpid()
    _CPU["x"] = _cpu_chunk_input(seed, n)
    _CPU["n"] = n
    _CPU["kind"] = _cpu_kind()
    if _CPU["kind"] == "reference":
        import logging
        from oracle import ref_shim
        ref_shim.load()
        logging.disable(logging.CRITICAL)
    _cpu_step(1, warm=True)


def _cpu_step(reps, warm=False):
    """`reps` chunks through the chain exactly as decode_noaa.__audio (decode_noaa.py:613-627) runs
    them for the crude sync -- fresh filter / demodulator / chunker / output signal per pass, the
    accumulated audio extended chunk by chunk; returns the seconds it took."""
    x, n = _CPU["x"], _CPU["n"]
    if warm:
        x, n = x[:200000], 200000
    if _CPU["kind"] == "reference":
        from directdemod import chunker, comm, demod_fm, filters           # the staged, unmodified reference
        src = _LoopSource(x, reps)
        t0 = time.perf_counter()
        audio_out = comm.commSignal(CRUDE_RATE)
        bh = filters.blackmanHarris(NTAPS)
        fm = demod_fm.demod_fm()
        ck = chunker.chunker(src, n)
        for i in ck.getChunks:
            sig = comm.commSignal(src.sampFreq, src.read(*i), ck).offsetFreq(F_OFF).filter(bh) \
                .bwLim(BW, uniq="First").funcApply(fm.demod).bwLim(CRUDE_RATE, False)
            audio_out.extend(sig)
        return time.perf_counter() - t0
    from oracle import ddoracle as O
    taps = taps_bh151()
    t0 = time.perf_counter()
    st = O.ChainState(taps)
    for _ in range(reps):
        O.chain_chunk(x, FS, F_OFF, taps, BW, st)
    return time.perf_counter() - t0


def _cpu_sample_text(kind, reps, n, procs, cores):
    what = ("the unmodified reference (oracle/_ref: commSignal.offsetFreq.filter(blackmanHarris).bwLim.funcApply("
            "demod_fm).bwLim + extend, decode_noaa.py:613-627)" if kind == "reference" else
            "oracle port of the reference chain (oracle/ddoracle.chain_chunk; no staged reference found)")
    return "%d chunk(s) x %d samples per process through %s, %d process(es), host has %d logical cores" \
        % (reps, n, what, procs, cores)


def cpu_baseline_single(reps=3, n=CHUNK):
    _cpu_init(n)
    sec = _cpu_step(reps)
    kind = _CPU["kind"]
    _CPU.clear()
    return {"value": round(reps * n / sec / 1e6, 3), "unit": UNIT, "cores": 1, "kind": kind,
            "sample": _cpu_sample_text(kind, reps, n, 1, os.cpu_count() or 1)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all usable host
    cores, one independent stream per process (the reference is single threaded per stream).
    A step = every process pushes one PROC_CHUNKSIZE chunk through the chain; the step time is the
    slowest process (input synthesis is outside the timed section)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16 << 30
    cores = os.cpu_count() or 1
    n = CHUNK
    per_worker = n * (8 + 16 * 5 + 8)                  # c64 input + complex128 temporaries
    workers = max(1, min(cores, args.ref_procs, int(avail * 0.5 // per_worker)))
    ctx = mp.get_context("fork")
    total = 0.0
    kind = _cpu_kind()
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(n,)) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_step, [1] * workers, chunksize=1)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_step, [1] * workers, chunksize=1)
            total += time.perf_counter() - t0
    value = args.steps * workers * n / total / 1e6
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": workers, "kind": kind,
                         "sample": _cpu_sample_text(kind, 1, n, workers, cores) + "; per step"},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def _cpus_near_gpu(torch, dev_index):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when that cannot be told."""
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(path) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return (node, cpus) if cpus else None
    except Exception:
        return None


class _Env:
    """Rank bookkeeping + device-timed sections shared by the legs of the bench."""

    def __init__(self, torch, dist, world, rank, local, dev):
        self.torch, self.dist = torch, dist
        self.world, self.rank, self.local, self.dev = world, rank, local, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def timed(self, fn, reps, warm=2):
        """Average device time (ms) of fn over `reps` back-to-back calls bracketed by barrier +
        synchronize, max over ranks."""
        torch = self.torch
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks([e0.elapsed_time(e1) / reps])[0]

    def fill_noise(self, x, seed):
        torch = self.torch
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(seed)
        xr = torch.view_as_real(x) if x.is_complex() else x
        flat = xr.reshape(-1)
        slab = 1 << 27
        for a in range(0, flat.numel(), slab):
            flat[a:a + slab].normal_(0.0, 40.0, generator=gen)

    def free_bytes(self):
        self.torch.cuda.empty_cache()
        return self.torch.cuda.mem_get_info(self.dev)[0]


def leg_timeshard_chain(env, args, peak):
    """ONE stream of 8 x the C2 pass (14.7 G samples), split in time over the ranks (strong scaling):
    TimeShardedChain.run = NCCL neighbour exchange of the raw-input halo + the fused launch."""
    torch = env.torch
    from directdemod_b200 import _lib, shard
    n_total = 8 * N_PASS
    fit = (env.free_bytes() - (6 << 30)) // 9 * env.world          # 8 B in + out + slack per sample
    note = None
    if n_total > fit:
        n_total = int(fit // (DECIM * env.world) * (DECIM * env.world))
        note = "stream shortened to fit this GPU's free memory"
    ts = shard.TimeShardedChain(taps_bh151(), DECIM, F_OFF, FS, n_total, env.rank, env.world, device=env.local)
    n = ts.end - ts.start
    x = torch.empty(n, dtype=torch.complex64, device=env.dev)
    env.fill_noise(x, 77 + env.rank)
    l0 = _lib.launch_count()
    y = ts.run(x)
    per_run = _lib.launch_count() - l0
    reps = max(3, min(args.steps, 10))
    ms = env.timed(lambda: ts.run(x), reps)
    send = x[-ts.halo_len:].contiguous()
    ex_ms = env.timed(lambda: shard.exchange_halo(send, env.rank, env.world), 20) if env.world > 1 else 0.0
    chk = ts.boundary_check(x, y, width=2000)
    err, neq, cnt = env.max_over_ranks([chk["max_abs_err"], 0.0 if chk["bit_equal"] else 1.0, chk["samples"]])
    rec = {
        "workload": "ONE stream of %d cf32 samples (8 x the C2 pass) split in time over %d rank(s); chain as the "
                    "headline; halo = %d raw samples per seam over NCCL send/recv" % (n_total, env.world, ts.halo_len),
        "scaling": "strong", "samples_total": n_total, "samples_per_rank": n, "ms": round(ms, 4),
        "msps_total": round(n_total / ms / 1e3, 1),
        "exchange_halo_us": round(ex_ms * 1e3, 1), "halo_samples": ts.halo_len, "launches_per_run": per_run,
        "roofline": {"bound": "hbm", "achieved": round(n * ALG_BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9, 1), "peak": peak,
                     "unit": "GB/s per GPU", "frac": round(n * ALG_BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9 / peak, 4)},
        "timeshard_parity": bool(err <= 1e-6) if env.world > 1 else None,
        "parity": {"check": "+-2000 outputs around every seam recomputed by one chain on one rank", "max_abs_err_rad": err,
                   "tolerance_rad": 1e-6, "bit_equal": neq == 0.0, "samples_per_seam": int(cnt)},
    }
    if note:
        rec["note"] = note
    del x, y, ts
    return rec, per_run * reps


def leg_c4_cascade(env, args, peak):
    """BASELINE configs[3]: 1 h @ 2.4 Msps through remez-1023 -> butter-8 as ONE stream split in time;
    every rank holds the 1.08 G-sample slab of the 8-GPU split (on N < 8 ranks: the first N/8 of the
    hour), halo = (ntaps-1) + IIR warm-up raw samples per seam over NCCL."""
    torch = env.torch
    import numpy as np
    from directdemod_b200 import _lib, filters, shard
    slab = C4_SLAB
    fit = (env.free_bytes() - (6 << 30)) // 26                      # in + FIR out + IIR out (+ slack), 8 B each
    note = None
    if slab > fit:
        slab = int(fit)
        note = "slab shortened to fit this GPU's free memory"
    n_total = slab * env.world
    fir = filters.remez(C4_FS, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    iir = filters.butter(C4_FS, 100000, n=8)
    b1, b2, a2 = np.asarray(fir.getB, dtype=np.float64), np.asarray(iir.getB), np.asarray(iir.getA)
    tf = shard.TimeShardedFilters([fir, iir], n_total, env.rank, env.world)
    n = tf.end - tf.start
    x = torch.empty(n, dtype=torch.complex64, device=env.dev)
    env.fill_noise(x, 99 + env.rank)
    l0 = _lib.launch_count()
    y = tf.run(x)
    per_run = _lib.launch_count() - l0
    reps = max(3, min(args.steps, 5))
    ms = env.timed(lambda: tf.run(x), reps, warm=1)
    ex_ms = 0.0
    if env.world > 1:
        send = x[-tf.halo_len:].contiguous()
        ex_ms = env.timed(lambda: shard.exchange_halo(send, env.rank, env.world), 20)
    chk = tf.boundary_check(x, y, lambda: [filters.filter(b1, [1]), filters.filter(b2, a2)], width=2000)
    err, cnt = env.max_over_ranks([chk["max_rel_err"], chk["samples"]])
    gbs = n * 16.0 / (ms * 1e-3) / 1e9
    rec = {
        "workload": "C4 (BASELINE configs[3]): remez(1023 taps) -> butter(n=8) over ONE %d-sample cf32 stream @ 2.4 Msps "
                    "split in time, %d samples per rank, %d rank(s)" % (n_total, n, env.world),
        "scaling": "weak", "samples_total": n_total, "samples_per_rank": n, "ms": round(ms, 4),
        "msps_total": round(n_total / ms / 1e3, 1), "msps_per_gpu": round(n / ms / 1e3, 1),
        "exchange_halo_us": round(ex_ms * 1e3, 1), "halo_samples": tf.halo_len, "launches_per_run": per_run,
        "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s per GPU",
                     "frac": round(gbs / peak, 4), "algorithmic_bytes_per_sample": 16.0,
                     "direct_form_equivalent_tflops_per_gpu": round(n * C4_FLOPS_DIRECT / (ms * 1e-3) / 1e12, 1),
                     "fp32_peak_tflops": round(FP32_PEAK_TFLOPS, 1),
                     "note": "the FIR runs as an overlap-save FFT, so the direct-form figure (4160 flop/sample, "
                             "SURVEY 8d) counts flops the kernel does not execute; the binding roofline is HBM"},
        "timeshard_parity": bool(err <= 1e-5) if env.world > 1 else None,
        "parity": {"check": "+-2000 outputs around every seam recomputed by one cascade on one rank from zero state",
                   "max_err_over_rms": err, "tolerance": 1e-5, "samples_per_seam": int(cnt)},
    }
    if note:
        rec["note"] = note
    del x, y, tf
    return rec, per_run * reps


def leg_c5_captures(env, args, peak):
    """BASELINE configs[4]: 256 independent 1 s @ 10 Msps captures, offsetFreq -> bh151 -> bwLim(200 kHz)
    [D = 50] -> demod_fm, dealt over the ranks with shard.unit_range; each rank's captures in ONE launch."""
    torch = env.torch
    from directdemod_b200 import _lib, shard
    from directdemod_b200.fused import FusedChain
    first, last = shard.unit_range(C5_CAPTURES, env.world, env.rank)
    mine = last - first
    d = int(C5_FS / C5_BW)
    fit = (env.free_bytes() - (4 << 30)) // (C5_LEN * 9)
    note = None
    if mine > fit:
        mine = int(fit)
        note = "fewer captures than this rank's share fit its free memory"
    ch = FusedChain(taps_bh151(), d, 250000.0, C5_FS, demod=True, device=env.local)
    x = torch.empty((mine, C5_LEN), dtype=torch.complex64, device=env.dev)
    env.fill_noise(x, 100 + first)
    out = torch.empty((mine, ch.out_count(C5_LEN)), dtype=torch.float32, device=env.dev)
    l0 = _lib.launch_count()
    y = ch.apply_batch(x, out=out)
    per_run = _lib.launch_count() - l0
    reps = max(3, min(args.steps, 10))
    ms = env.timed(lambda: ch.apply_batch(x, out=out), reps)
    # self-check: one capture of the batch equals the same capture run alone as a fresh stream
    k = mine // 2
    ch1 = FusedChain(taps_bh151(), d, 250000.0, C5_FS, demod=True, device=env.local)
    alone = ch1.apply(x[k])
    same = bool(torch.equal(alone, y[k]))
    same = env.max_over_ranks([0.0 if same else 1.0])[0] == 0.0
    total = env.max_over_ranks([float(mine)])[0]                     # ranks hold 256/N each (+-1)
    bps = 8.0 + 4.0 / d
    gbs = mine * C5_LEN * bps / (ms * 1e-3) / 1e9
    rec = {
        "workload": "C5 (BASELINE configs[4]): %d captures x %d cf32 samples, D=%d, dealt over %d rank(s), no "
                    "inter-GPU traffic" % (C5_CAPTURES, C5_LEN, d, env.world),
        "scaling": "strong", "captures_per_rank": mine, "ms": round(ms, 4),
        "msps_total": round(C5_CAPTURES * C5_LEN / ms / 1e3, 1) if not note else round(total * env.world * C5_LEN / ms / 1e3, 1),
        "launches_per_run": per_run,
        "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s per GPU",
                     "frac": round(gbs / peak, 4), "algorithmic_bytes_per_sample": round(bps, 4)},
        "batch_equals_single_capture": same,
    }
    if note:
        rec["note"] = note
    del x, out, y
    return rec, per_run * reps


def run_ours(args):
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist
    from directdemod_b200 import _lib
    from directdemod_b200.fused import FusedChain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    # stdout carries exactly ONE line, the JSON record: whatever libraries print there (NCCL's version
    # banner) is sent to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    env = _Env(torch, dist, world, rank, local, dev)
    barrier = env.barrier

    n = args.samples
    taps = taps_bh151()
    # ---- synthetic capture, generated on the device (seeded per rank) ----
    x = torch.empty(n, dtype=torch.complex64, device=dev)
    env.fill_noise(x, 1234 + rank)
    chain = FusedChain(taps, DECIM, F_OFF, FS, demod=True, device=local)
    out = torch.empty(chain.out_count(n) + 1, dtype=torch.float32, device=dev)

    def one_pass():
        chain.set_position(0, 0, False)     # every step demodulates the capture from n0 = 0
        return chain.apply(x, out=out)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        y = one_pass()
    barrier()
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for e0, e1 in evs:
        chain.set_position(0, 0, False)
        e0.record()
        y = chain.apply(x, out=out)
        e1.record()
    t_end.record()
    barrier()
    launches = _lib.launch_count() - launches0
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    kern_avg_ms = sum(kern_ms) / len(kern_ms)
    checksum = float(y.double().sum().item())

    # ---- e2e: pinned host capture through ddm_chain_apply_host, 20 M-sample chunks ----
    # The same quarter pass per rank at every N (the leg is PCIe-bound: its rate does not depend on the
    # length, and 8 ranks x a full pass would pin most of the box's RAM).
    e2e = None
    e2e_ms = None
    n_e2e = min(n, args.e2e_samples)
    # several ranks stream from host memory at once: keep each rank's pinned staging buffers (first
    # touch) and its copy-issuing thread on the NUMA node of its GPU for the duration of this leg
    numa_note = None
    old_affinity = None
    if world > 1 and not args.no_numa_bind:
        near = _cpus_near_gpu(torch, local)
        if near is not None:
            try:
                old_affinity = os.sched_getaffinity(0)
                os.sched_setaffinity(0, near[1])
                numa_note = "rank bound to NUMA node %d (%d cpus) for the end-to-end leg" % (near[0], len(near[1]))
            except Exception:
                old_affinity = None
    e2e_steps = max(1, args.steps if args.e2e_steps <= 0 else min(args.steps, args.e2e_steps))
    try:
        host = torch.empty(n_e2e, dtype=torch.complex64, pin_memory=True)
        host.copy_(x[:n_e2e])
        host_np = host.numpy()
        out_host = torch.empty(n_e2e // DECIM + 2, dtype=torch.float32, pin_memory=True).numpy()
        h2d = d2h = 0

        def e2e_pass():
            nonlocal h2d, d2h
            chain.set_position(0, 0, False)
            h2d = d2h = 0
            pos = 0
            for a in range(0, n_e2e, CHUNK):
                b = min(n_e2e, a + CHUNK)
                got = chain.apply_host(host_np[a:b], out=out_host[pos:])
                pos += got.size
                h2d += (b - a) * 8
                d2h += got.size * 4
            return pos

        e2e_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            produced = e2e_pass()
        torch.cuda.synchronize()
        e2e_wall = time.perf_counter() - t0
        barrier()
        e2e_launches = e2e_steps * ((n_e2e + CHUNK - 1) // CHUNK)
        e2e_ms = e2e_wall / e2e_steps * 1e3
        e2e = {"samples_per_step": n_e2e, "steps": e2e_steps, "produced": int(produced),
               "h2d": int(h2d), "d2h": int(d2h)}
        del host
    except Exception as exc:  # pinned allocation can fail on a small host
        e2e = {"error": "%s: %s" % (type(exc).__name__, exc)}
        e2e_launches = 0
    # ---- the same two measurements fed with raw unsigned 8-bit I/Q, the format the reference's sources
    # actually read (source.py:117-118, :209-210); 2 B/sample instead of 8.  Every rank runs it. ----
    u8 = None
    u8_dev_ms = u8_host_ms = 0.0
    n_u8 = min(n, args.e2e_samples)
    if not args.no_u8:
        try:
            chain8 = FusedChain(taps, DECIM, F_OFF, FS, demod=True, device=local, in_format="cu8")
            xu = torch.empty((n_u8, 2), dtype=torch.uint8, device=dev)
            slab = 1 << 26
            for a in range(0, n_u8, slab):
                b = min(n_u8, a + slab)
                xu[a:b] = (torch.view_as_real(x[a:b]) + 127.5).clamp_(0, 255).to(torch.uint8)
            out8 = torch.empty(chain8.out_count(n_u8) + 2, dtype=torch.float32, device=dev)

            def pass8():
                chain8.set_position(0, 0, False)
                return chain8.apply(xu, out=out8)
            reps_dev8 = max(3, min(args.steps, 10))
            u8_dev_ms = env.timed(pass8, reps_dev8, warm=3)
            host8 = torch.empty((n_u8, 2), dtype=torch.uint8, pin_memory=True)
            host8.copy_(xu)
            h8 = host8.numpy()
            oh8 = torch.empty(n_u8 // DECIM + 2, dtype=torch.float32, pin_memory=True).numpy()

            def e2e8():
                chain8.set_position(0, 0, False)
                pos = 0
                for a in range(0, n_u8, CHUNK):
                    b = min(n_u8, a + CHUNK)
                    pos += chain8.apply_host(h8[a:b], out=oh8[pos:]).size
                return pos
            e2e8()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                produced8 = e2e8()
            torch.cuda.synchronize()
            u8_host_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
            barrier()
            u8 = {"produced": int(produced8)}
            del xu, host8
        except Exception as exc:
            u8 = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if old_affinity is not None:
        try:
            os.sched_setaffinity(0, old_affinity)
        except Exception:
            pass
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- reduce over ranks: slowest rank defines the time ----
    total_ms, kern_avg_ms, e2e_ms, u8_host_ms = env.max_over_ranks(
        [total_ms, kern_avg_ms, e2e_ms if e2e_ms else 0.0, u8_host_ms])

    # ---- sub-records: the time-sharded stream, the C4 cascade, the C5 captures ----
    peak, peak_src = _peaks()
    del x, out, y
    extra = {}
    extra_launches = 0
    if not args.no_extra:
        for name, leg in (("timeshard_chain", leg_timeshard_chain), ("c4_cascade", leg_c4_cascade),
                          ("c5_captures", leg_c5_captures)):
            if args.only_extra and name not in args.only_extra.split(","):
                continue
            try:
                torch.cuda.empty_cache()
                rec, cnt = leg(env, args, peak)
                extra[name] = rec
                extra_launches += cnt
            except Exception as exc:
                extra[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            torch.cuda.empty_cache()

    if rank == 0:
        value = world * n * args.steps / (total_ms * 1e-3) / 1e6
        achieved = n * ALG_BYTES_PER_SAMPLE / (kern_avg_ms * 1e-3) / 1e9
        traffic = _traffic(n)
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "mode": "device-resident, 1 fused launch per pass",
            "roofline": {
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": round(traffic["traffic_bytes_per_launch"] / 1e9, 3) if traffic else None,
                "traffic_unit": "GB per launch",
                "traffic_source": traffic.get("source") if traffic else
                "not measured in this run (ncu --set full capture of this launch size: profiles/)",
                "algorithmic_gb_per_launch": round(n * ALG_BYTES_PER_SAMPLE / 1e9, 3),
                "peak_source": peak_src, "kernel": "ddm::chain_stream_kernel<Q=5,MIX,FM>, 16 warps per SM, one TMA stage per warp (1 launch per step)",
                "note": "peak is the driver's copy (read+write) figure; a read-only stream reaches ~7350 GB/s on "
                        "this part (profiles/r01_microbench.txt), so a 99.6%-read kernel can exceed frac 1.0",
                "algorithmic_bytes_per_sample": round(ALG_BYTES_PER_SAMPLE, 4),
                "kernel_ms_avg": round(kern_avg_ms, 4), "kernel_ms_min": round(kern_ms[0], 4),
            },
            "clocks": sampler.summary(),
            "gpu_launches": int(launches),
            "gpu_launches_other_legs": int(e2e_launches + extra_launches),
            "checksum": checksum,
        }
        if e2e_ms:
            e2e_val = world * e2e["samples_per_step"] / (e2e_ms * 1e-3) / 1e6
            line["e2e"] = {"value": round(e2e_val, 1), "unit": UNIT,
                           "h2d_bytes_per_step": world * e2e["h2d"], "d2h_bytes_per_step": world * e2e["d2h"],
                           "ms_per_step": round(e2e_ms, 3), "steps": e2e["steps"],
                           "samples_per_step": e2e["samples_per_step"],
                           "path": "ddm_chain_apply_host per %d-sample chunk, pinned host buffers; the same "
                                   "%d samples per rank at every N" % (CHUNK, e2e["samples_per_step"])}
            if numa_note:
                line["e2e"]["numa"] = numa_note
        else:
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": str(e2e)}
        if u8 is not None:
            if "error" in u8:
                line["cu8_input"] = u8
            else:
                line["cu8_input"] = {
                    "samples_per_step": n_u8, "device_ms": round(u8_dev_ms, 4),
                    "device_resident_msps": round(world * n_u8 / u8_dev_ms / 1e3, 1),
                    "e2e": {"value": round(world * n_u8 / u8_host_ms / 1e3, 1), "unit": UNIT,
                            "ms_per_step": round(u8_host_ms, 3), "steps": e2e_steps,
                            "h2d_bytes_per_step": world * 2 * n_u8, "d2h_bytes_per_step": world * 4 * u8["produced"]},
                    "note": "same chain and capture quantised to interleaved unsigned 8-bit I/Q (what source.py reads "
                            "from a recording); all %d rank(s), slowest rank's time" % world}
        line.update(extra)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single(reps=args.cpu_reps)
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=N_PASS, help="samples per GPU per step")
    ap.add_argument("--e2e-samples", type=int, default=E2E_SAMPLES)
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-rank runs: do not bind ranks to their GPU's NUMA node for the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (0 = --steps)")
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-u8", action="store_true", help="skip the unsigned 8-bit input measurement")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the timeshard_chain / c4_cascade / c5_captures sub-records")
    ap.add_argument("--only-extra", default="", help="comma-separated subset of the sub-records")
    ap.add_argument("--ref-procs", type=int, default=64, help="max processes of the reference arm")
    ap.add_argument("--configs", action="store_true",
                    help="per-config table (BASELINE configs C1..C5, one JSON object per line) instead of the "
                         "headline line: scripts/bench_configs.py with the oracle CPU legs switched on")
    ap.add_argument("--quick", action="store_true", help="with --configs: reduced sizes")
    ap.add_argument("--only", default="", help="with --configs: comma-separated subset, e.g. c2,c4")
    args = ap.parse_args()
    if args.configs:
        # cpu_baseline leg of the per-config table: the one place besides the headline's
        # cpu_baseline / --impl reference where bench.py executes oracle/
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_configs
        from oracle import ddoracle
        bench_configs.set_cpu_oracle(ddoracle)
        return bench_configs.main((["--quick"] if args.quick else []) + (["--only", args.only] if args.only else []))
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
