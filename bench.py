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
timed region.  `--impl reference` times the oracle port of the reference's scipy path
(oracle/ddoracle.py; the reference itself is Python and cannot travel to the GPU box) on
the host cores.
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
ALG_BYTES_PER_SAMPLE = 8.0 + 4.0 / DECIM   # cf32 in + f32 out every D samples
METRIC = "IQ Msps (shift->FIR->decim->FM demod)"
UNIT = "Msamples/s"


def _traffic():
    """DRAM bytes per launch of the fused kernel from the committed ncu --set full capture
    (profiles/r01_traffic.json); None if the capture is missing or was taken on another size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as fh:
            t = json.load(fh)
        return t if int(t["samples_per_launch"]) > 0 else None
    except Exception:
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


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: oracle port of the reference's scipy path
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


def _cpu_init(n):
    """Per-process setup: synthesise one chunk of input and a fresh chain state."""
    from oracle import ddoracle as O
    seed = os.getpid()
    _CPU["x"] = _cpu_chunk_input(seed, n)
    _CPU["taps"] = taps_bh151()
    _CPU["st"] = O.ChainState(_CPU["taps"])
    O.chain_chunk(_CPU["x"][:200000], FS, F_OFF, _CPU["taps"], BW, _CPU["st"])   # warm up


def _cpu_step(reps):
    """Process `reps` chunks like decode_noaa.__audio does; returns the seconds it took."""
    from oracle import ddoracle as O
    t0 = time.perf_counter()
    for _ in range(reps):
        O.chain_chunk(_CPU["x"], FS, F_OFF, _CPU["taps"], BW, _CPU["st"])
    return time.perf_counter() - t0


def cpu_baseline_single(reps=3, n=CHUNK):
    _cpu_init(n)
    sec = _cpu_step(reps)
    _CPU.clear()
    return {"value": round(reps * n / sec / 1e6, 3), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d chunks x %d samples of the same chain (oracle/ddoracle.chain_chunk: numpy mixer + "
                      "scipy.signal.lfilter + stride decimation + np.angle, float64 like the reference), "
                      "1 process" % (reps, n)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on all usable host cores, one
    independent stream per process (the reference itself is single threaded per stream).
    A step = every process pushes one PROC_CHUNKSIZE chunk through the chain; the step time is
    the slowest process (input synthesis is outside the timed section)."""
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
    per_worker = n * (8 + 16 * 4 + 8)                  # c64 input + complex128 temporaries
    workers = max(1, min(cores, args.ref_procs, int(avail * 0.25 // per_worker)))
    ctx = mp.get_context("fork")
    total = 0.0
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(n,)) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_step, [1] * workers, chunksize=1)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_step, [1] * workers, chunksize=1)
            total += time.perf_counter() - t0
    value = args.steps * workers * n / total / 1e6
    sample = ("%d steps x %d processes x one %d-sample chunk each (independent streams, "
              "host has %d logical cores)" % (args.steps, workers, n, cores))
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(total / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1, "reference-cpu"),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus, mode):
    return {
        "workload": "NOAA APT pass (BASELINE configs[1]): %d s @ %d sps = %d cf32 samples per GPU; "
                    "offsetFreq(%g) -> blackmanHarris(%d) -> bwLim(%d) [D=%d] -> demod_fm"
                    % (PASS_SECONDS, FS, N_PASS, F_OFF, NTAPS, BW, DECIM),
        "samples_per_gpu": N_PASS, "decim": DECIM, "ntaps": NTAPS, "mode": mode,
        "l2": "input per step is 14.7 GB (>> 126 MB L2); no flush needed",
        "sharding": "independent capture per GPU, no data-path collective" if n_gpus > 1 else "single GPU",
    }


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


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from directdemod_b200 import _lib
    from directdemod_b200.fused import FusedChain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.samples
    taps = taps_bh151()
    # ---- synthetic capture, generated on the device (seeded per rank) ----
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    x = torch.empty(n, dtype=torch.complex64, device=dev)
    xr = torch.view_as_real(x)
    slab = 1 << 26
    for a in range(0, n, slab):
        b = min(n, a + slab)
        xr[a:b].normal_(0.0, 40.0, generator=gen)
    chain = FusedChain(taps, DECIM, F_OFF, FS, demod=True, device=local)
    out = torch.empty(chain.out_count(n) + 1, dtype=torch.float32, device=dev)

    def one_pass():
        chain.set_position(0, 0, False)     # every step demodulates the capture from n0 = 0
        return chain.apply(x, out=out)

    for _ in range(max(args.warmup, 3)):
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
    e2e = None
    e2e_ms = None
    # with several ranks on one host the pinned staging buffers add up (8 x 14.7 GB would pin most of
    # the box's RAM): the end-to-end leg then streams a quarter pass per rank -- it is PCIe-bound, its
    # rate does not depend on the length
    if world > 1 and args.e2e_samples == N_PASS:
        args.e2e_samples = N_PASS // 4
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
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            produced = e2e_pass()
        torch.cuda.synchronize()
        e2e_wall = time.perf_counter() - t0
        barrier()
        e2e_ms = e2e_wall / e2e_steps * 1e3
        e2e = {"samples_per_step": n_e2e, "steps": e2e_steps, "produced": int(produced),
               "h2d": int(h2d), "d2h": int(d2h)}
    except Exception as exc:  # pinned allocation can fail on a small host
        e2e = {"error": "%s: %s" % (type(exc).__name__, exc)}
    finally:
        if old_affinity is not None:
            try:
                os.sched_setaffinity(0, old_affinity)
            except Exception:
                pass
    # ---- extra (not the headline): the same pass fed with raw unsigned 8-bit I/Q, the format the
    # reference's sources actually read (source.py:117-118); 2 B/sample instead of 8 ----
    u8 = None
    if not args.no_u8:
        try:
            n_u8 = min(n, args.e2e_samples)
            chain8 = FusedChain(taps, DECIM, F_OFF, FS, demod=True, device=local, in_format="cu8")
            xu = torch.empty((n_u8, 2), dtype=torch.uint8, device=dev)
            for a in range(0, n_u8, slab):
                b = min(n_u8, a + slab)
                xu[a:b] = (torch.view_as_real(x[a:b]) + 127.5).clamp_(0, 255).to(torch.uint8)
            out8 = torch.empty(chain8.out_count(n_u8) + 2, dtype=torch.float32, device=dev)

            def pass8():
                chain8.set_position(0, 0, False)
                return chain8.apply(xu, out=out8)
            for _ in range(3):
                pass8()
            torch.cuda.synchronize()
            ts = []
            for _ in range(max(3, min(args.steps, 10))):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pass8()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            dev_ms = sum(ts) / len(ts)
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
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps8 = max(1, min(args.steps, args.e2e_steps))
            for _ in range(reps8):
                e2e8()
            torch.cuda.synchronize()
            host_ms = (time.perf_counter() - t0) / reps8 * 1e3
            u8 = {"samples_per_step": n_u8, "device_resident_msps": round(n_u8 / dev_ms / 1e3, 1),
                  "device_ms": round(dev_ms, 4), "e2e_msps": round(n_u8 / host_ms / 1e3, 1),
                  "e2e_ms": round(host_ms, 3), "h2d_bytes_per_step": 2 * n_u8,
                  "note": "same chain and capture quantised to unsigned 8-bit I/Q; per GPU, rank 0"}
            del xu, host8
        except Exception as exc:
            u8 = {"error": "%s: %s" % (type(exc).__name__, exc)}
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- reduce over ranks: slowest rank defines the time ----
    vals = torch.tensor([total_ms, kern_avg_ms, e2e_ms if e2e_ms else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    total_ms, kern_avg_ms, e2e_ms = [float(v) for v in vals.tolist()]

    if rank == 0:
        peak, peak_src = _peaks()
        value = world * n * args.steps / (total_ms * 1e-3) / 1e6
        achieved = n * ALG_BYTES_PER_SAMPLE / (kern_avg_ms * 1e-3) / 1e9
        traffic = _traffic()
        traffic_bytes = None
        if traffic is not None:
            # per launch like `achieved`: scaled by samples if this run uses another capture size
            traffic_bytes = traffic["traffic_bytes_per_launch"] * (n / float(traffic["samples_per_launch"]))
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, "device-resident, 1 fused launch per pass"),
            "roofline": {
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": round(traffic_bytes / 1e9, 3) if traffic_bytes else None, "traffic_unit": "GB per launch",
                "algorithmic_gb_per_launch": round(n * ALG_BYTES_PER_SAMPLE / 1e9, 3),
                "peak_source": peak_src, "kernel": "ddm::chain_fused_kernel<Q=5,MIX,FM> (1 launch per step)",
                "note": "peak is the driver's copy (read+write) figure; a read-only stream reaches ~7350 GB/s on "
                        "this part (profiles/r01_microbench.txt), so a 99.6%-read kernel can exceed frac 1.0",
                "algorithmic_bytes_per_sample": round(ALG_BYTES_PER_SAMPLE, 4),
                "kernel_ms_avg": round(kern_avg_ms, 4), "kernel_ms_min": round(kern_ms[0], 4),
            },
            "clocks": sampler.summary(),
            "gpu_launches": int(launches),
            "checksum": checksum,
        }
        if e2e_ms:
            e2e_val = world * e2e["samples_per_step"] / (e2e_ms * 1e-3) / 1e6
            line["e2e"] = {"value": round(e2e_val, 1), "unit": UNIT,
                           "h2d_bytes_per_step": world * e2e["h2d"], "d2h_bytes_per_step": world * e2e["d2h"],
                           "ms_per_step": round(e2e_ms, 3), "steps": e2e["steps"],
                           "samples_per_step": e2e["samples_per_step"],
                           "path": "ddm_chain_apply_host per %d-sample chunk, pinned host buffers" % CHUNK}
            if numa_note:
                line["e2e"]["numa"] = numa_note
        else:
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "note": str(e2e)}
        if u8 is not None:
            line["cu8_input"] = u8
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single(reps=args.cpu_reps)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=N_PASS, help="samples per GPU per step")
    ap.add_argument("--e2e-samples", type=int, default=N_PASS)
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="multi-rank runs: do not bind ranks to their GPU's NUMA node for the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-reps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-u8", action="store_true", help="skip the extra unsigned 8-bit input measurement")
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
