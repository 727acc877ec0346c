#!/usr/bin/env python
"""(Round 1; bench.py now carries both measurements as the sub-records `timeshard_chain` and `c4_cascade`
of its JSON line, with seam self-checks.)
One long stream split in time across the GPUs of a box (torchrun, NCCL halo exchange):
(a) the fused chain on a C2-like stream, (b) the C4 cascade (1023-tap Remez + 8th-order
Butterworth).  Strong scaling: the stream length is fixed, every rank holds 1/world of it.
Device time per rank by CUDA events (halo exchange included), max over ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_timeshard.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import scipy.signal as sps
import torch
import torch.distributed as dist

from directdemod_b200 import filters, shard


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def timed(fn, reps=5):
        fn()
        best = None
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t) if best is None else min(best, float(t))
        return best

    # (a) fused chain: 4 x the C2 capture as ONE stream (7.37e9 samples), split in time
    n_total = 4 * 1843200000
    ts = shard.TimeShardedChain(sps.windows.blackmanharris(151), 34, 30000.0, 2048000, n_total, rank, world, device=local)
    n = ts.end - ts.start
    x = torch.empty(n, dtype=torch.complex64, device=dev)
    torch.view_as_real(x).normal_(0, 40)
    ms = timed(lambda: ts.run(x))
    if rank == 0:
        print(json.dumps({"workload": "fused chain, one stream of %d samples split in time" % n_total, "n_gpus": world,
                          "samples_per_rank": n, "ms": round(ms, 3), "msps_total": round(n_total / ms / 1e3, 1),
                          "halo_samples": ts.halo_len}), flush=True)
    del x, ts
    torch.cuda.empty_cache()

    # (b) C4: 1 h @ 2.4 Msps through remez-1023 + butter-8, split in time (on < 8 GPUs: the first
    # world/8 of the hour, so that a rank always holds the 1.08e9-sample slab of the 8-GPU split)
    fs = 2400000
    n_total = fs * 3600 // 8 * world
    fir = filters.remez(fs, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    iir = filters.butter(fs, 100000, n=8)
    tf = shard.TimeShardedFilters([fir, iir], n_total, rank, world)
    n = tf.end - tf.start
    x = torch.empty(n, dtype=torch.complex64, device=dev)
    torch.view_as_real(x).normal_(0, 40)
    ms = timed(lambda: tf.run(x), reps=3)
    if rank == 0:
        print(json.dumps({"workload": "C4 cascade remez1023 + butter8, stream of %d samples split in time" % n_total,
                          "n_gpus": world, "samples_per_rank": n, "ms": round(ms, 3),
                          "msps_total": round(n_total / ms / 1e3, 1), "halo_samples": tf.halo_len}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
