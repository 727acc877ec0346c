"""Time-sharded fused chain on real GPUs over NCCL (needs >= 2 devices; the driver's 1-GPU run
skips it).  Per-rank outputs must concatenate to the single-GPU result (same kernel, same inputs,
same global positions: equal to a float32 ulp) and match the oracle within the stated tolerance."""

import os
import socket

import numpy as np
import pytest

from oracle import ddoracle as O
from tests.util import TOL, fm_tone_c64, wrap_rel_rms

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    from directdemod_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    fs, f, decim = 2048000, 30000.0, 34
    taps = O.taps_blackman_harris(151)[0]
    x = fm_tone_c64(3, n, fs, f, 1300.0, 2.0)
    ts = shard.TimeShardedChain(taps, decim, f, fs, n, rank, world, device=rank)
    slab = torch.from_numpy(x[ts.start:ts.end].copy()).cuda()
    y = ts.run(slab)
    chk = ts.boundary_check(slab, y, width=300)          # the self-check bench.py reports
    assert chk["max_abs_err"] <= 1e-6 and chk["samples"] == (600 if rank else 0), chk
    np.save(os.path.join(out_dir, "part%d.npy" % rank), y.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_chain_over_nccl(tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    from directdemod_b200.fused import FusedChain
    n = 6000000
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("part%d.npy" % r)) for r in range(world)])
    fs, f, decim = 2048000, 30000.0, 34
    taps = O.taps_blackman_harris(151)[0]
    x = fm_tone_c64(3, n, fs, f, 1300.0, 2.0)
    single = FusedChain(taps, decim, f, fs).apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert got.shape == single.shape
    # same kernel, same samples, same global positions; only the float64 block rotators are advanced
    # from different anchor tiles, which may move a result by one float32 ulp
    assert np.max(np.abs(np.angle(np.exp(1j * (got.astype(np.float64) - single))))) <= 1e-6
    want, _ = O.chain_stream(x, fs, f, taps, fs / decim)
    assert wrap_rel_rms(got, want) <= TOL


def _worker_filters(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    from directdemod_b200 import filters, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(9)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    fir = filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
    iir = filters.butter(2400000, 100000, n=8)
    ts = shard.TimeShardedFilters([fir, iir], n, rank, world)
    slab = torch.from_numpy(x[ts.start:ts.end].copy()).cuda()
    y = ts.run(slab)
    b1, b2, a2 = np.asarray(fir.getB), np.asarray(iir.getB), np.asarray(iir.getA)
    chk = ts.boundary_check(slab, y, lambda: [filters.filter(b1, [1]), filters.filter(b2, a2)], width=500)
    assert chk["max_rel_err"] <= 1e-5 and chk["samples"] == (1000 if rank else 0), chk
    np.save(os.path.join(out_dir, "fpart%d.npy" % rank), y.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_c4_cascade_over_nccl(tmp_path):
    """BASELINE config 4 in miniature: 1023-tap Remez + 8th-order Butterworth on a stream split in
    time across GPUs, halo moved by NCCL; equals the oracle's whole-stream result."""
    import scipy.signal as sps
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp
    n = 4000000
    mp.spawn(_worker_filters, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("fpart%d.npy" % r)) for r in range(world)])
    rng = np.random.default_rng(9)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    b1, a1 = O.taps_remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], 1023)
    b2, a2 = O.taps_butter(2400000, 100000, n=8)
    want, _ = sps.lfilter(b1, a1, x.astype(np.complex128), zi=sps.lfilter_zi(b1, a1))
    want, _ = sps.lfilter(b2, a2, want, zi=sps.lfilter_zi(b2, a2))
    assert got.shape == want.shape
    assert O.rel_rms(got, want) <= TOL


def test_time_sharded_c4_cascade_slabs_on_one_gpu():
    """The same cascade with the slabs of a 3-way split run one after the other on ONE device, the
    halo handed over directly: what every rank computes, without needing several GPUs."""
    import scipy.signal as sps
    import torch
    from directdemod_b200 import filters, shard
    n, world = 1500000, 3
    rng = np.random.default_rng(10)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    parts = []
    for rank in range(world):
        fir = filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=1023)
        iir = filters.butter(2400000, 100000, n=8)
        ts = shard.TimeShardedFilters([fir, iir], n, rank, world)
        halo = xd[ts.start - ts.halo_len:ts.start] if rank else xd[:0]
        assert ts.halo_len >= 1022
        parts.append(ts.run(xd[ts.start:ts.end], halo=halo).cpu().numpy())
    got = np.concatenate(parts)
    b1, a1 = O.taps_remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], 1023)
    b2, a2 = O.taps_butter(2400000, 100000, n=8)
    want, _ = sps.lfilter(b1, a1, x.astype(np.complex128), zi=sps.lfilter_zi(b1, a1))
    want, _ = sps.lfilter(b2, a2, want, zi=sps.lfilter_zi(b2, a2))
    assert got.shape == want.shape
    assert O.rel_rms(got, want) <= TOL
    for a in np.cumsum([len(p) for p in parts])[:-1]:                              # no seam at the slab joins
        assert O.rel_rms(got[a - 500:a + 500], want[a - 500:a + 500]) <= TOL


def test_time_sharded_chain_body_then_head_on_one_gpu():
    """What a rank > 0 of a time-sharded stream launches -- the slab's body first (no neighbour data
    needed), its head once the halo has arrived -- run for every rank of a 3-way split on ONE device with
    the halo handed over directly: the parts concatenate to the single-launch result, also for a slab
    too short to be split and for an odd decimation factor."""
    import torch
    from directdemod_b200 import shard
    from directdemod_b200.fused import FusedChain
    for decim, fs, n in ((34, 2048000, 9000000), (33, 2048000, 7000003), (34, 2048000, 2500000)):
        f, world = 30000.0, 3
        taps = O.taps_blackman_harris(151)[0]
        x = fm_tone_c64(31, n, fs, f, 1300.0, 2.0)
        xd = torch.from_numpy(x).cuda()
        single = FusedChain(taps, decim, f, fs).apply(xd).cpu().numpy()
        parts = []
        for rank in range(world):
            ts = shard.TimeShardedChain(taps, decim, f, fs, n, rank, world, device=0)
            slab = xd[ts.start:ts.end]
            if rank == 0:
                ts.chain.set_position(0, 0, False)
                parts.append(ts.chain.apply(slab).cpu().numpy())
                continue
            out, head, m_head = ts._launch_body(slab)
            assert (head < slab.numel()) == (slab.numel() > (1 << 20) + 2 * ts.halo_len + 2 * decim)
            y = ts._launch_head(slab, xd[ts.start - ts.halo_len:ts.start].contiguous(), out, head, m_head)
            parts.append(y.cpu().numpy())
        got = np.concatenate(parts)
        assert got.shape == single.shape, (decim, n)
        assert np.max(np.abs(np.angle(np.exp(1j * (got.astype(np.float64) - single))))) <= 1e-6, (decim, n)
