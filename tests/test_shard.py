"""Multi-process sharding logic on CPU (gloo, world_size 2 and 3): slab bounds, the neighbour
halo exchange, and -- with the oracle standing in for the kernel -- that per-rank results of a
time-sharded stream concatenate to exactly the single-process result.  The same flow on real
GPUs over NCCL is tests/test_shard_gpu.py."""

import os
import socket

import numpy as np
import pytest

from oracle import ddoracle as O
from tests.util import fm_tone_c64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_unit_range_and_slab_bounds():
    from directdemod_b200 import shard
    for n, w in ((256, 8), (10, 3), (3, 8), (0, 2)):
        got = [shard.unit_range(n, w, r) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        assert max(e - s for s, e in got) - min(e - s for s, e in got) <= 1
    b = shard.slab_bounds(8640000000, 8, 34)
    assert b[0][0] == 0 and b[-1][1] == 8640000000
    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
    assert all(s % 34 == 0 for s, _ in b)
    assert shard.decim_offset_at(b[3][0], 34) == 0
    assert shard.decim_offset_at(35, 34) == 33 and shard.decim_offset_at(0, 34) == 0
    with pytest.raises(ValueError):
        shard.unit_range(4, 2, 2)


def _worker(rank, world, port, n, out_dir):
    import torch
    import torch.distributed as dist
    from directdemod_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fs, f, decim, ntaps = 2048000, 30000.0, 34, 151
    taps = O.taps_blackman_harris(ntaps)[0]
    x = fm_tone_c64(3, n, fs, f, 1300.0, 2.0)
    bounds = shard.slab_bounds(n, world, decim)
    start, end = bounds[rank]
    # halo length of the fused chain (csrc/chain.cu): (ceil((K+1)/D) + 1) * D, made even
    H = ((ntaps + 1 + decim - 1) // decim + 1) * decim
    H += H & 1
    slab = torch.from_numpy(x[start:end].copy())
    tail = slab[-H:] if rank + 1 < world else torch.empty(H, dtype=slab.dtype)
    halo = shard.exchange_halo(tail, rank, world)
    off = shard.decim_offset_at(start, decim)
    if rank == 0:
        assert halo is None
        st = O.ChainState(taps)
        y, _ = O.chain_chunk(x[start:end], fs, f, taps, fs / decim, st)
    else:
        assert np.array_equal(halo.numpy(), x[start - H:start])
        # what the kernel does with a halo: recompute the FIR state and the previous decimated
        # sample from raw history, mixer phase from the global index
        ext = np.concatenate([halo.numpy(), x[start:end]])
        mixed, _ = O.mix(ext, f, fs, start - H)
        yy = O.filt_stateless(taps, [1], mixed.astype(np.complex128))
        dec = yy[H + off - decim::decim]
        y, _ = O.fm_discriminator(dec, None, store_state=False)
    np.save(os.path.join(out_dir, "part%d.npy" % rank), y)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_time_sharded_stream_equals_single_process(tmp_path, world):
    import torch.multiprocessing as mp
    n = 90000
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / ("part%d.npy" % r)) for r in range(world)]
    got = np.concatenate(parts)
    fs, f, decim = 2048000, 30000.0, 34
    x = fm_tone_c64(3, n, fs, f, 1300.0, 2.0)
    want, _ = O.chain_stream(x, fs, f, O.taps_blackman_harris(151)[0], fs / decim)
    assert got.shape == want.shape
    assert np.max(np.abs(np.angle(np.exp(1j * (got - want))))) < 1e-9


def _worker_filters(rank, world, port, n, out_dir):
    import scipy.signal as sps
    import torch
    import torch.distributed as dist
    from directdemod_b200 import filters, shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(9)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    fir = filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=255)
    iir = filters.butter(2400000, 100000, n=8)
    H = fir.lookback() + iir.lookback()                 # host-only analysis, no device needed
    bounds = shard.slab_bounds(n, world, 1)
    start, end = bounds[rank]
    slab = torch.from_numpy(x[start:end].copy())
    tail = slab[-H:] if rank + 1 < world else torch.empty(H, dtype=slab.dtype)
    halo = shard.exchange_halo(tail, rank, world)
    b1, a1 = np.asarray(fir.getB, dtype=np.float64), [1.0]
    b2, a2 = iir.getB, iir.getA
    if rank == 0:
        y, _ = sps.lfilter(b1, a1, x[start:end].astype(np.complex128), zi=sps.lfilter_zi(b1, a1))
        y, _ = sps.lfilter(b2, a2, y, zi=sps.lfilter_zi(b2, a2))
    else:
        ext = np.concatenate([halo.numpy(), x[start:end]]).astype(np.complex128)
        y = sps.lfilter(b2, a2, sps.lfilter(b1, a1, ext))[H:]        # zero state, drop the halo
    np.save(os.path.join(out_dir, "fpart%d.npy" % rank), y)
    dist.barrier()
    dist.destroy_process_group()


def test_time_sharded_fir_iir_cascade_equals_single_process(tmp_path):
    """C4-shaped cascade (Remez FIR -> 8th-order Butterworth): one raw-input halo of
    (ntaps-1) + W samples per slab boundary reproduces the single-process stream."""
    import scipy.signal as sps
    import torch.multiprocessing as mp
    from directdemod_b200 import constants, filters
    n, world = 60000, 3
    mp.spawn(_worker_filters, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / ("fpart%d.npy" % r)) for r in range(world)])
    rng = np.random.default_rng(9)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
    b1, a1 = O.taps_remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], 255)
    b2, a2 = O.taps_butter(2400000, 100000, n=8)
    want, _ = sps.lfilter(b1, a1, x.astype(np.complex128), zi=sps.lfilter_zi(b1, a1))
    want, _ = sps.lfilter(b2, a2, want, zi=sps.lfilter_zi(b2, a2))
    assert got.shape == want.shape
    assert O.rel_rms(got, want) <= 1e-9
    # the NOAA band-pass only runs as a sequential replay: it must refuse to be time-sharded
    with pytest.raises(ValueError):
        filters.butter(60235, 400, 4400, n=6, typeFlt=constants.FLT_BP).lookback()
