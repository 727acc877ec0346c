"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed).

Two seams, both taken from the reference's own structure:

* independent units -- separate captures, separate channels of one recording, the
  accurate-sync windows (decode_noaa.py:844,865): ``unit_range`` deals them out, no
  communication.
* one long stream in time -- the chunker's seam (chunker.py): rank g owns a contiguous slab
  of the stream.  The only data that crosses a slab boundary is what the reference carries
  from chunk to chunk, expressed as raw input history: the last ``halo_len`` samples of the
  previous slab (FIR delay line + previous decimated sample for the FM discriminator; for the
  segment-parallel IIR the same halo is its warm-up).  ``exchange_halo`` moves it with ONE
  neighbour send/recv per rank (NCCL point-to-point over NVLink; a few KB, latency bound:
  45-63 us measured), which ``TimeShardedChain.run`` overlaps with the launch over the slab's
  body.  The mixer phase and the decimation phase need no exchange: both are functions of the
  global sample index.  ``boundary_check`` recomputes the outputs around a seam on one rank
  alone -- the self-check bench.py reports at every N.

The sequential (bit-exact replay) IIR mode cannot be time-sharded -- its state is a serial
dependency by construction; such filters shard by independent units only.
"""

from __future__ import annotations


def unit_range(n_units, world, rank):
    """[first, last) of the units rank ``rank`` of ``world`` processes (contiguous blocks,
    sizes differing by at most one)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    base, extra = divmod(int(n_units), world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def slab_bounds(n_samples, world, decim=1):
    """Time slabs [start, end) per rank, starts aligned to the decimation factor so every
    slab begins on a kept sample (comm.py:124: kept samples are those with global index
    == offset (mod j); offset is 0 at the start of a stream)."""
    n_samples, decim = int(n_samples), int(decim)
    if world < 1 or decim < 1:
        raise ValueError("bad world/decimation")
    starts = [(g * n_samples // world) // decim * decim for g in range(world)]
    ends = starts[1:] + [n_samples]
    return [(s, e) for s, e in zip(starts, ends)]


def decim_offset_at(start, decim, stream_offset=0):
    """Chunker variable "bwlim" a chunk starting at global index ``start`` would see."""
    return (stream_offset - start) % decim


def exchange_halo(tail, rank, world, group=None):
    """Ring shift g -> g+1 of the slab tails.  ``tail``: this rank's last halo_len input
    samples (tensor on the device of the process group's backend; every rank passes the same
    shape and dtype).  Returns the tensor received from rank-1, or None on rank 0 (which starts
    from the reference's initial condition)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return None
    ops = []
    recv = torch.empty_like(tail) if rank > 0 else None
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, tail.contiguous(), rank + 1, group))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, recv, rank - 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return recv


class TimeShardedChain:
    """The fused chain over one slab of a time-sharded stream.

    Every rank builds the same chain (taps, decimation, mixer); ``run`` receives the slab's
    raw cf32 samples (cuda tensor), exchanges halos with the neighbours and returns this rank's
    part of the output.  Concatenating the parts of ranks 0..world-1 gives exactly what one
    GPU (and the reference) produces for the whole stream."""

    def __init__(self, taps, decim, freq_offset, samp_rate, n_samples, rank, world, demod=True,
                 device=None, group=None):
        from .fused import FusedChain
        self.rank, self.world, self.group = int(rank), int(world), group
        self.decim = int(decim)
        self.bounds = slab_bounds(n_samples, world, decim)
        self.start, self.end = self.bounds[rank]
        self.chain = FusedChain(taps, decim, freq_offset, samp_rate, demod=demod, device=device)
        self.halo_len = self.chain.halo_len
        for s, e in self.bounds[:-1]:
            if e - s < self.halo_len:
                raise ValueError("slab of %d samples is shorter than the %d-sample halo" % (e - s, self.halo_len))

    def run(self, x_slab):
        """The halo exchange is posted first and awaited last: the BODY of the slab (everything behind
        its first ``head`` samples) needs no neighbour data -- its history is the end of the head -- so
        its launch is queued while the NCCL send/recv is in flight, and only the short head launch
        waits for the halo.  Same samples, same global positions, same results as one launch."""
        import torch.distributed as dist
        n = self.end - self.start
        if x_slab.numel() != n:
            raise ValueError("slab has %d samples, expected %d" % (x_slab.numel(), n))
        if self.rank + 1 < self.world and n < self.halo_len:
            raise ValueError("slab shorter than the halo")
        off = decim_offset_at(self.start, self.decim)
        ch = self.chain
        if self.rank == 0 or self.world == 1:
            reqs = []
            if self.rank + 1 < self.world:
                reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, x_slab[-self.halo_len:].contiguous(),
                                                          self.rank + 1, self.group)])
            ch.set_position(0, off, False)
            y = ch.apply(x_slab)
            for r in reqs:
                r.wait()
            return y
        ops = [dist.P2POp(dist.irecv, self._recv_buf(x_slab), self.rank - 1, self.group)]
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, x_slab[-self.halo_len:].contiguous(), self.rank + 1, self.group))
        reqs = dist.batch_isend_irecv(ops)
        out, head, m_head = self._launch_body(x_slab)
        for r in reqs:
            r.wait()
        return self._launch_head(x_slab, self._recv, out, head, m_head)

    def _launch_body(self, x_slab):
        """Everything behind the slab's first ``head`` samples: needs no neighbour data (its history is
        the end of the head).  Returns (output tensor with the body part filled in, head, head outputs)."""
        import torch
        ch = self.chain
        n = self.end - self.start
        off = decim_offset_at(self.start, self.decim)
        # head: a whole number of decimation periods, at least the halo, ~1 M samples; the body starts
        # on a kept sample like the slab itself
        step = 2 * self.decim                                      # even: the body stays 16-byte aligned
        head = (max(self.halo_len, 1 << 20) + step - 1) // step * step
        if head >= n - self.halo_len:
            head = n                                               # short slab: one launch after the wait
        m_head = ch.count_for(head, off, True)
        off_body = decim_offset_at(self.start + head, self.decim)
        m_body = ch.count_for(n - head, off_body, True) if head < n else 0
        dt = torch.float32 if ch.demod else torch.complex64
        out = torch.empty(m_head + m_body, dtype=dt, device=x_slab.device)
        if head < n:
            ch.set_position(self.start + head, off_body, True, x_slab[head - self.halo_len:head])
            ch.apply(x_slab[head:], out=out[m_head:])
        return out, head, m_head

    def _launch_head(self, x_slab, halo, out, head, m_head):
        """The slab's first ``head`` samples, behind the halo received from the previous rank."""
        ch = self.chain
        ch.set_position(self.start, decim_offset_at(self.start, self.decim), True, halo)
        ch.apply(x_slab[:head], out=out[:m_head])
        return out

    def _recv_buf(self, like):
        import torch
        if getattr(self, "_recv", None) is None or self._recv.device != like.device:
            self._recv = torch.empty(self.halo_len, dtype=like.dtype, device=like.device)
        return self._recv

    def boundary_check(self, x_slab, y_slab, width=2000):
        """Self-check of the seam between rank-1 and this rank: the ``width`` outputs either side of
        the slab start are recomputed by ONE chain on this rank alone -- from ``width*decim`` (+ halo)
        raw samples of the previous slab, fetched with a second ring shift, and the head of this slab
        -- and compared with what the two ranks produced (previous rank's output tail ++ this rank's
        output head).  Returns {"max_abs_err", "bit_equal", "samples"} (rank 0: nothing to compare,
        zeros).  Collective: every rank must call it."""
        import torch
        from .fused import FusedChain
        w_in = width * self.decim
        need = w_in + self.halo_len
        for s, e in self.bounds:
            if e - s < need + self.decim:
                raise ValueError("slabs are too short for a %d-output boundary check" % width)
        tail_in = exchange_halo(x_slab[-need:].contiguous(), self.rank, self.world, self.group)
        tail_out = exchange_halo(y_slab[-width:].contiguous(), self.rank, self.world, self.group)
        if self.rank == 0:
            return {"max_abs_err": 0.0, "bit_equal": True, "samples": 0}
        c = self.chain
        chk = FusedChain(c._taps, self.decim, c.freq_offset, c.samp_rate, demod=c.demod, device=c.device)
        g0 = self.start - w_in
        chk.set_position(g0, decim_offset_at(g0, self.decim), True, tail_in[:self.halo_len].contiguous())
        win = torch.cat([tail_in[self.halo_len:], x_slab[:w_in]])
        got = chk.apply(win)
        want = torch.cat([tail_out, y_slab[:width]])
        if got.numel() != want.numel():
            raise RuntimeError("boundary window produced %d samples, expected %d" % (got.numel(), want.numel()))
        if c.demod:                    # phases: compare modulo 2 pi
            d = got.double() - want.double()
            err = torch.atan2(torch.sin(d), torch.cos(d)).abs().max()
        else:
            err = (got - want).abs().max()
        return {"max_abs_err": float(err), "bit_equal": bool(torch.equal(got, want)), "samples": int(got.numel())}


class TimeShardedFilters:
    """A cascade of filters.filter objects (BASELINE config 4: 1023-tap Remez, then an 8th-order
    Butterworth) over one slab of a time-sharded stream.

    Each filter needs ``filter.lookback()`` samples of history: ntaps-1 inputs for a FIR, the
    warm-up length W for the segment-parallel IIR.  The histories add up along the cascade, and
    all of it is RAW input of the previous slab, so ONE neighbour exchange of
    ``halo_len = sum(lookbacks)`` samples (NCCL point-to-point) serves the whole cascade and does
    not wait for the neighbour's compute.  Rank 0 runs the filters statefully from the reference's
    initial conditions (lfilter_zi); ranks > 0 run them from zero state over [halo ++ slab] -- every
    stage statefully over the halo first, then over the slab, so the concatenation is never
    materialised -- and keep the slab part: by then the zero-input response of every stage has
    decayed below one ulp.  A filter that only runs as a sequential replay raises at construction."""

    def __init__(self, filts, n_samples, rank, world, group=None, fuse=True):
        """``fuse``: run the cascade as ONE equivalent filter (filters.cascade: the stages' combined
        impulse response, dead to 1e-9 after a finite number of taps, applied by overlap-save FFT in a
        single pass) when the stages allow it; the halo is then that filter's taps - 1."""
        self.filts = list(filts)
        self.fused = False
        if fuse and len(self.filts) > 1:
            from . import filters as _filters
            try:
                self.filts = [_filters.cascade(self.filts)]
                self.fused = True
            except ValueError:
                pass
        self.rank, self.world, self.group = int(rank), int(world), group
        self.halo_len = int(sum(f.lookback() for f in self.filts)) if world > 1 else 0
        self.bounds = slab_bounds(n_samples, world, 1)
        self.start, self.end = self.bounds[rank]
        for s, e in self.bounds[:-1]:
            if world > 1 and e - s < self.halo_len:
                raise ValueError("slab of %d samples is shorter than the %d-sample halo" % (e - s, self.halo_len))

    def run(self, x_slab, halo=None):
        """Filter this rank's slab.  ``halo`` (the last ``halo_len`` raw samples of the previous
        slab, on the device) may be passed by a caller that moved it itself; otherwise it is
        exchanged here over the process group."""
        import numpy as np
        import torch
        from . import _dev, _lib
        if x_slab.numel() != self.end - self.start:
            raise ValueError("slab has %d samples, expected %d" % (x_slab.numel(), self.end - self.start))
        if self.world > 1 and halo is None:
            send = x_slab[-self.halo_len:] if self.rank + 1 < self.world else \
                torch.empty(self.halo_len, dtype=x_slab.dtype, device=x_slab.device)
            halo = exchange_halo(send.contiguous(), self.rank, self.world, self.group)
        if self.rank == 0 or self.world == 1:
            y = x_slab
            for f in self.filts:
                y = f._apply_dev(y)
            return y
        if halo.numel() != self.halo_len:
            raise ValueError("halo has %d samples, expected %d" % (halo.numel(), self.halo_len))
        # [halo ++ slab] from zero state without materialising the concatenation: every stage runs
        # statefully over the halo first (outputs feed the next stage's warm-up and are then dropped)
        # and carries its delay line / recursion state into the slab
        l = _lib.lib()
        st = _dev.stream_ptr(x_slab.device.index)
        y_h, y_s = halo.contiguous(), x_slab
        for f in self.filts:
            f._unshare()                      # the carried state below must be this cascade's own
            f.setState(np.zeros(f._state_len(), dtype=np.complex128))
            outs = []
            for part in (y_h, y_s):
                out = torch.empty_like(part)
                _lib.check(l.ddm_filter_apply_dev(f._handle(), _dev.ptr(part), part.numel(), int(part.is_complex()),
                                                  _dev.ptr(out), 1, st), "ddm_filter_apply_dev")
                outs.append(out)
            y_h, y_s = outs
        return y_s

    def boundary_check(self, x_slab, y_slab, make_filters, width=2000):
        """Self-check of the seam between rank-1 and this rank: a fresh copy of the cascade
        (``make_filters()``) runs on this rank alone from zero state over
        [halo_len + width raw samples of the previous slab ++ width samples of this slab] and its last
        2*width outputs are compared with (previous rank's output tail ++ this rank's output head).
        Returns {"max_rel_err", "samples"}: maximum absolute difference over the RMS of the window
        (the segment-parallel IIR cuts its segments differently in the two runs, so this is equality
        to the recursion's roundoff, not bit equality).  Collective: every rank must call it."""
        import torch
        need = self.halo_len + width
        for s, e in self.bounds:
            if e - s < need:
                raise ValueError("slabs are too short for a %d-sample boundary check" % width)
        tail_in = exchange_halo(x_slab[-need:].contiguous(), self.rank, self.world, self.group)
        tail_out = exchange_halo(y_slab[-width:].contiguous(), self.rank, self.world, self.group)
        if self.rank == 0 or self.world == 1:
            return {"max_rel_err": 0.0, "samples": 0}
        chk = TimeShardedFilters(make_filters(), self.bounds[-1][1], self.rank, self.world, self.group)
        win = torch.cat([tail_in[self.halo_len:], x_slab[:width]])
        got = chk.run_window(win, tail_in[:self.halo_len].contiguous())
        want = torch.cat([tail_out, y_slab[:width]])
        rms = float(want.abs().double().pow(2).mean().sqrt())
        err = float((got - want).abs().max())
        return {"max_rel_err": err / rms if rms > 0 else err, "samples": int(got.numel())}

    def run_window(self, x, halo):
        """The cascade from zero state over [halo ++ x], keeping the x part (what ``run`` does on ranks
        > 0, for an arbitrary window)."""
        keep_rank, keep_bounds = self.rank, (self.start, self.end)
        try:
            self.rank = max(self.rank, 1)
            self.start, self.end = 0, x.numel()
            return self.run(x, halo=halo)
        finally:
            self.rank = keep_rank
            self.start, self.end = keep_bounds
