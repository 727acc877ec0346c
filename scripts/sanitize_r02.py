#!/usr/bin/env python
"""Small end-to-end exercises of the kernels added in round 2, for compute-sanitizer (memcheck /
racecheck): the warp-autonomous fused chain (cf32, u8, odd D, batch, ragged chunks, 16- and 8/12-warp
geometries), the paired Hilbert envelope, the cascade.   compute-sanitizer python scripts/sanitize_r02.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.signal as sps
import torch

from directdemod_b200 import fftops, filters
from directdemod_b200.fused import FusedChain

torch.cuda.set_device(0)
rng = np.random.default_rng(5)
bh = sps.windows.blackmanharris(151)


def wrap(a, b):
    return float(np.max(np.abs(np.angle(np.exp(1j * (a.astype(np.float64) - b.astype(np.float64)))))))


n = 700003
x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 40).astype(np.complex64)
xd = torch.from_numpy(x).cuda()
for d, fs in ((34, 2048000), (33, 2048000), (17, 1024000), (24, 2048000), (200, 2048000)):
    os.environ["DDM_CHAIN_LEGACY"] = "1"
    ref = FusedChain(bh, d, 30000.0, fs).apply(xd).cpu().numpy() if d != 200 else None
    os.environ.pop("DDM_CHAIN_LEGACY")
    ch = FusedChain(bh, d, 30000.0, fs)
    got = torch.cat([ch.apply(xd[a:b]) for a, b in ((0, 250001), (250001, 250002), (250002, 250040), (250040, n))]).cpu().numpy()
    print("chain D=%d pieces vs cta-tiled whole: %s" % (d, "n/a" if ref is None else "%.2e" % wrap(got, ref)), got.shape, flush=True)
xb = xd[:600000].reshape(3, 200000)
yb = FusedChain(bh, 34, 30000.0, 2048000).apply_batch(xb)
y1 = FusedChain(bh, 34, 30000.0, 2048000).apply(xb[1].contiguous())
print("batch row equals single:", bool(torch.equal(yb[1], y1)), flush=True)
xu = (torch.view_as_real(xd) + 127.5).clamp_(0, 255).to(torch.uint8)
c8 = FusedChain(bh, 34, 30000.0, 2048000, in_format="cu8")
g8 = torch.cat([c8.apply(xu[a:b]) for a, b in ((0, 300001), (300001, n))])
print("u8 chain:", tuple(g8.shape), flush=True)
# paired Hilbert envelope: odd number of chunks + ragged tail
aud = torch.from_numpy(rng.standard_normal(5 * 24000 + 777).astype(np.float32)).cuda()
env = fftops.hilbert_envelope(aud, 24000).cpu().numpy()
want = np.concatenate([np.abs(sps.hilbert(aud.cpu().numpy()[a:a + 24000].astype(np.float64))) for a in range(0, aud.numel(), 24000)])
print("hilbert pairs rel %.2e" % (np.sqrt(np.mean((env - want) ** 2) / np.mean(want ** 2))), flush=True)
# cascade
fir = filters.remez(2400000, [[0, 100000], [120000, 1199999]], [1, 0], ntaps=255)
iir = filters.butter(2400000, 100000, n=8)
xc = x[:1200007]
w = xc.astype(np.complex128)
for f in (fir, iir):
    w, _ = sps.lfilter(f.getB, f.getA, w, zi=sps.lfilter_zi(f.getB, f.getA))
cas = filters.cascade([fir, iir]).setFIRMode(2)
g = np.concatenate([cas.applyOn(xc[:1100000]), cas.applyOn(xc[1100000:])])
print("cascade rel %.2e" % np.sqrt(np.mean(np.abs(g - w) ** 2) / np.mean(np.abs(w) ** 2)), len(cas.getB), flush=True)
