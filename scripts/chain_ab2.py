#!/usr/bin/env python
"""Same-box, interleaved A/B of the D = 34 chain kernels (one process per variant, three rounds):
CTA-tiled kernel, warp-autonomous generic, warp-autonomous specialised (D, zero-tap prefix compile-time)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, scipy.signal as sps
from directdemod_b200.fused import FusedChain
torch.cuda.set_device(0)
n = 1843200000
x = torch.empty(n, dtype=torch.complex64, device="cuda")
xr = torch.view_as_real(x).reshape(-1)
for a in range(0, xr.numel(), 1 << 27):
    xr[a:a + (1 << 27)].normal_(0.0, 40.0)
ch = FusedChain(sps.windows.blackmanharris(151), 34, 30000.0, 2048000)
out = torch.empty(ch.out_count(n) + 2, dtype=torch.float32, device="cuda")
for _ in range(5):
    ch.set_position(0, 0, False); ch.apply(x, out=out)
torch.cuda.synchronize()
ts = []
for _ in range(30):
    ch.set_position(0, 0, False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.apply(x, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(json.dumps({"ms_med": round(ts[15], 4), "ms_min": round(ts[0], 4), "ms_p90": round(ts[27], 4)}))
''' % ROOT
variants = [("warp w8s2 (default)", {}),
            ("warp w16s1 (128-register build)", {"DDM_STREAM_WARPS": "16", "DDM_STREAM_STAGES": "1"}),
            ("warp w12s1", {"DDM_STREAM_WARPS": "12", "DDM_STREAM_STAGES": "1"}),
            ("warp w8s1", {"DDM_STREAM_WARPS": "8", "DDM_STREAM_STAGES": "1"}),
            ("warp w12s2", {"DDM_STREAM_WARPS": "12", "DDM_STREAM_STAGES": "2"})]
for rnd in range(2):
    for name, env in variants:
        e = dict(os.environ)
        for k in ("DDM_CHAIN_LEGACY", "DDM_STREAM_GENERIC", "DDM_STREAM_WARPS", "DDM_STREAM_STAGES"):
            e.pop(k, None)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", CHILD], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
        print(json.dumps({"round": rnd, "variant": name, "result": line}), flush=True)
