import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, scipy.signal as sps
from directdemod_b200.fused import FusedChain
torch.cuda.set_device(0)
n = 1843200000
D = int(os.environ["DSWEEP_D"])
x = torch.empty(n, dtype=torch.complex64, device="cuda")
xr = torch.view_as_real(x).reshape(-1)
for a in range(0, xr.numel(), 1 << 27):
    xr[a:a + (1 << 27)].normal_(0.0, 40.0)
try:
    ch = FusedChain(sps.windows.blackmanharris(151), D, 30000.0, 2048000)
    out = torch.empty(ch.out_count(n) + 2, dtype=torch.float32, device="cuda")
    for _ in range(4):
        ch.set_position(0, 0, False); ch.apply(x, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(16):
        ch.set_position(0, 0, False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch.apply(x, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(json.dumps({"ms_med": round(ts[8], 4), "ms_min": round(ts[0], 4)}))
except Exception as exc:
    print(json.dumps({"error": str(exc)[:100]}))
''' % ROOT
cases = {50: [(8, 1), (12, 1), (16, 1), (8, 2), (6, 2)], 68: [(8, 1), (12, 1), (4, 3), (6, 2), (4, 2)],
         100: [(8, 1), (4, 2), (6, 1)], 40: [(8, 2), (12, 1), (16, 1)], 20: [(12, 2), (16, 1), (16, 2)]}
for D, geos in cases.items():
    variants = [("auto", {}), ("cta-tiled", {"DDM_CHAIN_LEGACY": "1"})] + \
        [("w%ds%d" % g, {"DDM_STREAM_WARPS": str(g[0]), "DDM_STREAM_STAGES": str(g[1])}) for g in geos]
    for name, env in variants:
        e = dict(os.environ, DSWEEP_D=str(D))
        for k in ("DDM_CHAIN_LEGACY", "DDM_STREAM_WARPS", "DDM_STREAM_STAGES"):
            e.pop(k, None)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", CHILD], env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        print(D, name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-200:], flush=True)
