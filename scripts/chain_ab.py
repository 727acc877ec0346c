#!/usr/bin/env python
"""A/B of the fused-chain kernels on one GPU: the CTA-tiled kernel (DDM_CHAIN_LEGACY=1) against the
warp-autonomous kernel for several (warps, stages) geometries, full C2 pass, CUDA events, with a
bit-equality check of the outputs.  The knobs are read by ddm_chain_create, so one process can walk
through them.   python scripts/chain_ab.py [samples]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scipy.signal as sps
import torch

from directdemod_b200.fused import FusedChain

torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1843200000
x = torch.empty(n, dtype=torch.complex64, device="cuda")
xr = torch.view_as_real(x).reshape(-1)
for a in range(0, xr.numel(), 1 << 27):
    xr[a:a + (1 << 27)].normal_(0.0, 40.0)
bh = sps.windows.blackmanharris(151)


def run_cfg(env, fs, f, d, fmt="cf32", xin=None, reps=10):
    for k in ("DDM_CHAIN_LEGACY", "DDM_STREAM_WARPS", "DDM_STREAM_STAGES"):
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        ch = FusedChain(bh, d, f, fs, in_format=fmt)
    except Exception as exc:
        return None, str(exc)
    xi = x if xin is None else xin
    nn = xi.numel() if fmt == "cf32" else xi.numel() // 2
    out = torch.empty(ch.out_count(nn) + 2, dtype=torch.float32, device="cuda")

    def go():
        ch.set_position(0, 0, False)
        return ch.apply(xi, out=out)
    for _ in range(3):
        y = go()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        ch.set_position(0, 0, False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = ch.apply(xi, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return (sum(ts) / len(ts), ts[0], y.clone()), None


def main():
    cfgs = [("legacy", {"DDM_CHAIN_LEGACY": "1"}), ("auto", {})]
    for w, s in ((8, 2), (8, 3), (9, 2), (10, 2), (11, 2), (12, 2)):
        cfgs.append(("w%ds%d" % (w, s), {"DDM_STREAM_WARPS": str(w), "DDM_STREAM_STAGES": str(s)}))
    ref = None
    for tag, env in cfgs:
        res, err = run_cfg(env, 2048000, 30000.0, 34)
        if res is None:
            print(json.dumps({"cfg": tag, "error": err}), flush=True)
            continue
        avg, best, y = res
        if ref is None:
            ref = y
        d = (y.double() - ref.double())
        d = torch.atan2(torch.sin(d), torch.cos(d)).abs()
        rec = {"cfg": tag, "D": 34, "ms_avg": round(avg, 4), "ms_min": round(best, 4),
               "GBps": round(n * (8 + 4 / 34) / avg / 1e6, 1), "bit_equal_to_legacy": bool(torch.equal(y, ref)),
               "max_abs_diff_to_legacy": float(d.max()), "rms_diff": float(d.pow(2).mean().sqrt())}
        del d
        print(json.dumps(rec), flush=True)
        del y
    del ref
    # other decimation factors and the u8 path: legacy vs auto
    xu = None
    for tag, fs, f, d, fmt in (("D=50", 10000000, 125000.0, 50, "cf32"), ("D=68", 2048000, 30000.0, 68, "cf32"),
                               ("D=17", 1024000, 30000.0, 17, "cf32"), ("D=33", 2048000, 30000.0, 33, "cf32"),
                               ("D=24", 2048000, 30000.0, 24, "cf32"), ("D=100", 2048000, 30000.0, 100, "cf32"),
                               ("D=34 u8", 2048000, 30000.0, 34, "cu8")):
        xin = None
        if fmt == "cu8":
            m = n // 2
            xu = torch.empty((m, 2), dtype=torch.uint8, device="cuda")
            for a in range(0, m, 1 << 26):
                b = min(m, a + (1 << 26))
                xu[a:b] = (torch.view_as_real(x[a:b]) + 127.5).clamp_(0, 255).to(torch.uint8)
            xin = xu
        outs = {}
        envs = [("legacy", {"DDM_CHAIN_LEGACY": "1"}), ("auto", {})]
        if fmt == "cf32":
            envs += [("w%ds2" % w, {"DDM_STREAM_WARPS": str(w), "DDM_STREAM_STAGES": "2"}) for w in (12, 10, 8, 6, 4)]
        else:
            envs += [("w%ds%d" % (w, st), {"DDM_STREAM_WARPS": str(w), "DDM_STREAM_STAGES": str(st)})
                     for w, st in ((12, 4), (12, 8), (8, 8))]
        for name, env in envs:
            res, err = run_cfg(env, fs, f, d, fmt, xin, reps=5)
            if res is None:
                print(json.dumps({"cfg": name, "case": tag, "error": err}), flush=True)
                continue
            outs[name] = res
            nn = n if fmt == "cf32" else n // 2
            print(json.dumps({"cfg": name, "case": tag, "ms_avg": round(res[0], 4), "ms_min": round(res[1], 4),
                              "Gsps": round(nn / res[0] / 1e6, 1)}), flush=True)
        if "legacy" in outs and "auto" in outs:
            dd = outs["legacy"][2].double() - outs["auto"][2].double()
            dd = torch.atan2(torch.sin(dd), torch.cos(dd)).abs()
            print(json.dumps({"case": tag, "max_abs_diff_legacy_vs_auto": float(dd.max())}), flush=True)
        outs.clear()


if __name__ == "__main__":
    main()
