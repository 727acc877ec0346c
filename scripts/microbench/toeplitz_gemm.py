#!/usr/bin/env python
"""MEASUREMENT ONLY (cuBLAS through torch.matmul, not a product path): how fast could the 1023-tap FIR
run as a Toeplitz GEMM on the tensor cores?

    y[128 g + j] = sum_m T[j][m] x[128 g - (K-1) + m],   T[j][m] = h[K-1 + j - m]   (128 x 1152, Toeplitz)

i.e. Y[g][j] = X[g][:] . T[j][:] with X[g] = the 1152-sample window in front of output group g.  The
window matrix is MATERIALISED here (9x the signal) so that the library GEMM can run at its own speed:
a hand-written tcgen05 kernel that builds the windows in shared memory could at best match that rate.
TF32 keeps 10 mantissa bits, so the 1e-5 tolerance needs the 3-term split (x_hi h_hi + x_lo h_hi +
x_hi h_lo); a complex signal is two real ones.  Prints the GEMM rates and the equivalent complex
samples per second next to the overlap-save FFT kernel's measured figure."""
import json
import sys

import numpy as np
import scipy.signal as sps
import torch

torch.cuda.set_device(0)
K, G = 1023, 128
KW = K - 1 + G                       # 1150
KP = (KW + 31) // 32 * 32            # 1152
h = sps.remez(K, [0, 100000, 120000, 1199999], [1, 0], fs=2400000)
T = np.zeros((G, KP))
for j in range(G):
    for m in range(KW):
        k = K - 1 + j - m
        if 0 <= k < K:
            T[j, m] = h[k]
groups = 1 << 19                     # 67 M real samples per GEMM
n = groups * G
x = torch.randn(n + KP, device="cuda", dtype=torch.float32) * 40
X = x.unfold(0, KP, G)[:groups].contiguous()            # [groups, 1152], materialised windows
Tt = torch.from_numpy(T.T.copy()).cuda()                 # [1152, 128] float64


def split_tf32(a):
    hi = (a.view(torch.int32) & ~0x1FFF).view(torch.float32)       # 10 explicit mantissa bits
    return hi, a - hi


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


flop = 2.0 * groups * KP * G
out = {}
# --- TF32 tensor cores
torch.backends.cuda.matmul.allow_tf32 = True
T32 = Tt.float()
ms = timed(lambda: X @ T32)
out["tf32_gemm_ms"] = round(ms, 3)
out["tf32_tflops"] = round(flop / ms / 1e9, 1)
xh, xl = split_tf32(X)
th, tl = split_tf32(T32)
ms3 = timed(lambda: (xh @ th) + (xl @ th) + (xh @ tl))
y3 = (xh @ th) + (xl @ th) + (xh @ tl)
y1 = X @ T32
want = (X[:4096].double() @ Tt)
err3 = float(((y3[:4096].double() - want).pow(2).mean() / want.pow(2).mean()).sqrt())
err1 = float(((y1[:4096].double() - want).pow(2).mean() / want.pow(2).mean()).sqrt())
out["tf32_x1_rel_rms"] = err1
out["tf32_x3_rel_rms"] = err3
out["tf32_x3_ms_incl_adds"] = round(ms3, 3)
# complex samples per second: two real signals, three products each
out["tf32_x3_equiv_complex_gsps"] = round(n / (2 * 3 * ms) / 1e6, 1)
# --- BF16 tensor cores (rate only: a 3-term bf16 split holds 24 bits but needs 6 products)
Xb, Tb = X.bfloat16(), T32.bfloat16()
msb = timed(lambda: Xb @ Tb)
out["bf16_gemm_ms"] = round(msb, 3)
out["bf16_tflops"] = round(flop / msb / 1e9, 1)
out["bf16_x6_equiv_complex_gsps"] = round(n / (2 * 6 * msb) / 1e6, 1)
# --- plain fp32 (CUDA cores) for reference
torch.backends.cuda.matmul.allow_tf32 = False
msf = timed(lambda: X @ T32, reps=2)
out["fp32_gemm_ms"] = round(msf, 3)
out["fp32_tflops"] = round(flop / msf / 1e9, 1)
out["shape"] = "[%d x %d] x [%d x %d], %d real samples per GEMM" % (groups, KP, KP, G, n)
out["fft_kernel_complex_gsps_measured"] = 162.8
print(json.dumps(out))
