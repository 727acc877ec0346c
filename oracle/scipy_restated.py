"""First-principles restatement (pure numpy) of the scipy.signal routines the
reference's hot path delegates to.

TEST INFRASTRUCTURE ONLY (see oracle/ddoracle.py).  The reference's arithmetic lives in
scipy.signal (third-party, pinned scipy==1.0.0 in the reference's requirements.txt:11,
not vendored under the reference tree).  These functions restate the *published*
algorithms so the oracle does not rest on "scipy is a black box"; tests/test_oracle.py
checks each against the installed scipy 1.18 on seeded inputs.  They are plain loops /
FFTs and are only meant for small cases.

Call sites in the reference: lfilter filters.py:69,75; lfilter_zi filters.py:45;
lfiltic filters.py:67; filtfilt filters.py:73; hilbert demod_am.py:29; resample
comm.py:114; correlate decode_noaa.py:671; np.convolve decode_noaa.py:672.
"""

from __future__ import annotations

import numpy as np


def lfilter(b, a, x, zi=None):
    """Direct-form II transposed IIR/FIR filter (the structure scipy's lfilter uses).

        y[n]   = b0 x[n] + z0
        z[i]   = b[i+1] x[n] + z[i+1] - a[i+1] y[n]      (i < order-1)
        z[o-1] = b[o]   x[n]          - a[o]   y[n]

    with b, a normalised by a[0] and zero-padded to equal length.  Returns y, or
    (y, zf) when an initial state is given.
    """
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    n = max(len(a), len(b))
    b = np.concatenate([b, np.zeros(n - len(b))]) / a[0]
    a = np.concatenate([a, np.zeros(n - len(a))]) / a[0]
    x = np.asarray(x)
    dt = np.result_type(x.dtype, np.float64, zi.dtype if zi is not None else np.float64)
    z = np.zeros(n - 1, dtype=dt) if zi is None else np.array(zi, dtype=dt)
    y = np.zeros(len(x), dtype=dt)
    for k in range(len(x)):
        xk = x[k]
        yk = b[0] * xk + (z[0] if n > 1 else 0.0)
        for i in range(n - 2):
            z[i] = b[i + 1] * xk + z[i + 1] - a[i + 1] * yk
        if n > 1:
            z[n - 2] = b[n - 1] * xk - a[n - 1] * yk
        y[k] = yk
    return y if zi is None else (y, z)


def lfilter_zi(b, a):
    """Steady-state DF-II-T state for a unit step input: solve (I - A^T) zi = B with the
    companion matrix A of ``a`` and B = b[1:] - a[1:] b[0]."""
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if a[0] != 1.0:
        b = b / a[0]
        a = a / a[0]
    n = max(len(a), len(b))
    a = np.concatenate([a, np.zeros(n - len(a))])
    b = np.concatenate([b, np.zeros(n - len(b))])
    # companion(a).T : first column is -a[1:], ones on the super-diagonal
    comp_t = np.zeros((n - 1, n - 1))
    comp_t[:, 0] = -a[1:]
    for i in range(n - 2):
        comp_t[i, i + 1] = 1.0
    rhs = b[1:] - a[1:] * b[0]
    return np.linalg.solve(np.eye(n - 1) - comp_t, rhs)


def lfiltic(b, a, y, x=None):
    """Initial DF-II-T state from past outputs y[-1], y[-2].. and inputs x[-1], x[-2].."""
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    nb, na = len(b) - 1, len(a) - 1          # N, M in scipy's notation (b order, a order)
    k = max(nb, na)
    y = np.asarray(y)
    if len(y) < na:
        y = np.concatenate([y, np.zeros(na - len(y))])
    if x is None:
        x = np.zeros(nb)
    else:
        x = np.asarray(x)
        if len(x) < nb:
            x = np.concatenate([x, np.zeros(nb - len(x))])
    zi = np.zeros(k, dtype=np.result_type(y.dtype, x.dtype, np.float64))
    for m in range(nb):
        zi[m] = np.sum(b[m + 1:] * x[:nb - m])
    for m in range(na):
        zi[m] -= np.sum(a[m + 1:] * y[:na - m])
    return zi / a[0]


def filtfilt(b, a, x):
    """Zero-phase filter with scipy's defaults: odd extension of 3*max(len a, len b)
    samples at both ends, forward pass seeded with zi*ext[0], backward pass seeded with
    zi*fwd[-1], then the padding is trimmed."""
    x = np.asarray(x)
    ntaps = max(len(a), len(b))
    edge = 3 * ntaps
    if len(x) <= edge:
        raise ValueError("input too short for the default filtfilt padding")
    left = 2 * x[0] - x[edge:0:-1]
    right = 2 * x[-1] - x[-2:-edge - 2:-1]
    ext = np.concatenate([left, x, right])
    zi = lfilter_zi(b, a)
    fwd, _ = lfilter(b, a, ext, zi * ext[0])
    bwd, _ = lfilter(b, a, fwd[::-1], zi * fwd[-1])
    return bwd[::-1][edge:-edge]


def hilbert(x):
    """Analytic signal by FFT masking: keep DC (and Nyquist for even N) once, double the
    positive frequencies, zero the negative ones."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    spec = np.fft.fft(x)
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1
        h[1:n // 2] = 2
    else:
        h[0] = 1
        h[1:(n + 1) // 2] = 2
    return np.fft.ifft(spec * h)


def resample(x, num):
    """Fourier-domain resampling of a real signal: keep the lowest min(num, n)//2+1 rfft bins.  Down,
    ``num`` even: the new Nyquist bin collects both aliases (doubled, real part survives the irfft).
    Up, ``n`` even: the old Nyquist bin is shared between the two frequencies it now stands for
    (halved), whatever the parity of ``num``.  Rescale by num/n."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    spec = np.fft.rfft(x)
    keep = num // 2 + 1
    out = np.zeros(keep, dtype=np.complex128)
    m = min(keep, len(spec))
    out[:m] = spec[:m]
    if num % 2 == 0 and num < n:
        out[num // 2] *= 2.0
    elif num > n and n % 2 == 0:
        out[n // 2] *= 0.5
    return np.fft.irfft(out, num) * (float(num) / n)


def correlate_same(h, k):
    """cross-correlation, 'same' mode: out[i] = sum_j h[i - M//2 + j... ] k[j] aligned the
    way scipy centres the 'full' result (start index (M-1)//2 of the full output)."""
    h = np.asarray(h, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    full = np.convolve(h, k[::-1])           # correlation == convolution with reversed k
    start = (len(k) - 1) // 2
    return full[start:start + len(h)]


def convolve_same(h, k):
    """np.convolve(h, k, 'same') for len(h) >= len(k)."""
    full = np.convolve(np.asarray(h, dtype=np.float64), np.asarray(k, dtype=np.float64))
    start = (len(k) - 1) // 2
    return full[start:start + len(h)]
