"""Host <-> device handoff helpers shared by the operator classes.

PyTorch is used here only to own device memory and streams ("tensor handoff"); all
arithmetic is done by libddemod.so.  Device signals are float32 (real) or complex64;
arrays handed back to the caller use the reference's dtypes (float64 / complex128, what
numpy/scipy produce in the reference) so downstream host code sees what it always saw.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


_checked = False


def require_cuda():
    global _checked
    t = torch()
    if not _checked:
        if not t.cuda.is_available():
            raise RuntimeError("directdemod_b200 needs a CUDA device; there is no CPU fallback")
        _lib.lib()
        _checked = True
    return t


def device_index():
    return torch().cuda.current_device()


def stream_ptr(dev=None):
    t = torch()
    return C.c_void_p(t.cuda.current_stream(device_index() if dev is None else dev).cuda_stream)


def is_tensor(x):
    t = torch()
    return isinstance(x, t.Tensor)


def ptr(tensor):
    return C.c_void_p(tensor.data_ptr())


def to_device(x):
    """Any 1-D array-like -> contiguous cuda tensor, float32 or complex64."""
    t = require_cuda()
    if isinstance(x, t.Tensor):
        if x.dim() != 1:
            raise TypeError("The signal array must be 1-D")
        if x.is_cuda and x.dtype in (t.complex64, t.float32) and x.is_contiguous():
            return x                                     # the common case on the chunk loops: nothing to do
        want = t.complex64 if x.is_complex() else t.float32
        if not x.is_cuda:
            x = x.to("cuda")
        return x.to(want).contiguous()
    a = np.asarray(x)
    if a.ndim != 1:
        raise TypeError("The signal array must be 1-D")
    a = np.ascontiguousarray(a, dtype=np.complex64 if np.iscomplexobj(a) else np.float32)
    return t.from_numpy(a).to("cuda")


def to_host(tensor):
    """cuda tensor -> numpy array in the reference's dtype (float64 / complex128)."""
    a = tensor.detach().cpu().numpy()
    return a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)


def empty_like_kind(n, complex_, dev=None):
    t = torch()
    return t.empty(int(n), dtype=t.complex64 if complex_ else t.float32,
                   device="cuda:%d" % (device_index() if dev is None else dev))
