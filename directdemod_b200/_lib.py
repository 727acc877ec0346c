"""ctypes binding of csrc/libddemod.so (the C ABI declared in include/ddemod.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module
raises.  PyTorch is used by the callers only to own device memory and streams; nothing in
the signatures below is a torch type.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDEMOD_LIB") or os.path.join(_HERE, "csrc", "libddemod.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY, ERR_UNSUPPORTED = -1, -2, -3, -4, -5

CHAIN_OUT_FM, CHAIN_OUT_IQ = 0, 1
IN_CF32, IN_CU8 = 0, 1

_vp, _i64, _int, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
_pi64, _pint = C.POINTER(C.c_int64), C.POINTER(C.c_int)
_pdbl = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/ddemod.h one to one
SIGNATURES = {
    "ddm_version": (_int, []),
    "ddm_last_error": (C.c_char_p, []),
    "ddm_launch_count": (_i64, []),
    "ddm_device_count": (_int, [_pint]),
    "ddm_release_scratch": (_int, [_int]),
    "ddm_chain_create": (_int, [_int, _pdbl, _int, _int, _dbl, _dbl, _int, _int,
                                C.POINTER(_vp)]),
    "ddm_chain_destroy": (_int, [_vp]),
    "ddm_chain_reset": (_int, [_vp]),
    "ddm_chain_halo_len": (_int, [_vp, _pi64]),
    "ddm_chain_out_count": (_int, [_vp, _i64, _pi64]),
    "ddm_chain_get_position": (_int, [_vp, _pi64, _pi64, _pint]),
    "ddm_chain_set_position": (_int, [_vp, _i64, _i64, _int, _vp, _vp]),
    "ddm_chain_get_halo": (_int, [_vp, _vp, _vp]),
    "ddm_chain_export_state": (_int, [_vp, _pdbl, _pdbl, _vp]),
    "ddm_chain_apply_dev": (_int, [_vp, _vp, _i64, _vp, _i64, _pi64, _vp]),
    "ddm_chain_apply_batch_dev": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _pi64, _vp]),
    "ddm_chain_apply_host": (_int, [_vp, _vp, _i64, _vp, _i64, _pi64, _vp]),
    "ddm_mix_cf32": (_int, [_int, _vp, _i64, _dbl, _dbl, _i64, _vp]),
    "ddm_mix_var_cf32": (_int, [_int, _vp, _vp, _i64, _dbl, _i64, _vp]),
    "ddm_fm_demod": (_int, [_int, _vp, _i64, _vp, _vp, _pi64, _vp]),
    "ddm_fm_angle_diff": (_int, [_int, _vp, _i64, _vp, _vp, _vp, _pi64, _vp]),
    "ddm_abs": (_int, [_int, _vp, _i64, _int, _vp, _vp]),
    "ddm_sign": (_int, [_int, _vp, _i64, _vp, _vp]),
    "ddm_stride_copy": (_int, [_int, _vp, _i64, _int, _i64, _i64, _vp, _pi64, _vp]),
    "ddm_medfilt": (_int, [_int, _vp, _i64, _int, _vp, _vp]),
    "ddm_cu8_to_cf32": (_int, [_int, _vp, _i64, _vp, _vp]),
    "ddm_filter_create": (_int, [_int, _pdbl, _int, _pdbl, _int, C.POINTER(_vp)]),
    "ddm_filter_destroy": (_int, [_vp]),
    "ddm_filter_state_len": (_int, [_vp, _pint]),
    "ddm_filter_set_state": (_int, [_vp, _pdbl, _vp]),
    "ddm_filter_get_state": (_int, [_vp, _pdbl, _vp]),
    "ddm_filter_reset": (_int, [_vp, _vp]),
    "ddm_filter_set_zi_base": (_int, [_vp, _pdbl]),
    "ddm_filter_set_iir_mode": (_int, [_vp, _int]),
    "ddm_filter_set_iir_auto_floor": (_int, [_vp, _dbl]),
    "ddm_filter_set_fir_mode": (_int, [_vp, _int]),
    "ddm_filter_info": (_int, [_vp, _pint, _pi64, _pdbl]),
    "ddm_filter_apply_dev": (_int, [_vp, _vp, _i64, _int, _vp, _int, _vp]),
    "ddm_filter_filtfilt_dev": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "ddm_iir_analyse": (_int, [_pdbl, _int, _pdbl, _int, _pi64, _pdbl]),
    "ddm_lfilter_zi": (_int, [_pdbl, _int, _pdbl, _int, _pdbl]),
    "ddm_fft_create": (_int, [_int, C.POINTER(_vp)]),
    "ddm_fft_destroy": (_int, [_vp]),
    "ddm_am_hilbert": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "ddm_resample": (_int, [_vp, _vp, _i64, _int, _i64, _vp, _vp]),
    "ddm_mix_rows_cf32": (_int, [_int, _vp, _i64, _i64, _dbl, _dbl, _vp]),
    "ddm_filter_filtfilt_rows_dev": (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp]),
    "ddm_fm_demod_rows": (_int, [_int, _vp, _i64, _i64, _vp, _vp]),
    "ddm_rows_argmax": (_int, [_int, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "ddm_rows_mean": (_int, [_int, _vp, _vp, _i64, _i64, _vp, _vp]),
    "ddm_resample_rows": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "ddm_row_medians": (_int, [_int, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "ddm_correlate": (_int, [_int, _vp, _i64, _int, _pdbl, _int, _int, _vp, _vp]),
    "ddm_topk_sums": (_int, [_int, _vp, _i64, _i64, _pdbl, _pdbl, _vp]),
    "ddm_compact_above": (_int, [_int, _vp, _i64, _dbl, _vp, _vp, _i64, _pi64, _vp]),
    "ddm_pick_peaks": (_int, [_int, _vp, _i64, _dbl, _dbl, _pi64, _i64, _pi64, _vp]),
    "ddm_group_peaks": (_int, [_pi64, _pdbl, _i64, _dbl, _pi64, _i64, _pi64]),
    "ddm_bank4": (_int, [_int, _vp, _i64, _int, _pdbl, _int, _vp, _vp]),
}

_lib = None


class DdmError(RuntimeError):
    def __init__(self, code, where, msg):
        super().__init__("%s failed (%d): %s" % (where, code, msg))
        self.code = code


def lib():
    """The loaded library.  Raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libddemod.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C directdemod_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code, where):
    if code != OK:
        msg = lib().ddm_last_error()
        raise DdmError(code, where, msg.decode("utf-8", "replace") if msg else "")
    return code


def launch_count() -> int:
    return int(lib().ddm_launch_count())
