"""ctypes binding of libdwm_b200.so (the C ABI declared in include/dwm_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdwm_b200.so")

DWM_BF16, DWM_F16, DWM_F32 = 0, 1, 2
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_SILU = 0, 1, 2, 3
EPI_STORE, EPI_GEGLU, EPI_QKNORM, EPI_RESID, EPI_F32 = 0, 1, 2, 3, 4

_i64 = ctypes.c_int64
_p = ctypes.c_void_p


class LinearArgs(ctypes.Structure):
    _fields_ = [
        ("M", _i64), ("N", _i64), ("K", _i64),
        ("A", _p), ("lda", _i64),
        ("W", _p), ("ldw", _i64),
        ("bias", _p),
        ("dtype", ctypes.c_int), ("epilogue", ctypes.c_int),
        ("act", ctypes.c_int),
        ("out", _p), ("ldo", _i64),
        ("rows_per_item", _i64), ("out_item_stride", _i64),
        ("out_row_offset", _i64),
        ("q_norm_weight", _p), ("k_norm_weight", _p), ("qk_region", _i64),
        ("eps", ctypes.c_float),
        ("resid", _p), ("ldr", _i64), ("resid_row_mod", _i64),
        ("gate", _p), ("gate_ld", _i64),
        ("blend_x", _p), ("ldx", _i64),
        ("alpha", _p), ("rows_per_batch", _i64),
    ]


_lib = None

# name -> (restype, argtypes); every symbol include/dwm_b200.h declares.
SYMBOLS = {
    "dwm_b200_version": (ctypes.c_char_p, []),
    "dwm_b200_last_error": (ctypes.c_char_p, []),
    "dwm_b200_linear": (ctypes.c_int, [ctypes.POINTER(LinearArgs), _p]),
}


def load():
    """Loads the shared library (building nothing). Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "{} not found: run `python -m opendwm_b200.build` (no CPU "
                "fallback exists)".format(LIB_PATH))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().dwm_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError("{} failed (rc={}): {}".format(what, rc, msg))
