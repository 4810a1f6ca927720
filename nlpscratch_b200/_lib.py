"""ctypes binding of libnlps_b200.so (the C ABI declared in include/nlps_b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if no CUDA
device is visible every compute call raises ``RuntimeError`` (from the library's own status).
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnlps_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nlps_b200.h")

PREC = {"fp32": 0, "tf32": 1, "tf32x3": 2}
PARAM_IDS = {n: i for i, n in enumerate(
    ("We", "Wq", "Wk", "Wv", "Wo", "W1", "W2", "b1", "b2", "W_vocab", "b_vocab",
     "ln1_gamma", "ln1_beta", "ln2_gamma", "ln2_beta"))}
BUF_PARAM, BUF_GRAD, BUF_M, BUF_V = 0, 1, 2, 3


class Config(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("embed_dim", C.c_int32), ("num_heads", C.c_int32),
                ("ff_dim", C.c_int32), ("num_layers", C.c_int32), ("max_len", C.c_int32),
                ("dropout_p", C.c_float), ("use_adam", C.c_int32), ("precision", C.c_int32),
                ("seed", C.c_uint64)]


class View(C.Structure):
    _fields_ = [("d_ptr", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 3),
                ("stride", C.c_int64 * 3)]


_lib = None
_lock = threading.Lock()
current_stream = C.c_void_p(0)        # cudaStream_t all shim ops and models launch on (default stream)


def declared_symbols():
    """Every function name declared in include/nlps_b200.h."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nlps_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load (building first if needed) the shared library; idempotent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import build
            build.build_library()
        lib = C.CDLL(LIB_PATH)
        lib.nlps_last_error.restype = C.c_char_p
        for name in declared_symbols():
            fn = getattr(lib, name)           # raises AttributeError if the .so lacks a declared symbol
            if name != "nlps_last_error":
                fn.restype = C.c_int
        _lib = lib
    return _lib


def check(status: int):
    if status != 0:
        msg = load().nlps_last_error()
        raise RuntimeError("libnlps_b200: " + (msg.decode() if msg else "error %d" % status))


def call(name: str, *args):
    check(getattr(load(), name)(*args))


def device_count() -> int:
    n = C.c_int(0)
    call("nlps_device_count", C.byref(n))
    return n.value


def require_device():
    if device_count() <= 0:
        raise RuntimeError("nlpscratch_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def stream_ptr():
    return current_stream


def set_stream(ptr: int):
    """Route shim ops (and models created afterwards) to another cudaStream_t."""
    global current_stream
    current_stream = C.c_void_p(int(ptr))


def vp(ptr: int) -> C.c_void_p:
    return C.c_void_p(int(ptr))
