"""Thin ctypes helpers so the GPU tests can call single operators of the C ABI (test infrastructure)."""
import ctypes as C

import numpy as np

from nlpscratch_b200 import _lib
from nlpscratch_b200.array import DeviceArray

S0 = C.c_void_p(0)


def dev(a, dtype=None):
    return DeviceArray.from_host(np.asarray(a), dtype)


def gemm(a, b, prec="fp32", trans_a=False, trans_b=False, bias=None, relu=False, gate=None, residual=None,
         accumulate_into=None):
    """C = op(A) op(B) through nlps_gemm; a/b are host arrays in their STORED layout."""
    A, B = dev(a, np.float32), dev(b, np.float32)
    M = a.shape[1] if trans_a else a.shape[0]
    K = a.shape[0] if trans_a else a.shape[1]
    N = b.shape[0] if trans_b else b.shape[1]
    out = dev(accumulate_into, np.float32) if accumulate_into is not None else DeviceArray.empty((M, N), np.float32)
    bias_d = dev(bias, np.float32) if bias is not None else None
    gate_d = dev(gate, np.float32) if gate is not None else None
    res_d = dev(residual, np.float32) if residual is not None else None
    _lib.call("nlps_gemm", S0, C.c_int(_lib.PREC[prec]), C.c_int(M), C.c_int(N), C.c_int(K),
              _lib.vp(A.ptr), C.c_int64(a.shape[1]), C.c_int(int(trans_a)),
              _lib.vp(B.ptr), C.c_int64(b.shape[1]), C.c_int(int(trans_b)),
              _lib.vp(out.ptr), C.c_int64(N),
              _lib.vp(bias_d.ptr) if bias_d is not None else None, C.c_int(int(relu)),
              _lib.vp(gate_d.ptr) if gate_d is not None else None, C.c_int64(N),
              _lib.vp(res_d.ptr) if res_d is not None else None, C.c_int64(N),
              C.c_int(1 if accumulate_into is not None else 0))
    return out.get()


def gemm_ref(a, b, trans_a=False, trans_b=False, bias=None, relu=False, gate=None, residual=None, acc=None):
    A = a.astype(np.float64).T if trans_a else a.astype(np.float64)
    B = b.astype(np.float64).T if trans_b else b.astype(np.float64)
    c = A @ B
    if bias is not None:
        c = c + bias
    if relu:
        c = np.maximum(c, 0)
    if gate is not None:
        c = np.where(gate > 0, c, 0)
    if residual is not None:
        c = c + residual
    if acc is not None:
        c = c + acc
    return c
