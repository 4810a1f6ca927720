import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nlpscratch_b200 import _lib
from nlpscratch_b200.array import DeviceArray
n = 256 * 1024 * 1024  # floats = 1 GiB
a = DeviceArray.empty((n,), np.float32); b = DeviceArray.empty((n,), np.float32)
def t(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: _lib.call("nlps_memset", _lib.vp(a.ptr), C.c_int(0), C.c_size_t(n * 4), C.c_void_p(0)))
print("cudaMemset 1 GiB: %.3f ms -> %.0f GB/s write" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: _lib.call("nlps_memcpy_d2d", _lib.vp(b.ptr), _lib.vp(a.ptr), C.c_size_t(n * 4), C.c_void_p(0)))
print("cudaMemcpy d2d 1 GiB: %.3f ms -> %.0f GB/s read+write" % (ms, 2 * n * 4 / ms / 1e6))
ms = t(lambda: _lib.call("nlps_unary_f32", C.c_void_p(0), C.c_int(4), C.c_int64(n), _lib.vp(a.ptr), _lib.vp(b.ptr)))
print("neg kernel (scalar ld/st) 1 GiB: %.3f ms -> %.0f GB/s read+write" % (ms, 2 * n * 4 / ms / 1e6))
o = DeviceArray.empty((1,), np.float32)
ms = t(lambda: _lib.call("nlps_sumsq", C.c_void_p(0), C.c_int64(n), _lib.vp(a.ptr), _lib.vp(o.ptr)))
print("sumsq (float4 loads) 1 GiB: %.3f ms -> %.0f GB/s read" % (ms, n * 4 / ms / 1e6))
