"""Time the tcgen05 GEMM on every shape of one c4 training step (CUDA events, 20 reps each)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nlpscratch_b200 import _lib
from nlpscratch_b200.array import DeviceArray

N, E, F, V = 32768, 512, 2048, 8192
shapes = [  # name, M, N, K, trans_a, trans_b
    ("fwd qkv   NN", N, 3 * E, E, 0, 0), ("fwd wo    NN", N, E, E, 0, 0), ("fwd ffn1  NN", N, F, E, 0, 0),
    ("fwd ffn2  NN", N, E, F, 0, 0), ("fwd vocab NN", N, V, E, 0, 0),
    ("dgrad dz    NT", N, F, E, 0, 1), ("dgrad dy1   NT", N, E, F, 0, 1), ("dgrad dctx  NT", N, E, E, 0, 1),
    ("dgrad dx    NT", N, E, 3 * E, 0, 1), ("dgrad dH    NT", N, E, V, 0, 1),
    ("wgrad dW2   TN", F, E, N, 1, 0), ("wgrad dW1   TN", E, F, N, 1, 0), ("wgrad dWo   TN", E, E, N, 1, 0),
    ("wgrad dWqkv TN", E, 3 * E, N, 1, 0), ("wgrad dWvoc TN", E, V, N, 1, 0),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(0)
tot_ms, tot_fl = 0.0, 0.0
for name, M, Nn, K, ta, tb in shapes:
    if only and only not in name:
        continue
    a = DeviceArray.from_host(rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32))
    b = DeviceArray.from_host(rng.standard_normal((Nn, K) if tb else (K, Nn)).astype(np.float32))
    c = DeviceArray.empty((M, Nn), np.float32)
    def launch():
        _lib.call("nlps_gemm", C.c_void_p(0), C.c_int(1), C.c_int(M), C.c_int(Nn), C.c_int(K), _lib.vp(a.ptr),
                  C.c_int64(a.shape[1]), C.c_int(ta), _lib.vp(b.ptr), C.c_int64(b.shape[1]), C.c_int(tb), _lib.vp(c.ptr),
                  C.c_int64(Nn), None, C.c_int(0), None, C.c_int64(0), None, C.c_int64(0), C.c_int(0))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * M * Nn * K
    mult = 1 if "voc" in name or "dH" in name else 6
    tot_ms += ms * mult; tot_fl += fl * mult
    print("%-16s M=%6d N=%5d K=%6d  %7.3f ms  %6.1f TFLOP/s" % (name, M, Nn, K, ms, fl / ms / 1e9), flush=True)
    del a, b, c
print("step total (6 layers + vocab): %.2f ms, %.1f TFLOP/s average" % (tot_ms, tot_fl / tot_ms / 1e9))
