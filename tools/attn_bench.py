"""Time the fused attention kernels alone at a workload's shapes (CUDA events)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nlpscratch_b200 import _lib
from nlpscratch_b200.array import DeviceArray

B, S, h, d = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (64, 512, 8, 64)))
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
E = h * d
rng = np.random.default_rng(0)
qkv = DeviceArray.from_host(rng.standard_normal((B * S, 3 * E)).astype(np.float32))
x = DeviceArray.from_host(rng.integers(1, 100, (B, S)).astype(np.int32))
dctx = DeviceArray.from_host(rng.standard_normal((B * S, E)).astype(np.float32))
ctx, lse, dqkv = DeviceArray.empty((B * S, E)), DeviceArray.empty((B, h, S)), DeviceArray.empty((B * S, 3 * E))
S0 = C.c_void_p(0)
def fwd():
    _lib.call("nlps_attention_forward", S0, C.c_int(1), C.c_int(B), C.c_int(S), C.c_int(h), C.c_int(d),
              _lib.vp(qkv.ptr), _lib.vp(x.ptr), _lib.vp(ctx.ptr), _lib.vp(lse.ptr))
def bwd():
    _lib.call("nlps_attention_backward", S0, C.c_int(1), C.c_int(B), C.c_int(S), C.c_int(h), C.c_int(d),
              _lib.vp(qkv.ptr), _lib.vp(x.ptr), _lib.vp(ctx.ptr), _lib.vp(lse.ptr), _lib.vp(dctx.ptr), _lib.vp(dqkv.ptr))
def t(fn):
    # median of 5 trials of `reps` launches after a generous warm-up (first launches pay module load and clocks)
    import time
    t0 = time.time()
    while time.time() - t0 < (0.02 if os.environ.get('NLPS_BENCH_FAST') else 0.4):
        for _ in range(10): fn()
        torch.cuda.synchronize()
    out = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps)
    if os.environ.get('NLPS_BENCH_VERBOSE'): print('   trials (ms):', ' '.join('%.3f' % v for v in out))
    return sorted(out)[2]
flops_fwd = 2.0 * 2 * B * h * S * S * d / 2      # causal half
tf, tb = t(fwd), t(bwd)
print("attention B=%d S=%d h=%d d=%d: fwd %.3f ms (%.1f TFLOP/s causal), bwd %.3f ms (%.1f TFLOP/s, 2.5x fwd flops)" % (
    B, S, h, d, tf, flops_fwd / tf / 1e9, tb, 2.5 * flops_fwd / tb / 1e9))
