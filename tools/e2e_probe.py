"""Where does the end-to-end leg's extra time per step come from?  Variants of the timed loop on the c4 workload."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nlpscratch_b200 import HostScalarSlot, Transformer, _lib
from nlpscratch_b200.array import DeviceArray

L, E, h, F, S, B, V = 6, 512, 8, 2048, 512, 64, 8192
np.random.seed(1234)
model = Transformer(V, E, h, F, L, S, dropout_p=0.1, precision="tf32")
rng = np.random.default_rng(0)
xs = [rng.integers(1, V, (B, S)).astype(np.int32) for _ in range(4)]
ys = [rng.integers(1, V, (B, S)).astype(np.int32) for _ in range(4)]
xd = [DeviceArray.from_host(a) for a in xs]
yd = [DeviceArray.from_host(a) for a in ys]
pin = []
for a in xs + ys:
    p = C.c_void_p()
    _lib.call("nlps_host_alloc_pinned", C.byref(p), C.c_size_t(a.nbytes))
    C.memmove(p.value, a.ctypes.data, a.nbytes)
    pin.append(p)
x_in, y_in = DeviceArray.empty((B, S), np.int32), DeviceArray.empty((B, S), np.int32)
x_in.bounds = y_in.bounds = (1, V - 1)
nbytes = xs[0].nbytes
slots = [HostScalarSlot(model), HostScalarSlot(model)]
mstream = C.c_void_p()
_lib.call("nlps_model_stream", model._h, C.byref(mstream))


def run(name, n, body):
    for i in range(6):
        body(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        body(i)
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / n
    torch.cuda.synchronize()
    print("%-34s %.3f ms/step (host enqueue %.3f ms/step)" % (name, e0.elapsed_time(e1) / n, host_ms), flush=True)


def resident(i):
    model.train_step(xd[i % 4], yd[i % 4], 1e-4, 1.0)


def resident_one(i):
    model.train_step(x_in, y_in, 1e-4, 1.0)


def resident_slot(i):
    l = model.train_step(xd[i % 4], yd[i % 4], 1e-4, 1.0)
    slots[i % 2].fetch(l)
    if i:
        slots[(i - 1) % 2].value()


def h2d(stream):
    def body(i):
        _lib.call("nlps_memcpy_h2d", _lib.vp(x_in.ptr), pin[i % 4], C.c_size_t(nbytes), stream)
        _lib.call("nlps_memcpy_h2d", _lib.vp(y_in.ptr), pin[4 + i % 4], C.c_size_t(nbytes), stream)
        l = model.train_step(x_in, y_in, 1e-4, 1.0)
        slots[i % 2].fetch(l)
        if i:
            slots[(i - 1) % 2].value()
    return body


def h2d_sync_loss(i):
    _lib.call("nlps_memcpy_h2d", _lib.vp(x_in.ptr), pin[i % 4], C.c_size_t(nbytes), _lib.stream_ptr())
    _lib.call("nlps_memcpy_h2d", _lib.vp(y_in.ptr), pin[4 + i % 4], C.c_size_t(nbytes), _lib.stream_ptr())
    float(model.train_step(x_in, y_in, 1e-4, 1.0))


run("device-resident, 4 batches", 20, resident)
run("device-resident, one buffer", 20, resident_one)
run("resident + pipelined loss read", 20, resident_slot)
run("H2D on stream 0 + pipelined read", 20, h2d(_lib.stream_ptr()))
run("H2D on model stream + pipelined", 20, h2d(mstream))
run("H2D on stream 0 + float(loss)", 20, h2d_sync_loss)
run("device-resident, 4 batches (again)", 20, resident)
