"""Print the measured tcgen05 kind::tf32 issue ceiling (csrc/tf32_peak.cu) for cta_group 1 and 2."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlpscratch_b200 import _lib  # noqa: E402


def measure(cta_group, iters=4000, reps=5):
    tf, ms = C.c_float(), C.c_float()
    _lib.call("nlps_tf32_mma_peak", C.c_int(0), C.c_int(cta_group), C.c_int(iters), C.c_int(reps), C.byref(tf), C.byref(ms))
    return {"cta_group": cta_group, "tflops": tf.value, "ms": ms.value, "iters": iters}


if __name__ == "__main__":
    out = [measure(g, it) for g in (1, 2) for it in (1000, 4000, 16000)]
    print(json.dumps(out))
