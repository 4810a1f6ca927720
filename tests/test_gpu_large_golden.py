"""The benchmarked configurations against fixtures of the REAL reference (tests/golden/make_golden_large.py), through the
public API: reference-default model (full-sequence and the driver's last-token path), c3 (vocabulary 32768), the c4 model
(BASELINE.json configs[3]; B=2, S=512) and the c5 model (configs[4]; B=1, S=2048) -- in all three arithmetic modes.

In `tf32` mode these shapes run the kernels the benchmark times (asserted through nlps_kernel_stats): the CTA-pair
tcgen05 GEMM and the persistent tcgen05 attention.  Tolerances (norm-wise relative on deterministic samples, and on f64 norms):
  fp32, tf32x3 : 1e-4 logits / hidden / loss / every gradient tensor; updated weights 1e-4
  tf32         : 5e-3 logits / hidden, 2e-3 loss, 5e-2 gradients (10-bit operand mantissas: the reference's own cuBLAS TF32 mode),
                 updated weights within 2.5*lr*steps (Adam's first steps are +-lr*sign(g))
Observed errors are written to gpurun_out/parity_large.json when that directory exists.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, LARGE_CASES, NAMES, load_case, perturb_params, rel_err, sample_idx, sampled, sha16

pytestmark = pytest.mark.gpu

from nlpscratch_b200 import Transformer, _lib  # noqa: E402

TOL = {"fp32": dict(act=1e-4, grad=1e-4, loss=1e-5, param=1e-4),
       "tf32x3": dict(act=1e-4, grad=1e-4, loss=1e-5, param=1e-4),
       "tf32": dict(act=5e-3, grad=5e-2, loss=2e-3, param=None)}
REPORT = {}


def kernel_stats(reset=False):
    out = (C.c_int64 * 9)()
    _lib.call("nlps_kernel_stats", out, C.c_int(9), C.c_int(int(reset)))
    return dict(zip(("gemm_pair", "gemm_tc_single", "gemm_simt", "gemm_x3", "attn_ts_fwd", "attn_ts_bwd", "attn_mma",
                     "attn_fp32", "attn_ss"), list(out)))


def norm64(a):
    return float(np.sqrt(np.sum(np.square(np.asarray(a, dtype=np.float64)))))


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("name", LARGE_CASES)
def test_benchmarked_sizes_against_reference_fixtures(name, precision):
    case = load_case(name)
    cfg = [int(v) for v in case["cfg"]]
    seed = int(case["seed"])
    with open(os.path.join(GOLDEN, "manifest_large.json")) as fh:
        shas = json.load(fh)["cases"][name]["p0_sha"]
    np.random.seed(seed)                                   # the reference's draw order => the reference's weights
    m = Transformer(*cfg, dropout_p=0.0, precision=precision)
    params = perturb_params(m.get_params(), seed + 1)
    m.load_params(params)
    for k in NAMES:
        assert sha16(m.get_params()[k]) == shas[k], k
    tol = TOL[precision]
    lr, clip, mode = float(case["lr"]), float(case["clip"]), str(case["mode"])
    x, y = case["x"], case["y"]
    steps = sum(1 for k in case if k.startswith("loss"))
    rep = {}
    kernel_stats(reset=True)
    for s in range(steps):
        if mode == "full":
            logits, cache = m.forward(x)
            loss, dl = m.calculate_loss_dlogits(logits, y)
            grads = m.backward(cache, dl)
            H = cache["H_final"]
        else:                                              # utils/train.py:104-131 on the model.np surface
            npb = m.np
            xb = npb.array(x)
            lens = npb.sum(xb != 0, axis=1).astype(npb.int32)
            xb = xb[:, :int(npb.max(lens))]
            H, cache = m.forward(xb, return_hidden_only=True)
            t = lens - 1
            logits = H[npb.arange(len(y)), t, :] @ m.W_vocab + m.b_vocab
            loss, dl = m.calculate_loss_dlogits(logits, npb.array(y))
            grads = m.backward_last_token(cache, t, dl[:, 0, :])
        want = float(case["loss%d" % s])
        rep["loss%d" % s] = abs(float(loss) - want) / max(1.0, abs(want))
        assert rep["loss%d" % s] < tol["loss"], (s, float(loss), want)
        if s == 0:
            rep["logits"] = rel_err(sampled(logits.get()), case["logits_s"])
            rep["hidden"] = rel_err(sampled(H.get()), case["H_s"])
            assert rep["logits"] < tol["act"] and rep["hidden"] < tol["act"], rep
            ln = float(case["logits_norm"])
            assert abs(norm64(logits.get()) - ln) < tol["act"] * ln
            for k in NAMES:
                g = grads[k].get()
                assert g.dtype == np.float32 and g.shape == params[k].shape
                e = rel_err(sampled(g), case["graw_s_" + k])
                gn = float(case["graw_norm_" + k])
                rep["grad_" + k] = e
                assert e < tol["grad"], (k, e)
                assert abs(norm64(g) - gn) < tol["grad"] * max(gn, 1e-12), (k, norm64(g), gn)
        if clip > 0:
            m.clip_gradients(clip)                          # clip_grads(list(grads.values()), clip), function.py:27-38
        m.step(grads, lr)
        got = m.get_params()
        for k in NAMES:
            w = case["p%d_s_%s" % (s + 1, k)]
            if tol["param"] is not None:
                assert rel_err(sampled(got[k]), w) < tol["param"], (k, s)
            else:
                assert float(np.max(np.abs(sampled(got[k]).astype(np.float64) - w))) <= 2.5 * lr * (s + 1), (k, s)
    stats = kernel_stats()
    rep["kernels"] = stats
    if precision == "tf32":
        # the kernels the benchmark times produced these numbers
        assert stats["gemm_pair"] > 0 and stats["gemm_simt"] == 0, stats
        assert stats["attn_ts_fwd"] > 0 and stats["attn_ts_bwd"] > 0 and stats["attn_fp32"] == 0 and stats["attn_mma"] == 0, stats
    elif precision == "tf32x3":
        assert stats["gemm_x3"] > 0, stats
    else:
        assert stats["gemm_pair"] == 0 and stats["gemm_tc_single"] == 0 and stats["attn_ts_fwd"] == 0, stats
    m.close()
    REPORT["%s/%s" % (name, precision)] = rep
    out_dir = os.path.join(os.path.dirname(GOLDEN), os.pardir, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_large.json"), "w") as fh:
            json.dump(REPORT, fh, indent=1, sort_keys=True)
