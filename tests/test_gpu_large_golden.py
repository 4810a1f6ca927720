"""The benchmarked configurations against fixtures of the REAL reference (tests/golden/make_golden_large.py), through the
public API: reference-default model (full-sequence and the driver's last-token path), c3 (vocabulary 32768), the c4 model
(BASELINE.json configs[3]; B=2, S=512) and the c5 model (configs[4]; B=1, S=2048) -- in all three arithmetic modes.

In `tf32` mode these shapes run the kernels the benchmark times (asserted through nlps_kernel_stats): the CTA-pair
tcgen05 GEMM and the persistent tcgen05 attention.  Tolerances (norm-wise relative on deterministic samples, and on f64 norms):
  fp32, tf32x3 : 1e-4 logits / hidden / loss / every gradient tensor; updated weights 1e-4
  tf32         : 2e-3 logits / hidden, 2e-4 loss, 4e-2 gradients, per-tensor gradient norms within the same bound
                 (10-bit operand mantissas: the reference's own cuBLAS TF32 mode; observed 4.8e-4 / 1.2e-5 / 2.0e-2),
                 updated weights within 2.5*lr*steps (Adam's first steps are +-lr*sign(g)); the same bound replaces the 1e-4 one
                 in the FP32-faithful modes when a near-zero ReLU gate flipped (next paragraph): the flip perturbs the
                 gradients below it by ~1e-4, which turns the Adam sign of a few near-zero gradient elements (2*lr each)
ReLU gates within float32 rounding of zero.  The gradient of the network is discontinuous where a hidden pre-activation
z = y1 W1 + b1 crosses 0 (function.py:53-56 zeroes dz where z <= 0).  The reference evaluates z in float64 (NumPy >= 2
promotion); any float32 implementation lands on the other side of 0 for a few of the 10^6..10^8 gates of these
models (|z| ~ 1e-7 sigma), and ONE flipped gate moves the bias gradient b1 by ~4e-4 -- measured: the fp32 CUDA path
differs from the reference at default_full_seq by exactly the gate at z = -1.07e-7 (layer 1, row 60, column 773), and
flipping that one gate inside the float64 oracle reproduces every per-tensor error to two digits.  The fixtures therefore
list the gates with |z| < 2e-5 sigma; when a gradient misses the tolerance the test reads our cached ffn_hidden at those
gates, and if (and only if) some of THEM differ from the reference's sign it re-evaluates the reference gradient with
those gates flipped (oracle.gate_flips; the oracle is pinned to the reference at these sizes by
tests/test_oracle_golden_large.py) and holds the CUDA path to the same tolerance against that.  A flip outside the list,
or any other error, still fails.  Flip counts are reported.
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
from oracle.transformer_oracle import OracleTransformer  # noqa: E402

TOL = {"fp32": dict(act=1e-4, grad=1e-4, loss=1e-5, param=1e-4),
       "tf32x3": dict(act=1e-4, grad=1e-4, loss=1e-5, param=1e-4),
       "tf32": dict(act=2e-3, grad=4e-2, loss=2e-4, param=None)}     # observed: 4.8e-4 / 2.0e-2 / 1.2e-5 (profiles/r02_parity_large.json)
REPORT = {}


def kernel_stats(reset=False):
    out = (C.c_int64 * 9)()
    _lib.call("nlps_kernel_stats", out, C.c_int(9), C.c_int(int(reset)))
    return dict(zip(("gemm_pair", "gemm_tc_single", "gemm_simt", "gemm_x3", "attn_ts_fwd", "attn_ts_bwd", "attn_mma",
                     "attn_fp32", "attn_ss"), list(out)))


def norm64(a):
    return float(np.sqrt(np.sum(np.square(np.asarray(a, dtype=np.float64)))))


def fragile_gate_flips(case, cache):
    """{layer: flat indices} of the listed near-zero gates where our float32 gate (ffn_hidden > 0) differs from the
    reference's float64 one (z > 0)."""
    flips = {}
    layers = cache["layers"]
    for l in np.unique(case["frag_layer"]):
        sel = case["frag_layer"] == l
        idx, z = case["frag_index"][sel], case["frag_z"][sel]
        ours = layers[int(l)]["ffn_hidden"].get().reshape(-1)[idx] > 0
        bad = idx[ours != (z > 0)]
        if len(bad):
            flips[int(l)] = bad
    return flips


def flipped_z(case, flips):
    worst = 0.0
    for l, idx in flips.items():
        sel = case["frag_layer"] == l
        lut = dict(zip(case["frag_index"][sel].tolist(), case["frag_z"][sel].tolist()))
        worst = max([worst] + [abs(lut[int(i)]) for i in idx])
    return worst


def oracle_grads_with_flips(case, cfg, seed, flips, mode):
    """The reference gradient (float64 oracle, pinned to the reference at this size) evaluated with the listed gates flipped."""
    np.random.seed(seed)
    o = OracleTransformer(*cfg, dropout_p=0.0)
    perturb_params(o.params, seed + 1)
    o.gate_flips = flips
    x, y = case["x"], case["y"]
    if mode == "full":
        lo, co = o.forward(x)
        _, dl = o.calculate_loss_dlogits(lo, y)
        g = o.backward(co, dl)
    else:
        lens = (x != 0).sum(1)
        H, co = o.forward(x[:, :lens.max()], return_hidden_only=True)
        lg = H[np.arange(len(y)), lens - 1] @ o.params["W_vocab"] + o.params["b_vocab"]
        _, dl = o.calculate_loss_dlogits(lg, y)
        g = o.backward_last_token(co, lens - 1, dl[:, 0, :])
    return {k: sampled(g[k]) for k in NAMES}, {k: norm64(g[k]) for k in NAMES}


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
    rep, failures = {}, []

    def check(ok, what):
        if not ok:
            failures.append(what)
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
        check(rep["loss%d" % s] < tol["loss"], ("loss", s, float(loss), want))
        if s == 0:
            rep["logits"] = rel_err(sampled(logits.get()), case["logits_s"])
            rep["hidden"] = rel_err(sampled(H.get()), case["H_s"])
            check(rep["logits"] < tol["act"] and rep["hidden"] < tol["act"], ("act", rep["logits"], rep["hidden"]))
            ln = float(case["logits_norm"])
            check(abs(norm64(logits.get()) - ln) < tol["act"] * ln, ("logits norm",))
            got_g = {k: grads[k].get() for k in NAMES}
            want_s = {k: case["graw_s_" + k] for k in NAMES}
            want_n = {k: float(case["graw_norm_" + k]) for k in NAMES}
            worst = max(rel_err(sampled(got_g[k]), want_s[k]) for k in NAMES)
            rep["grad_worst_vs_reference_gates"] = worst
            rep["gate_flips"] = 0
            if worst >= tol["grad"]:
                flips = fragile_gate_flips(case, cache)
                rep["gate_flips"] = int(sum(len(v) for v in flips.values()))
                rep["gate_flips_max_abs_z_over_sigma"] = flipped_z(case, flips)
                if rep["gate_flips"]:
                    want_s, want_n = oracle_grads_with_flips(case, cfg, seed, flips, mode)
            for k in NAMES:
                g = got_g[k]
                assert g.dtype == np.float32 and g.shape == params[k].shape
                e = rel_err(sampled(g), want_s[k])
                rep["grad_" + k] = e
                rep["gradnorm_" + k] = abs(norm64(g) - want_n[k]) / max(want_n[k], 1e-12)
                check(e < tol["grad"], ("grad", k, e))
                check(rep["gradnorm_" + k] < tol["grad"], ("grad norm", k, norm64(g), want_n[k]))
        if clip > 0:
            m.clip_gradients(clip)                          # clip_grads(list(grads.values()), clip), function.py:27-38
        m.step(grads, lr)
        got = m.get_params()
        for k in NAMES:
            w = case["p%d_s_%s" % (s + 1, k)]
            if tol["param"] is not None and not rep.get("gate_flips"):
                check(rel_err(sampled(got[k]), w) < tol["param"], ("param", k, s))
            else:
                check(float(np.max(np.abs(sampled(got[k]).astype(np.float64) - w))) <= 2.5 * lr * (s + 1), ("param", k, s))
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
    rep["failures"] = [str(f) for f in failures]
    REPORT["%s/%s" % (name, precision)] = rep
    out_dir = os.path.join(os.path.dirname(GOLDEN), os.pardir, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_large.json"), "w") as fh:
            json.dump(REPORT, fh, indent=1, sort_keys=True)
    assert not failures, failures
