"""Pins the oracle to the reference at the benchmarked model sizes (tests/golden/make_golden_large.py): the reference's
default hyper-parameters, the BPE-sized vocabulary (c3), the c4 model and the c5 model.  Parameters are seed-derived on
both sides (SHA-256 pinned), large tensors are compared through deterministic samples and f64 norms."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, NAMES, load_case, perturb_params, rel_err, sampled, sha16
from oracle.transformer_oracle import OracleTransformer, global_norm_clip

CPU_CASES = ("default_full_seq", "default_last_token", "c3_full_seq", "c4_full_seq")   # c5 costs ~40 s on CPU: GPU test only


def oracle_for(case):
    cfg = [int(v) for v in case["cfg"]]
    np.random.seed(int(case["seed"]))
    m = OracleTransformer(*cfg, dropout_p=0.0)
    perturb_params(m.params, int(case["seed"]) + 1)
    return m


def manifest_large():
    with open(os.path.join(GOLDEN, "manifest_large.json")) as fh:
        return json.load(fh)


def norm64(a):
    return float(np.sqrt(np.sum(np.square(np.asarray(a, dtype=np.float64)))))


@pytest.mark.parametrize("name", CPU_CASES)
def test_oracle_matches_reference_at_benchmarked_sizes(name):
    case = load_case(name)
    m = oracle_for(case)
    shas = manifest_large()["cases"][name]["p0_sha"]
    for k in NAMES:
        assert sha16(m.params[k]) == shas[k], k              # same weights as the reference, bit for bit
    lr, clip, mode = float(case["lr"]), float(case["clip"]), str(case["mode"])
    x, y = case["x"], case["y"]
    steps = sum(1 for k in case if k.startswith("loss"))
    for s in range(steps):
        if mode == "full":
            logits, cache = m.forward(x)
            loss, dl = m.calculate_loss_dlogits(logits, y)
            grads = m.backward(cache, dl)
            H = cache["H_final"]
        else:
            lens = (x != 0).sum(1)
            H, cache = m.forward(x[:, :lens.max()], return_hidden_only=True)
            t = lens - 1
            logits = H[np.arange(len(y)), t] @ m.params["W_vocab"] + m.params["b_vocab"]
            loss, dl = m.calculate_loss_dlogits(logits, y)
            grads = m.backward_last_token(cache, t, dl[:, 0, :])
        assert abs(float(loss) - float(case["loss%d" % s])) < 1e-6
        if s == 0:
            assert rel_err(sampled(logits), case["logits_s"]) < 1e-5
            assert rel_err(sampled(H), case["H_s"]) < 1e-5
            assert abs(norm64(logits) - float(case["logits_norm"])) < 1e-5 * float(case["logits_norm"])
            for k in NAMES:
                assert rel_err(sampled(grads[k]), case["graw_s_" + k]) < 2e-5, k
                assert abs(norm64(grads[k]) - float(case["graw_norm_" + k])) < 2e-5 * max(float(case["graw_norm_" + k]), 1e-12), k
        if clip > 0:
            global_norm_clip(list(grads.values()), clip)
        m.step(grads, lr)
        for k in NAMES:
            assert rel_err(sampled(m.params[k]), case["p%d_s_%s" % (s + 1, k)]) < 1e-6, k
