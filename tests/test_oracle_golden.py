"""Pins oracle/transformer_oracle.py to fixtures produced by the real reference."""
import hashlib

import numpy as np
import pytest

from oracle.transformer_oracle import (OracleTransformer, global_norm_clip, last_token_step,
                                        sinusoid_table, flops_per_token)
from helpers import FULL_SEQ_CASES, NAMES, load_case, manifest, params_from_case, rel_err

TOL = 1e-6      # layer-0 Q,K,V and scores are f32 GEMMs in both; only their summation order differs


def build(case, use_adam=True):
    cfg = [int(v) for v in case["cfg"]]
    m = OracleTransformer(*cfg, dropout_p=0.0, use_adam=use_adam)
    m.load_params(params_from_case(case, 0))
    return m


@pytest.mark.parametrize("name", FULL_SEQ_CASES)
def test_full_sequence_cases(name):
    case = load_case(name)
    use_adam = bool(case["use_adam"])
    m = build(case, use_adam)
    lr = float(case["lr"])
    steps = sum(1 for k in case if k.startswith("loss"))
    for s in range(steps):
        logits, cache = m.forward(case["x"])
        loss, dlogits = m.calculate_loss_dlogits(logits, case["y"])
        grads = m.backward(cache, dlogits)
        assert abs(float(loss) - float(case["loss%d" % s])) < 1e-7
        if s == 0:
            assert rel_err(logits, case["logits"]) < TOL
            assert rel_err(dlogits, case["dlogits"]) < TOL
            assert rel_err(cache["H_final"], case["H_final"]) < TOL
            assert rel_err(cache["layers"][0]["prob"], case["attn0"]) < TOL
            for k in NAMES:
                assert grads[k].dtype == np.float32
                assert rel_err(grads[k], case["g_" + k]) < 2e-6, k    # f32 accumulation buffers
        m.step(grads, lr)
        for k in NAMES:
            assert rel_err(m.params[k], case["p%d_%s" % (s + 1, k)]) < 1e-6, k


def test_leading_pad_rows_attend_to_all_keys():
    case = load_case("full_seq_leading_pad")
    m = build(case)
    _, cache = m.forward(case["x"])
    prob = cache["layers"][0]["prob"]
    # sequence 1 starts with three PADs: query rows 0..2 see every key, incl. future ones
    assert np.all(prob[1, :, 0, :] > 0)
    assert np.allclose(prob[1, :, 0, :].sum(-1), 1.0)
    # an ordinary row keeps exact zeros on masked keys
    assert np.all(prob[2, :, 2, 3:] == 0)


def test_last_token_path_with_clip():
    case = load_case("last_token_clip")
    m = build(case)
    lr, clip = float(case["lr"]), float(case["clip"])
    for s in range(2):
        before = {k: v.copy() for k, v in m.params.items()}
        loss, grads = last_token_step(m, case["x"], case["y"], lr=lr, clip=clip)
        assert abs(loss - float(case["loss%d" % s])) < 1e-7
        if s == 0:
            for k in NAMES:
                assert rel_err(grads[k], case["g_" + k]) < 2e-6, k
            # PAD row of the embedding gets exactly zero gradient in last-token mode
            assert np.all(grads["We"][0] == 0)
        for k in NAMES:
            assert rel_err(m.params[k], case["p%d_%s" % (s + 1, k)]) < 1e-6, k
        assert any(not np.array_equal(before[k], m.params[k]) for k in NAMES)


def test_reference_main_smoke_known_answers():
    ka = manifest()["smoke_main"]
    x = np.array([[1, 2, 3, 4, 5], [5, 4, 3, 2, 1]], dtype=np.int32)
    y = np.array([[2, 3, 4, 5, 6], [1, 1, 1, 1, 1]], dtype=np.int32)
    for p in (0.1, 0.0):
        np.random.seed(42)
        m = OracleTransformer(1000, 64, 8, 256, 2, 50, dropout_p=p)   # same draw order as the reference
        if p == 0.0:
            for k, sha in ka["param_sha256_16"].items():
                assert hashlib.sha256(np.ascontiguousarray(m.params[k]).tobytes()).hexdigest()[:16] == sha, k
        logits, cache = m.forward(x)
        loss, dlogits = m.calculate_loss_dlogits(logits, y)
        assert abs(float(loss) - ka["loss_p%g" % p]) < 1e-8
        if p == 0.0:
            assert np.allclose(logits[0, 0, :3], ka["logits_0_0_3"], rtol=0, atol=1e-7)
            grads = m.backward(cache, dlogits)
            for k, nrm in ka["grad_norms"].items():
                got = float(np.sqrt(np.sum(np.square(grads[k].astype(np.float64)))))
                assert abs(got - nrm) < 1e-6 * max(1.0, nrm), k
    assert np.allclose(sinusoid_table(50, 64)[1, :4], ka["pe_1_4"], atol=0)
    assert abs(float(loss) - 6.861496003409776) < 1e-8          # SURVEY section 4 known answer


def test_clip_is_joint_and_in_place():
    g = [np.full((4,), 3.0, np.float32), np.full((9,), 4.0, np.float32)]
    total = float(global_norm_clip(g, 1.0))
    assert abs(total - np.sqrt(4 * 9 + 9 * 16)) < 1e-5
    assert abs(np.sqrt(sum(float(np.sum(a * a)) for a in g)) - 1.0) < 1e-5
    g2 = [np.full((2,), 0.1, np.float32)]
    global_norm_clip(g2, 1.0)
    assert np.all(g2[0] == np.float32(0.1))


def test_flops_formula_matches_survey():
    assert abs(flops_per_token(6, 512, 2048, 512, 8192) / 1e6 - 147.85) < 0.01
    assert abs(flops_per_token(12, 768, 3072, 2048, 8192) / 1e6 - 660.60) < 0.01
