"""Dropout parity (Transformer._dropout, transformer.py:102-108, backward :377-379 / :416-418).

The reference draws its keep-masks from NumPy's global generator, which a GPU kernel cannot replay; the CUDA path draws
them from Philox4x32-10 keyed by (seed; layer, branch, forward-call counter) inside the GEMM epilogue and REGENERATES them
in the LayerNorm backward instead of storing them.  What must hold -- and is checked here -- is that (1) the forward applies
exactly the mask `nlps_dropout_mask` exports, scaled by 1/keep; (2) the backward applies the same mask; (3) with those masks
injected into the CPU oracle (`OracleTransformer.mask_source`) loss, gradients and updated weights agree to the FP32-faithful
tolerance 1e-4 (and to the TF32 tolerance on the tensor-core path the benchmark times); (4) the keep rate is keep +- 3 sigma.
"""
import numpy as np
import pytest

from helpers import NAMES, rel_err
from oracle.transformer_oracle import OracleTransformer, global_norm_clip

pytestmark = pytest.mark.gpu

from nlpscratch_b200 import Transformer  # noqa: E402

TOL = {"fp32": dict(loss=1e-5, grad=1e-4, param=1e-4), "tf32x3": dict(loss=1e-5, grad=1e-4, param=1e-4),
       "tf32": dict(loss=2e-3, grad=5e-2, param=None)}


def make_pair(cfg, p, precision, seed=11):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    oracle = OracleTransformer(*cfg, dropout_p=p)
    for n in ("b1", "b2", "b_vocab", "ln1_beta", "ln2_beta"):
        oracle.params[n][...] = rng.uniform(-0.1, 0.1, oracle.params[n].shape)
    m = Transformer(*cfg, dropout_p=p, precision=precision, seed=0xC0FFEE + seed)
    m.load_params(oracle.params)
    return oracle, m, rng


@pytest.mark.parametrize("precision", ["fp32", "tf32x3", "tf32"])
@pytest.mark.parametrize("p", [0.1, 0.25])
def test_full_sequence_step_with_injected_masks_matches_oracle(p, precision):
    cfg = (211, 128, 4, 256, 2, 48)                       # head_dim 32 -> tcgen05 attention in tf32 mode
    oracle, m, rng = make_pair(cfg, p, precision)
    x = rng.integers(1, cfg[0], (6, 48)).astype(np.int32)
    x[2, 30:] = 0
    y = rng.integers(0, cfg[0], (6, 48)).astype(np.int32)
    tol = TOL[precision]
    for s in range(3):
        step = m.dropout_counter                          # the Philox step this forward draws with
        oracle.mask_source = lambda layer, which, shape, _s=step: m.dropout_mask(layer, which, _s, shape)
        logits, cache = m.forward(x)
        loss, dl = m.calculate_loss_dlogits(logits, y)
        grads = m.backward(cache, dl)
        lo, co = oracle.forward(x)
        want, dlo = oracle.calculate_loss_dlogits(lo, y)
        go = oracle.backward(co, dlo)
        assert m.dropout_counter == step + 1
        assert abs(float(loss) - float(want)) < tol["loss"] * max(1.0, abs(float(want))), (s, float(loss), float(want))
        if s == 0 or tol["param"] is not None:
            # (TF32: from the second step on the weights differ by Adam's +-lr sign flips of near-zero gradients, so the
            #  gradients are compared only while the weights are still identical; loss and weight drift on every step)
            for k in NAMES:
                assert rel_err(grads[k].get(), go[k]) < tol["grad"], (s, k, rel_err(grads[k].get(), go[k]))
        m.step(grads, 1e-3)
        oracle.step(go, 1e-3)
    got = m.get_params()
    for k in NAMES:
        if tol["param"] is not None:
            assert rel_err(got[k], oracle.params[k]) < tol["param"], k
        else:
            assert float(np.max(np.abs(got[k].astype(np.float64) - oracle.params[k]))) <= 2.5 * 1e-3 * 3, k
    m.close()


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_fused_steps_with_injected_masks_match_oracle(precision):
    """The fused entry points (nlps_train_step: eager, captured, replayed; nlps_train_step_last_token) draw the same
    masks as the unfused API: 4 full-sequence steps (dropout 0.1, clip, Adam) then 2 last-token driver steps."""
    cfg = (211, 128, 4, 256, 2, 48)
    p = 0.1
    oracle, m, rng = make_pair(cfg, p, precision, seed=13)
    x = rng.integers(1, cfg[0], (6, 48)).astype(np.int32)
    y = rng.integers(0, cfg[0], (6, 48)).astype(np.int32)
    tol = TOL[precision]
    for s in range(4):
        step = m.dropout_counter
        oracle.mask_source = lambda layer, which, shape, _s=step: m.dropout_mask(layer, which, _s, shape)
        loss = m.train_step(x, y, learning_rate=1e-3, clip_grad_norm=1.0)
        lo, co = oracle.forward(x)
        want, dlo = oracle.calculate_loss_dlogits(lo, y)
        go = oracle.backward(co, dlo)
        global_norm_clip(list(go.values()), 1.0)
        oracle.step(go, 1e-3)
        assert abs(float(loss) - float(want)) < 10 * tol["loss"] * max(1.0, abs(float(want))), (s, float(loss), float(want))
    xl = np.zeros((6, 48), np.int32)
    lens = np.array([9, 48, 1, 17, 30, 5])
    for b, n in enumerate(lens):
        xl[b, :n] = rng.integers(1, cfg[0], n)
    yl = rng.integers(0, cfg[0], (6,)).astype(np.int32)
    for s in range(2):
        step = m.dropout_counter
        oracle.mask_source = lambda layer, which, shape, _s=step: m.dropout_mask(layer, which, _s, shape)
        loss = m.train_step_last_token(xl[:, :lens.max()], yl, lens=lens, learning_rate=1e-3, clip_grad_norm=1.0)
        H, co = oracle.forward(xl[:, :lens.max()], return_hidden_only=True)
        lg = H[np.arange(6), lens - 1] @ oracle.params["W_vocab"] + oracle.params["b_vocab"]
        want, dlo = oracle.calculate_loss_dlogits(lg, yl)
        go = oracle.backward_last_token(co, lens - 1, dlo[:, 0, :])
        global_norm_clip(list(go.values()), 1.0)
        oracle.step(go, 1e-3)
        assert abs(float(loss) - float(want)) < 10 * tol["loss"] * max(1.0, abs(float(want))), (s, float(loss), float(want))
    got = m.get_params()
    for k in NAMES:
        if tol["param"] is not None:
            assert rel_err(got[k], oracle.params[k]) < tol["param"], k
        else:
            assert float(np.max(np.abs(got[k].astype(np.float64) - oracle.params[k]))) <= 2.5 * 1e-3 * 6, k
    m.close()


@pytest.mark.parametrize("p", [0.1, 0.25, 0.5])
def test_keep_rate_scale_and_forward_backward_mask_equality(p):
    """One layer, weights chosen so that the attention branch's dropped tensor is visible from outside:
    the mask the forward applied (recovered from r1 = x + dropout(ctx Wo)) equals the exported one, the kept entries carry
    exactly 1/keep, the backward zeroes exactly the same entries, and the keep rate is binomial around keep."""
    keep = 1.0 - p
    E, S, B = 64, 40, 8
    cfg = (50, E, 2, 128, 1, S)
    np.random.seed(3)
    m = Transformer(*cfg, dropout_p=p, precision="fp32", seed=77)
    shape = (B, S, E)
    n = B * S * E
    rates = []
    for step in (0, 1, 5):
        for which in (0, 1):
            mask = m.dropout_mask(0, which, step, shape)
            assert set(np.unique(mask)) <= {0.0, 1.0}
            rates.append(mask.mean())
            sigma = np.sqrt(keep * p / n)
            # the threshold is round(keep * 2^16) / 2^16: include its quantisation in the bound
            assert abs(mask.mean() - keep) < 3 * sigma + 2.0 ** -16, (p, step, which, mask.mean())
    assert not np.array_equal(m.dropout_mask(0, 0, 0, shape), m.dropout_mask(0, 1, 0, shape))     # sites are independent
    assert not np.array_equal(m.dropout_mask(0, 0, 0, shape), m.dropout_mask(0, 0, 1, shape))     # fresh masks per forward
    # forward: compare train-mode and eval-mode pre-LayerNorm residuals through the oracle's algebra
    oracle = OracleTransformer(*cfg, dropout_p=p)
    oracle.load_params(m.get_params())
    x = np.random.default_rng(0).integers(1, 50, (B, S)).astype(np.int32)
    y = np.random.default_rng(1).integers(0, 50, (B, S)).astype(np.int32)
    step = m.dropout_counter
    drawn = {}

    def source(layer, which, shp):
        drawn[(layer, which)] = m.dropout_mask(layer, which, step, shp)
        return drawn[(layer, which)]
    oracle.mask_source = source
    logits, cache = m.forward(x)
    lo, co = oracle.forward(x)
    assert rel_err(logits.get(), lo) < 1e-4                 # forward applied exactly these masks with scale 1/keep
    # a wrong scale (e.g. no 1/keep) or a different mask would be an O(1) error:
    oracle.mask_source = lambda layer, which, shp: np.roll(drawn[(layer, which)], 1, axis=-1)
    lo_wrong, _ = oracle.forward(x)
    assert rel_err(logits.get(), lo_wrong) > 1e-2
    # backward regenerates the masks: gradients agree with the oracle that stored them
    oracle.mask_source = source
    loss, dl = m.calculate_loss_dlogits(logits, y)
    grads = m.backward(cache, dl)
    want, dlo = oracle.calculate_loss_dlogits(lo, y)
    go = oracle.backward(co, dlo)
    for k in NAMES:
        assert rel_err(grads[k].get(), go[k]) < 1e-4, k
    m.close()


def test_resume_checkpoint_continues_the_dropout_stream(tmp_path):
    """save_checkpoint stores the Philox seed and forward-call counter, so a restored model with dropout on
    continues bit for bit (before, it replayed the masks of steps 0..k)."""
    import os
    cfg = (97, 64, 2, 128, 2, 24)
    rng = np.random.default_rng(0)
    x = rng.integers(1, 97, (4, 24)).astype(np.int32)
    y = rng.integers(0, 97, (4, 24)).astype(np.int32)
    np.random.seed(1)
    a = Transformer(*cfg, dropout_p=0.1, precision="fp32", seed=1234)
    for _ in range(3):
        a.train_step(x, y, 1e-3, 1.0)
    ck = os.path.join(tmp_path, "ck.pkl")
    a.save_checkpoint(ck)
    for _ in range(2):
        a.train_step(x, y, 1e-3, 1.0)
    want = a.get_params()
    np.random.seed(2)
    b = Transformer(*cfg, dropout_p=0.1, precision="fp32", seed=999)       # other seed: the checkpoint's must win
    b.load_checkpoint(ck)
    assert b.dropout_counter == 3 and b.adam_t == 3
    for _ in range(2):
        b.train_step(x, y, 1e-3, 1.0)
    got = b.get_params()
    for k in want:
        assert np.array_equal(want[k], got[k]), k
    a.close(); b.close()
