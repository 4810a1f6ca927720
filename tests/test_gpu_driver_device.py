"""the evaluation pass, the learning-rate schedule and the top-k sampling of the driver loop
(utils/train.py:51-75, 86-88, 143-163) on the device, against the oracle running the reference's arithmetic."""
import numpy as np
import pytest

from oracle.transformer_oracle import OracleTransformer

pytestmark = pytest.mark.gpu

from nlpscratch_b200 import Transformer  # noqa: E402
from nlpscratch_b200.array import DeviceArray  # noqa: E402


def make(cfg, precision="fp32", seed=3):
    np.random.seed(seed)
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    rng = np.random.default_rng(seed)
    oracle.params["b_vocab"][...] = rng.uniform(-0.3, 0.3, oracle.params["b_vocab"].shape)
    m = Transformer(*cfg, dropout_p=0.1, precision=precision)      # dropout on: the eval pass must switch it off itself
    m.load_params(oracle.params)
    return oracle, m, rng


def prefix_rows(rng, n, max_len, V):
    X = np.zeros((n, max_len), np.int32)
    for i in range(n):
        L = int(rng.integers(1, max_len + 1))
        X[i, :L] = rng.integers(1, V, L)
    return X, rng.integers(0, V, n).astype(np.int32)


def oracle_eval(oracle, X, Y, bs):
    tot = 0.0
    oracle.eval()
    for i in range(0, len(Y), bs):
        xb, yb = X[i:i + bs], Y[i:i + bs]
        lens = (xb != 0).sum(1)
        H, _ = oracle.forward(xb[:, :lens.max()], return_hidden_only=True)
        lg = H[np.arange(len(yb)), lens - 1] @ oracle.params["W_vocab"] + oracle.params["b_vocab"]
        tot += float(oracle.calculate_loss_dlogits(lg, yb)[0]) * len(yb)
    oracle.train()
    return tot / len(Y)


def test_device_side_eval_pass_and_lr_halving():
    V, max_len = 97, 20
    oracle, m, rng = make((V, 32, 4, 64, 2, max_len))
    X, Y = prefix_rows(rng, 53, max_len, V)                      # ragged last batch (53 = 3*16 + 5)
    accum = DeviceArray.from_host(np.zeros(2, np.float32))
    sched = DeviceArray.from_host(np.array([0, 0, 1e-3, 0, 0], np.float32))

    def device_pass():
        for i in range(0, 53, 16):
            xb = X[i:i + 16]
            lens = (xb != 0).sum(1)
            m.eval_batch_last_token(xb[:, :lens.max()], Y[i:i + 16], accum, lens=lens if i else None)
        return m.eval_finish(accum, sched)
    want = oracle_eval(oracle, X, Y, 16)
    loss, lr, halved = device_pass()
    assert abs(loss - want) < 1e-5 * want and not halved and abs(lr - 1e-3) < 1e-10
    assert m.training                                            # the pass restored train mode
    loss2, lr2, halved2 = device_pass()                          # same loss again: not worse, no halving
    assert loss2 == loss and not halved2 and lr2 == lr
    # make the model worse: the next pass must halve the learning rate on the device (utils/train.py:86-88)
    bad = {k: v.copy() for k, v in oracle.params.items()}
    bad["b_vocab"] = bad["b_vocab"] + np.where(np.arange(V) % 2 == 0, 3.0, -3.0).astype(np.float32)
    m.load_params(bad)
    oracle.load_params(bad)
    want3 = oracle_eval(oracle, X, Y, 16)
    loss3, lr3, halved3 = device_pass()
    assert want3 > want and abs(loss3 - want3) < 1e-5 * want3
    assert halved3 and abs(lr3 - 0.5 * lr) < 1e-12
    assert np.array_equal(accum.get(), np.zeros(2, np.float32))  # the accumulator was cleared for the next pass
    m.close()


@pytest.mark.parametrize("V,k", [(97, 5), (574, 3), (8192, 8)])
def test_topk_sampling_matches_numpy(V, k):
    oracle, m, rng = make((V, 32, 4, 64, 1, 12))
    prompt = [2] + [int(v) for v in rng.integers(3, V, 6)]
    idx, prob = m.topk_next(prompt, k)
    H, _ = oracle.forward(np.array(prompt, np.int32)[None, :], return_hidden_only=True)
    nl = H[0, -1] @ oracle.params["W_vocab"] + oracle.params["b_vocab"]
    p = np.exp(nl - nl.max()); p /= p.sum()                       # utils/train.py:149-151
    want = np.array(p).argsort()[-k:][::-1]
    assert np.array_equal(idx, want), (idx, want)
    assert np.allclose(prob, p[want], rtol=1e-4, atol=1e-7)
    m.close()


def test_pipelined_loss_read_back_returns_every_steps_loss():
    """HostScalarSlot: step i's loss is copied to a pinned slot behind the step and read while step i+1 runs; the values
    are the ones a synchronous float(loss) returns (same steps, fresh model)."""
    from nlpscratch_b200 import HostScalarSlot
    cfg = (97, 32, 4, 64, 2, 20)
    rng = np.random.default_rng(0)
    x = rng.integers(1, 97, (4, 20)).astype(np.int32)
    y = rng.integers(0, 97, (4, 20)).astype(np.int32)
    got, want = [], []
    for mode in ("pipelined", "sync"):
        np.random.seed(5)
        m = Transformer(*cfg, dropout_p=0.1, precision="fp32", seed=9)
        if mode == "sync":
            want = [float(m.train_step(x, y, 1e-3, 1.0)) for _ in range(6)]
        else:
            slots = [HostScalarSlot(m), HostScalarSlot(m)]
            for i in range(6):
                slots[i % 2].fetch(m.train_step(x, y, 1e-3, 1.0))
                if i:
                    got.append(slots[(i - 1) % 2].value())
            got.append(slots[5 % 2].value())
            m.synchronize()
        m.close()
    assert got == want
