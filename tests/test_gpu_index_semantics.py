"""NumPy index semantics the reference relies on, for ids that only exist on the device:
negative targets wrap (transformer.py:538-546 ``log_probs[rows, y]``), out-of-range ids become IndexError at the next
synchronisation point, ``H[idx, t, :]`` wraps a negative t on its own axis (utils/train.py:113-115)."""
import numpy as np
import pytest

from helpers import rel_err
from oracle.transformer_oracle import OracleTransformer

pytestmark = pytest.mark.gpu

from nlpscratch_b200 import Transformer, b200np  # noqa: E402
from nlpscratch_b200.array import DeviceArray  # noqa: E402


def unbounded(a):
    d = DeviceArray.from_host(a)
    d.bounds = None                       # as if built on the device (data.py, gathers): no host-side range knowledge
    return d


@pytest.mark.parametrize("V", [64, 61])   # vectorised and scalar cross-entropy kernels
def test_negative_targets_wrap_like_numpy(V):
    cfg = (V, 32, 2, 64, 1, 8)
    np.random.seed(0)
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    m = Transformer(*cfg, dropout_p=0.0, precision="fp32")
    m.load_params(oracle.params)
    rng = np.random.default_rng(0)
    x = rng.integers(1, V, (3, 8)).astype(np.int32)
    y = rng.integers(0, V, (3, 8)).astype(np.int32)
    y[0, 0], y[1, 3], y[2, 7] = -1, -V, -5
    for yd in (y, unbounded(y)):          # host-known range and device-only ids
        logits, cache = m.forward(x)
        loss, dl = m.calculate_loss_dlogits(logits, yd)
        lo, co = oracle.forward(x)
        want, dlo = oracle.calculate_loss_dlogits(lo, y)
        assert abs(float(loss) - float(want)) < 1e-5
        assert rel_err(dl.get(), dlo) < 1e-5
    # the fused step takes the same targets
    l2 = float(m.train_step(x, unbounded(y), defer_update=True))
    assert abs(l2 - float(want)) < 1e-5
    m.close()


def test_out_of_range_ids_on_the_device_raise_index_error_at_the_next_sync():
    V = 40
    m = Transformer(V, 32, 2, 64, 1, 8, dropout_p=0.0, precision="fp32")
    x = np.random.default_rng(0).integers(1, V, (2, 8)).astype(np.int32)
    y = np.random.default_rng(1).integers(0, V, (2, 8)).astype(np.int32)
    bad_y = y.copy(); bad_y[1, 2] = V
    logits, _ = m.forward(x)
    loss, _ = m.calculate_loss_dlogits(logits, unbounded(bad_y))
    with pytest.raises(IndexError):
        float(loss)
    float(m.calculate_loss_dlogits(logits, y)[0])          # the flag was cleared: a clean call passes
    bad_x = x.copy(); bad_x[0, 0] = V + 3
    with pytest.raises(IndexError):                        # host array: caught before any launch, like NumPy
        m.forward(bad_x)
    m.forward(unbounded(bad_x))
    with pytest.raises(IndexError):
        m.synchronize()
    m.synchronize()
    with pytest.raises(IndexError):
        float(m.train_step(x, unbounded(bad_y)))
    m.close()


def test_fancy_index_wraps_each_axis_like_numpy():
    rng = np.random.default_rng(0)
    H = rng.standard_normal((5, 7, 12)).astype(np.float32)
    idx = np.arange(5)
    t = np.array([3, -1, 0, -7, 6])
    got = b200np.array(H)[unbounded(idx.astype(np.int32)), unbounded(t.astype(np.int32)), :].get()
    assert np.array_equal(got, H[idx, t, :])
    got2 = b200np.array(H)[b200np.arange(5), b200np.array(t.astype(np.int32)), :].get()
    assert np.array_equal(got2, H[idx, t, :])
    with pytest.raises(IndexError):
        b200np.array(H)[b200np.arange(5), b200np.array(np.array([0, 0, 0, 0, 7], np.int32)), :]


def test_driver_path_with_an_all_pad_row_matches_numpy_semantics():
    """lens == 0 for an all-PAD row makes t = -1: NumPy (and the reference) reads that row's LAST position in the forward
    gather and scatters the gradient there; gather and scatter must agree."""
    cfg = (30, 32, 2, 64, 1, 10)
    np.random.seed(2)
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    m = Transformer(*cfg, dropout_p=0.0, precision="fp32")
    m.load_params(oracle.params)
    rng = np.random.default_rng(5)
    x = np.zeros((4, 10), np.int32)
    for b, n in enumerate((6, 0, 10, 3)):                  # row 1 is all PAD
        x[b, :n] = rng.integers(1, 30, n)
    y = rng.integers(0, 30, (4,)).astype(np.int32)
    npb = m.np
    xb = npb.array(x)
    lens = npb.sum(xb != 0, axis=1).astype(npb.int32)
    xb = xb[:, :int(npb.max(lens))]
    H, cache = m.forward(xb, return_hidden_only=True)
    t = lens - 1
    logits_last = H[npb.arange(4), t, :] @ m.W_vocab + m.b_vocab
    loss, dl = m.calculate_loss_dlogits(logits_last, npb.array(y))
    grads = m.backward_last_token(cache, t, dl[:, 0, :])
    lens_h = (x != 0).sum(1)
    Ho, co = oracle.forward(x[:, :lens_h.max()], return_hidden_only=True)
    lo = Ho[np.arange(4), lens_h - 1] @ oracle.params["W_vocab"] + oracle.params["b_vocab"]
    want, dlo = oracle.calculate_loss_dlogits(lo, y)
    go = oracle.backward_last_token(co, lens_h - 1, dlo[:, 0, :])
    assert abs(float(loss) - float(want)) < 1e-5
    for k in go:
        assert rel_err(grads[k].get(), go[k]) < 1e-4, k
    m.close()
