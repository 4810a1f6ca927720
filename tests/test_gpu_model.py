"""End-to-end parity of the drop-in Transformer against the golden fixtures of the real reference and
against the oracle on seeded inputs -- through the public API only (ctypes -> C ABI -> CUDA).

Tolerances (norm-wise relative, written here as north_star asks):
  * precision="fp32"   : 1e-4  (FP32-faithful mode; observed ~1e-6)
  * precision="tf32x3" : 1e-4  (3xTF32 split on tcgen05)
  * precision="tf32"   : 5e-2 on gradients (the tiny, cancellation-heavy dWq/dWk set the bound; the
                          large tensors sit near 3e-3), 5e-3 on logits/loss (10-bit mantissa operands --
                          the reference's own cuBLAS TF32 mode, transformer.py:70-80).  Updated weights are
                          compared in units of the step size: Adam's first steps move every weight by
                          ~lr*sign(g), so a sign flip of a near-zero gradient costs 2*lr regardless of
                          precision; the bound is max|dW| <= 2.5*lr*steps.
"""
import hashlib
import os
import pickle

import numpy as np
import pytest

from helpers import FULL_SEQ_CASES, NAMES, load_case, manifest, params_from_case, rel_err
from oracle.transformer_oracle import OracleTransformer, full_sequence_step, global_norm_clip

pytestmark = pytest.mark.gpu

from nlpscratch_b200 import Transformer, b200np  # noqa: E402

TOL = {"fp32": dict(act=1e-4, grad=1e-4, loss=1e-5, param=1e-4),
       "tf32x3": dict(act=1e-4, grad=1e-4, loss=1e-5, param=1e-4),
       "tf32": dict(act=5e-3, grad=5e-2, loss=2e-3, param=None)}


def check_params(got, want, tol, lr, steps, key):
    if tol is not None:
        assert rel_err(got, want) < tol, (key, steps, rel_err(got, want))
    else:
        worst = float(np.max(np.abs(got.astype(np.float64) - want)))
        assert worst <= 2.5 * lr * steps, (key, steps, worst)


def build(case, precision, use_adam=True):
    cfg = [int(v) for v in case["cfg"]]
    m = Transformer(*cfg, dropout_p=0.0, use_adam=use_adam, precision=precision)
    m.load_params(params_from_case(case, 0))
    return m


def check_grads(grads, case, tol, prefix="g_"):
    for k in NAMES:
        got = grads[k].get()
        want = case[prefix + k]
        assert got.shape == want.shape and got.dtype == np.float32, k
        assert rel_err(got, want) < tol, (k, rel_err(got, want))


@pytest.mark.parametrize("precision", ["fp32", "tf32", "tf32x3"])
@pytest.mark.parametrize("name", FULL_SEQ_CASES)
def test_full_sequence_golden(name, precision):
    case = load_case(name)
    tol = TOL[precision]
    m = build(case, precision, bool(case["use_adam"]))
    lr = float(case["lr"])
    steps = sum(1 for k in case if k.startswith("loss"))
    for s in range(steps):
        logits, cache = m.forward(case["x"])
        loss, dlogits = m.calculate_loss_dlogits(logits, case["y"])
        grads = m.backward(cache, dlogits)
        assert abs(float(loss) - float(case["loss%d" % s])) < tol["loss"] * max(1.0, abs(float(case["loss%d" % s])))
        if s == 0:
            assert rel_err(logits.get(), case["logits"]) < tol["act"]
            assert rel_err(cache["H_final"].get(), case["H_final"]) < tol["act"]
            assert rel_err(dlogits.get(), case["dlogits"]) < tol["act"]
            check_grads(grads, case, tol["grad"])
        m.step(grads, lr)
        for k in NAMES:
            check_params(getattr(m, k).get(), case["p%d_%s" % (s + 1, k)], tol["param"], lr, s + 1, k)
    m.close()


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_last_token_driver_path_golden(precision):
    """The body of utils/train.py:104-131, call for call, on the model.np surface."""
    case = load_case("last_token_clip")
    tol = TOL[precision]
    m = build(case, precision)
    npb = m.np
    lr, clip = float(case["lr"]), float(case["clip"])
    x_train, y_train = npb.array(case["x"]), npb.array(case["y"])
    for s in range(2):
        xb, yb = x_train[0:len(y_train)], y_train[0:len(y_train)]
        lens = npb.sum(xb != 0, axis=1).astype(npb.int32)
        s_max = int(npb.max(lens))
        xb = xb[:, :s_max]
        H, cache = m.forward(xb, return_hidden_only=True)
        idx = npb.arange(len(yb))
        t = lens - 1
        H_last = H[idx, t, :]
        logits_last = H_last @ m.W_vocab + m.b_vocab
        loss, dlog_last = m.calculate_loss_dlogits(logits_last, yb)
        grads = m.backward_last_token(cache, t, dlog_last[:, 0, :])
        if s == 0:
            assert rel_err(H.get(), case["H"]) < tol["act"]
            assert rel_err(logits_last.get(), case["logits_last"]) < tol["act"]
            check_grads(grads, case, tol["grad"], prefix="graw_")
            assert np.all(grads["We"].get()[0] == 0)            # PAD row: exactly zero in last-token mode
        # clip_grads(grad_list, clip, lib=npb) -- function.py:27-38 restated on the same surface
        grad_list = list(grads.values())
        total_norm = npb.sqrt(sum(npb.sum(npb.square(g)) for g in grad_list))
        if total_norm > clip:
            scale = clip / (total_norm + 1e-6)
            for g in grad_list:
                g *= scale
        if s == 0:
            check_grads(grads, case, tol["grad"])
        m.step(grads, lr)
        assert abs(float(loss) - float(case["loss%d" % s])) < tol["loss"] * 10
        for k in NAMES:
            check_params(getattr(m, k).get(), case["p%d_%s" % (s + 1, k)], tol["param"], lr, s + 1, k)
        assert int(npb.sum(lens)) == int((case["x"] != 0).sum())
    m.close()


def test_reference_main_smoke_known_answer():
    """transformer.py:618-641 run against the drop-in: same seed => same weights => same loss."""
    ka = manifest()["smoke_main"]
    np.random.seed(42)
    model = Transformer(1000, 64, 8, 256, 2, 50, dropout_p=0.0, precision="fp32")
    for k, sha in ka["param_sha256_16"].items():
        assert hashlib.sha256(np.ascontiguousarray(getattr(model, k).get()).tobytes()).hexdigest()[:16] == sha, k
    assert np.allclose(model.pe.get()[1, :4], ka["pe_1_4"], rtol=0, atol=1e-7)
    x = np.array([[1, 2, 3, 4, 5], [5, 4, 3, 2, 1]], dtype=np.int32)
    y = np.array([[2, 3, 4, 5, 6], [1, 1, 1, 1, 1]], dtype=np.int32)
    logits, cache = model.forward(x)
    loss, dlogits = model.calculate_loss_dlogits(logits, y)
    assert abs(float(loss) - ka["loss_p0"]) < 1e-5
    assert np.allclose(logits.get()[0, 0, :3], ka["logits_0_0_3"], rtol=0, atol=1e-5)
    grads = model.backward(cache, dlogits)
    for k, nrm in ka["grad_norms"].items():
        got = float(np.sqrt(np.sum(np.square(grads[k].get().astype(np.float64)))))
        assert abs(got - nrm) < 1e-4 * max(nrm, 1e-3), k
    model.sgd_step(grads, learning_rate=0.001)
    # dropout on: the loss moves but stays a loss (masks are Philox, not NumPy's stream)
    np.random.seed(42)
    m2 = Transformer(1000, 64, 8, 256, 2, 50, precision="fp32")
    lo, _ = m2.forward(x)
    l2, _ = m2.calculate_loss_dlogits(lo, y)
    assert abs(float(l2) - ka["loss_p0.1"]) < 0.5
    model.close(); m2.close()


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_fused_train_step_matches_oracle_steps(precision):
    """nlps_train_step (forward+loss+backward+clip+Adam in one call) vs the oracle, 3 steps, with a
    padded ragged batch, d=32 heads and an odd vocabulary."""
    rng = np.random.default_rng(21)
    cfg = (301, 64, 2, 128, 2, 40)
    np.random.seed(5)
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    for n in ("b1", "b2", "b_vocab", "ln1_beta", "ln2_beta"):
        oracle.params[n][...] = rng.uniform(-0.1, 0.1, oracle.params[n].shape)
    m = Transformer(*cfg, dropout_p=0.0, precision=precision)
    m.load_params(oracle.params)
    x = rng.integers(1, 301, (5, 33)).astype(np.int32)
    for b, n in enumerate((33, 20, 7, 33, 1)):
        x[b, n:] = 0
    y = rng.integers(0, 301, (5, 33)).astype(np.int32)
    tol = TOL[precision]
    for s in range(3):
        want_loss, _ = full_sequence_step(oracle, x, y, lr=1e-3, clip=0.5)
        loss = m.train_step(x, y, learning_rate=1e-3, clip_grad_norm=0.5)
        assert abs(float(loss) - want_loss) < tol["loss"] * 10 * max(1.0, abs(want_loss)), s
    got = m.get_params()
    for k in NAMES:
        check_params(got[k], oracle.params[k], tol["param"], 1e-3, 3, k)
    assert m.adam_t == 3
    m.close()


def test_data_parallel_identity_on_one_gpu():
    """R shards with dlogits scaled by 1/R and gradients summed == one step on the concatenated batch
     -- the arithmetic the NCCL path relies on, checked without a second GPU."""
    rng = np.random.default_rng(8)
    cfg = (97, 32, 4, 64, 2, 16)
    np.random.seed(1)
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    x = rng.integers(1, 97, (4, 16)).astype(np.int32)
    y = rng.integers(0, 97, (4, 16)).astype(np.int32)
    whole = Transformer(*cfg, dropout_p=0.0, precision="fp32")
    whole.load_params(oracle.params)
    whole.train_step(x, y, defer_update=True)
    g_whole = whole.flat("grad").get()
    acc = np.zeros_like(g_whole)
    for r in range(2):
        shard = Transformer(*cfg, dropout_p=0.0, precision="fp32")
        shard.load_params(oracle.params)
        shard.train_step(x[2 * r:2 * r + 2], y[2 * r:2 * r + 2], grad_scale=0.5, defer_update=True)
        acc += shard.flat("grad").get()
        shard.close()
    assert rel_err(acc, g_whole) < 1e-5
    whole.close()


def test_dropout_is_inverted_and_statistical():
    np.random.seed(0)
    m = Transformer(50, 64, 4, 128, 1, 32, dropout_p=0.25, precision="fp32")
    x = np.random.default_rng(0).integers(1, 50, (8, 32)).astype(np.int32)
    m.eval()
    h_eval, _ = m.forward(x, return_hidden_only=True)
    h_eval2, _ = m.forward(x, return_hidden_only=True)
    assert np.array_equal(h_eval.get(), h_eval2.get())            # eval(): dropout off, deterministic
    m.train()
    h1, c1 = m.forward(x, return_hidden_only=True)
    h2, _ = m.forward(x, return_hidden_only=True)
    assert not np.array_equal(h1.get(), h2.get())                 # fresh masks per forward
    # gradient flows only through kept units: a finite-difference-free sanity check on determinism
    y = np.random.default_rng(1).integers(0, 50, (8, 32)).astype(np.int32)
    l1 = float(m.train_step(x, y, defer_update=True))
    g1 = m.flat("grad").get()
    assert np.isfinite(l1) and np.isfinite(g1).all() and np.abs(g1).sum() > 0
    m.close()


def test_save_produces_reference_pickle(tmp_path):
    np.random.seed(3)
    m = Transformer(30, 16, 2, 32, 2, 8, precision="fp32")
    path = os.path.join(tmp_path, "w.pkl")
    m.save(path)
    with open(path, "rb") as fh:
        w = pickle.load(fh)
    assert list(w.keys()) == list(NAMES)
    assert w["Wq"].shape == (2, 16, 16) and w["W_vocab"].shape == (16, 30) and w["b1"].shape == (2, 32)
    assert all(v.dtype == np.float32 for v in w.values())
    m2 = Transformer(30, 16, 2, 32, 2, 8, precision="fp32")
    m2.load(path)
    assert all(np.array_equal(m2.get_params()[k], w[k]) for k in NAMES)
    m.close(); m2.close()


def test_errors_are_python_exceptions():
    with pytest.raises(AssertionError):
        Transformer(10, 10, 3, 8, 1, 4)
    m = Transformer(10, 8, 2, 8, 1, 4, precision="fp32")
    with pytest.raises(IndexError):
        m.forward(np.array([[1, 11]], np.int32))
    with pytest.raises(RuntimeError):
        m.forward(np.ones((1, 9), np.int32))                      # longer than max_len
    m.close()


def test_baseline_model_size_tf32_agrees_with_fp32():
    """BASELINE.json's quoted model (6 layers, d_model 512, 8 heads, ff 2048, seq 512, vocab 8192) at a
    reduced batch: the tensor-core path (tcgen05 GEMMs + tcgen05 attention) against the FP32-faithful
    CUDA-core path on identical weights and inputs -- loss, every gradient tensor, and the linearity
    property of the step (gradients scale with grad_scale)."""
    L, E, h, F, S, V, B = 6, 512, 8, 2048, 512, 8192, 4
    rng = np.random.default_rng(42)
    x = rng.integers(1, V, (B, S)).astype(np.int32)
    x[1, 400:] = 0                                          # one padded sequence
    y = rng.integers(1, V, (B, S)).astype(np.int32)
    np.random.seed(9)
    ref = Transformer(V, E, h, F, L, S, dropout_p=0.0, precision="fp32")
    params = ref.get_params()
    l_ref = float(ref.train_step(x, y, defer_update=True))
    g_ref = {k: ref._view(1, k).get() for k in NAMES}
    ref.close()
    m = Transformer(V, E, h, F, L, S, dropout_p=0.0, precision="tf32")
    m.load_params(params)
    l_tf = float(m.train_step(x, y, defer_update=True))
    assert abs(l_tf - l_ref) < 2e-3 * abs(l_ref)
    for k in NAMES:
        assert rel_err(m._view(1, k).get(), g_ref[k]) < 5e-2, (k, rel_err(m._view(1, k).get(), g_ref[k]))
    g1 = m.flat("grad").get()
    m.train_step(x, y, grad_scale=0.25, defer_update=True)
    assert rel_err(m.flat("grad").get(), 0.25 * g1) < 1e-6      # linear in the loss scale, bit-stable reductions
    m.close()


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("cfg,B,S", [((13, 6, 2, 10, 1, 5), 1, 1),        # head_dim 3, nothing 4-aligned, S = B = 1
                                     ((31, 12, 3, 18, 2, 9), 2, 9),       # head_dim 4, ff not a multiple of 4
                                     ((64, 40, 5, 100, 1, 70), 3, 70),    # head_dim 8, S > one attention tile
                                     ((257, 96, 3, 200, 2, 33), 2, 33),   # head_dim 32 (tcgen05 attention), odd vocab
                                     # CTA-pair GEMMs with ragged row blocks (210 rows): db1 from the dz epilogue's partials
                                     ((300, 160, 5, 288, 1, 70), 3, 70)])
def test_odd_shapes_against_oracle(cfg, B, S, precision):
    """Ragged / unaligned dimensions: whatever mix of tensor-core and CUDA-core kernels the shapes
    allow must still match the oracle (the reference accepts any embed_dim % num_heads == 0)."""
    rng = np.random.default_rng(sum(cfg) + B + S)
    np.random.seed(sum(cfg))
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    for n in ("b1", "b2", "b_vocab", "ln1_beta", "ln2_beta"):
        oracle.params[n][...] = rng.uniform(-0.1, 0.1, oracle.params[n].shape)
    m = Transformer(*cfg, dropout_p=0.0, precision=precision)
    m.load_params(oracle.params)
    x = rng.integers(1, cfg[0], (B, S)).astype(np.int32)
    if S > 4:
        x[0, S - 2:] = 0
    y = rng.integers(0, cfg[0], (B, S)).astype(np.int32)
    logits, cache = m.forward(x)
    loss, dlogits = m.calculate_loss_dlogits(logits, y)
    grads = m.backward(cache, dlogits)
    lo, co = oracle.forward(x)
    want_loss, dlo = oracle.calculate_loss_dlogits(lo, y)
    go = oracle.backward(co, dlo)
    tol = TOL[precision]
    assert rel_err(logits.get(), lo) < tol["act"]
    assert abs(float(loss) - float(want_loss)) < tol["loss"] * max(1.0, abs(float(want_loss)))
    for k in NAMES:
        if np.linalg.norm(go[k]) < 1e-12:      # e.g. dWq, dWk at S == 1: exactly zero in exact arithmetic
            assert np.max(np.abs(grads[k].get())) < 1e-6, k
        else:
            assert rel_err(grads[k].get(), go[k]) < tol["grad"], (k, rel_err(grads[k].get(), go[k]))
    if S == 1:
        m.close()
        return                                  # Adam on exactly-zero gradients amplifies round-off signs
    m.step(grads, 1e-3)
    oracle.step(go, 1e-3)
    for k in NAMES:
        check_params(getattr(m, k).get(), oracle.params[k], tol["param"], 1e-3, 1, k)
    m.close()


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_fused_last_token_step_golden(precision):
    """nlps_train_step_last_token (the driver's whole iteration in one call) on the reference fixture."""
    case = load_case("last_token_clip")
    tol = TOL[precision]
    m = build(case, precision)
    lr, clip = float(case["lr"]), float(case["clip"])
    x, y = case["x"], case["y"]
    lens = (x != 0).sum(axis=1)
    for s in range(2):
        loss = m.train_step_last_token(x[:, :int(lens.max())], y, lens=lens if s == 0 else None,
                                       learning_rate=lr, clip_grad_norm=clip)
        assert abs(float(loss) - float(case["loss%d" % s])) < tol["loss"] * 10
        for k in NAMES:
            check_params(getattr(m, k).get(), case["p%d_%s" % (s + 1, k)], tol["param"], lr, s + 1, k)
    m.close()


def test_device_resident_train_loop_learns_and_matches_oracle_loop():
    """nlpscratch_b200.train.train_transformer (mirror of utils/train.py:30-174) on a synthetic prefix
    dataset: same eval losses as the oracle running the reference loop's arithmetic, loss goes down."""
    from nlpscratch_b200.train import train_transformer
    from oracle.transformer_oracle import last_token_step
    rng = np.random.default_rng(3)
    V, max_len, n = 60, 24, 96
    X = np.zeros((n, max_len), np.int32)
    for i in range(n):
        L = int(rng.integers(1, max_len))
        X[i, :L] = (np.arange(L) * 7 + i) % (V - 2) + 2          # learnable pattern
    Y = ((X[np.arange(n), (X != 0).sum(1) - 1] + 5) % (V - 2) + 2).astype(np.int32)
    cfg = (V, 32, 4, 64, 2, max_len)
    np.random.seed(4)
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    m = Transformer(*cfg, dropout_p=0.0, precision="fp32")
    m.load_params(oracle.params)
    np.random.seed(123)
    logs = []
    stats = train_transformer(m, X, Y, learning_rate=3e-3, nepoch=3, evaluation_loss_after=1, batch_size=16,
                              print_every=0, clip_grad_norm=1.0, log=logs.append)
    evals = [l for _, l in stats["losses"]]
    assert evals[-1] < evals[0] - 0.2                               # it learns
    # the oracle, driven by the same loop arithmetic and the same shuffles
    np.random.seed(123)
    xs, ys = X.copy(), Y.copy()

    def oracle_eval():
        tot = 0.0
        for i in range(0, n, 16):
            xb, yb = xs[i:i + 16], ys[i:i + 16]
            lens = (xb != 0).sum(1)
            H, _ = oracle.forward(xb[:, :lens.max()], return_hidden_only=True)
            lg = H[np.arange(len(yb)), lens - 1] @ oracle.params["W_vocab"] + oracle.params["b_vocab"]
            tot += float(oracle.calculate_loss_dlogits(lg, yb)[0]) * len(yb)
        return tot / n
    want = []
    for epoch in range(3):
        want.append(oracle_eval())
        perm = np.random.permutation(n)
        xs, ys = xs[perm], ys[perm]
        for i in range(0, n, 16):
            last_token_step(oracle, xs[i:i + 16], ys[i:i + 16], lr=3e-3, clip=1.0)
    assert np.allclose(evals, want, rtol=2e-3, atol=2e-3), (evals, want)
    m.close()


@pytest.mark.gpu
def test_fused_ce_statistics_variant_matches_default():
    """NLPS_FUSE_CE=1 (softmax statistics from the vocab-head GEMM epilogue) gives the same training
    trajectory as the default two-pass loss; run in a subprocess because the switch is read once per process."""
    import subprocess, sys, os, json
    code = r"""
import json, sys, numpy as np
sys.path.insert(0, %r)
from nlpscratch_b200 import Transformer
np.random.seed(3)
m = Transformer(512, 256, 4, 512, 2, 160, dropout_p=0.0, precision="tf32")
rng = np.random.default_rng(0)
x = rng.integers(1, 512, (4, 160)).astype(np.int32); x[1, 100:] = 0
y = rng.integers(0, 512, (4, 160)).astype(np.int32)
print(json.dumps([float(m.train_step(x, y, learning_rate=1e-3, clip_grad_norm=1.0)) for _ in range(3)]))
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = {}
    for flag in ("0", "1"):
        env = dict(os.environ, NLPS_FUSE_CE=flag)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        out[flag] = json.loads(res.stdout.strip().splitlines()[-1])
    for a, b in zip(out["0"], out["1"]):
        assert abs(a - b) < 2e-5 * max(1.0, abs(a)), (out["0"], out["1"])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_cuda_graph_replay_of_the_fused_step_is_bitwise_the_eager_step(precision):
    """nlps_train_step captures the step into a CUDA graph the second time it sees a (cache, x, y) signature and replays it
    afterwards with one patched node (dropout counter, lr, Adam bias corrections).  With device-resident batches the
    trajectory must equal the eager one (NLPS_GRAPH=0) bit for bit -- dropout on, learning rate changing."""
    import subprocess, sys, os, json
    code = r"""
import json, sys, hashlib, numpy as np
sys.path.insert(0, %r)
from nlpscratch_b200 import Transformer
from nlpscratch_b200.array import DeviceArray
np.random.seed(5)
m = Transformer(300, 128, 4, 256, 2, 96, dropout_p=0.1, precision=%r)
rng = np.random.default_rng(1)
xs = [DeviceArray.from_host(rng.integers(1, 300, (6, 96)).astype(np.int32)) for _ in range(2)]
ys = [DeviceArray.from_host(rng.integers(0, 300, (6, 96)).astype(np.int32)) for _ in range(2)]
losses = []
for i in range(9):                      # each batch is seen 4-5 times: eager, capture, replays
    lr = 1e-3 if i < 5 else 5e-4
    losses.append(float(m.train_step(xs[i %% 2], ys[i %% 2], learning_rate=lr, clip_grad_norm=1.0)).hex())
p = m.get_params()
print(json.dumps({"losses": losses, "params": hashlib.sha256(b"".join(np.ascontiguousarray(p[k]).tobytes() for k in sorted(p))).hexdigest()}))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), precision)
    out = {}
    for flag in ("0", "1"):
        env = dict(os.environ, NLPS_GRAPH=flag)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        out[flag] = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["0"]["losses"] == out["1"]["losses"]
    assert out["0"]["params"] == out["1"]["params"]


@pytest.mark.gpu
def test_cuda_graph_of_the_deferred_step_with_external_bucket_events_is_bitwise_the_eager_one():
    """The data-parallel half-step (defer_update: forward, loss, backward; the update follows the all-reduce) is captured
    too; its per-bucket events become external event-record nodes that a communication stream waits on after the graph
    launch.  One rank, NCCL group of size 1 so that DataParallel's bucket waits and all-reduces really run: the
    trajectory must equal the eager one (NLPS_GRAPH_DP=0) bit for bit, dropout on."""
    import subprocess, sys, os, json
    code = r"""
import json, sys, os, hashlib, numpy as np
sys.path.insert(0, %r)
import torch, torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29547")
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
from nlpscratch_b200 import Transformer
from nlpscratch_b200.array import DeviceArray
from nlpscratch_b200.dp import DataParallel
np.random.seed(5)
m = Transformer(300, 128, 4, 256, 2, 96, dropout_p=0.1, precision="tf32", device=0)
dp = DataParallel(m, min_bucket_floats=1)
rng = np.random.default_rng(1)
xs = [DeviceArray.from_host(rng.integers(1, 300, (6, 96)).astype(np.int32)) for _ in range(2)]
ys = [DeviceArray.from_host(rng.integers(0, 300, (6, 96)).astype(np.int32)) for _ in range(2)]
losses = []
for i in range(9):
    lr = 1e-3 if i < 5 else 5e-4
    l = m.train_step(xs[i %% 2], ys[i %% 2], learning_rate=lr, clip_grad_norm=1.0, grad_scale=0.5, defer_update=True)
    dp.allreduce_gradients()
    m.apply_update(lr, 1.0)
    losses.append(float(l).hex())
p = m.get_params()
print(json.dumps({"losses": losses, "launches": m.launch_count(), "params": hashlib.sha256(b"".join(np.ascontiguousarray(p[k]).tobytes() for k in sorted(p))).hexdigest()}))
dist.destroy_process_group()
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),)
    out = {}
    for flag in ("0", "1"):
        env = dict(os.environ, NLPS_GRAPH_DP=flag)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        out[flag] = json.loads([l for l in res.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert out["0"]["losses"] == out["1"]["losses"]
    assert out["0"]["params"] == out["1"]["params"]


@pytest.mark.gpu
@pytest.mark.parametrize("heads", [4, 8])
def test_delta_from_the_wo_dgrad_epilogue_matches_the_separate_pass(heads, tmp_path):
    """The row term of the softmax backward, delta = sum(dO * O) per (token, head) (function.py:59-65 moved through P@V),
    is a by-product of the dctx = d_a Wo^T GEMM epilogue (GemmArgs::rowdot_out).  NLPS_FUSE_DELTA=0 runs the separate
    pass instead; the attention-path gradients must agree to FP32 summation-order accuracy (head_dim 64 and 32, PAD tail,
    a row count that is not a multiple of the 256-row tile)."""
    import subprocess, sys, os
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from nlpscratch_b200 import Transformer
np.random.seed(3)
m = Transformer(512, 256, %d, 512, 2, 160, dropout_p=0.0, precision="tf32")
rng = np.random.default_rng(0)
x = rng.integers(1, 512, (5, 160)).astype(np.int32); x[1, 100:] = 0
y = rng.integers(0, 512, (5, 160)).astype(np.int32)
logits, cache = m.forward(x)
loss, dl = m.calculate_loss_dlogits(logits, y)
g = m.backward(cache, dl)
np.savez(sys.argv[1], **{k: np.asarray(g[k].get()) for k in ("Wq", "Wk", "Wv", "We")})
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), heads)
    out = {}
    for flag in ("0", "1"):
        path = os.path.join(tmp_path, "g%s.npz" % flag)
        env = dict(os.environ, NLPS_FUSE_DELTA=flag)
        res = subprocess.run([sys.executable, "-c", code, path], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        out[flag] = np.load(path)
    for k in ("Wq", "Wk", "Wv", "We"):
        a, b = out["0"][k].astype(np.float64), out["1"][k].astype(np.float64)
        assert np.isfinite(b).all()
        # delta itself moves by ~1e-7 relative; dS is then rounded onto the TF32 grid, where a flipped rounding is 2^-11
        assert np.linalg.norm(a - b) <= 5e-4 * np.linalg.norm(a), (k, np.linalg.norm(a - b) / np.linalg.norm(a))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_weight_gradient_stream_is_bitwise_the_single_stream_step(precision):
    """dW2 / dW1 / dWo / dWqkv / dW_vocab run on a second stream, ordered by events against the buffers they read
    (NLPS_WGRAD_STREAM=0 keeps them on the main stream).  Same kernels, same arguments, one owner per output: fused steps
    (eager, captured, replayed), the unfused forward/backward API and the last-token driver step must all give the same
    bits, dropout on."""
    import subprocess, sys, os, json
    code = r"""
import json, sys, hashlib, numpy as np
sys.path.insert(0, %r)
from nlpscratch_b200 import Transformer
from nlpscratch_b200.array import DeviceArray
np.random.seed(5)
m = Transformer(300, 128, 4, 256, 3, 96, dropout_p=0.1, precision=%r)
rng = np.random.default_rng(1)
xh = [rng.integers(1, 300, (6, 96)).astype(np.int32) for _ in range(2)]
xh[1][2, 60:] = 0
xs = [DeviceArray.from_host(a) for a in xh]
ys = [DeviceArray.from_host(rng.integers(0, 300, (6, 96)).astype(np.int32)) for _ in range(2)]
losses = []
for i in range(8):
    losses.append(float(m.train_step(xs[i %% 2], ys[i %% 2], learning_rate=1e-3, clip_grad_norm=1.0)).hex())
logits, cache = m.forward(xh[0])                      # the unfused API
loss, dl = m.calculate_loss_dlogits(logits, ys[0])
g = m.backward(cache, dl)
gh = hashlib.sha256(b"".join(np.ascontiguousarray(g[k].get()).tobytes() for k in sorted(g))).hexdigest()
m.step(g, 1e-3)
yl = rng.integers(0, 300, (6,)).astype(np.int32)      # the driver's last-token step
losses.append(float(m.train_step_last_token(xh[1], yl, learning_rate=1e-3, clip_grad_norm=1.0)).hex())
p = m.get_params()
print(json.dumps({"losses": losses, "grads": gh, "params": hashlib.sha256(b"".join(np.ascontiguousarray(p[k]).tobytes() for k in sorted(p))).hexdigest()}))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), precision)
    out = {}
    for flag in ("0", "1"):
        env = dict(os.environ, NLPS_WGRAD_STREAM=flag)
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        out[flag] = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["0"] == out["1"]


@pytest.mark.gpu
def test_resume_checkpoint_restores_parameters_and_adam_state(tmp_path):
    """save_checkpoint / load_checkpoint: a model restored from the checkpoint continues bit for bit like the
    one that wrote it (Adam moments and step count included); a reference-format ``save`` pickle loads with fresh Adam."""
    from nlpscratch_b200 import Transformer
    cfg = (97, 64, 2, 128, 2, 24)
    rng = np.random.default_rng(0)
    x = rng.integers(1, 97, (4, 24)).astype(np.int32)
    y = rng.integers(0, 97, (4, 24)).astype(np.int32)
    np.random.seed(1)
    a = Transformer(*cfg, dropout_p=0.0, precision="fp32")
    for _ in range(2):
        a.train_step(x, y, 1e-3, 1.0)
    ck, ref_fmt = os.path.join(tmp_path, "ck.pkl"), os.path.join(tmp_path, "ref.pkl")
    a.save_checkpoint(ck)
    a.save(ref_fmt)
    for _ in range(2):
        a.train_step(x, y, 1e-3, 1.0)
    want = a.get_params()
    np.random.seed(2)
    b = Transformer(*cfg, dropout_p=0.0, precision="fp32")
    b.load_checkpoint(ck)
    assert b.adam_t == 2
    for _ in range(2):
        b.train_step(x, y, 1e-3, 1.0)
    got = b.get_params()
    for k in want:
        assert np.array_equal(want[k], got[k]), k
    b.load_checkpoint(ref_fmt)                      # reference-format pickle: parameters only, Adam restarts
    assert b.adam_t == 0
    assert float(np.abs(b.m["We"].get()).max()) == 0.0
    a.close(); b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_fused_vocabulary_path_matches_the_materialising_one_and_the_oracle(precision, tmp_path):
    """NLPS_FUSE_VOCAB=1: vocabulary head, cross-entropy and vocabulary backward walk the rows in L2-sized
    chunks and never materialise the (N, V) logits / dlogits.  Same kernels and per-row arithmetic as the materialising path;
    only dW_vocab / db_vocab are summed chunk by chunk.  Four chunks here (the last one ragged), odd vocabulary, PAD tail,
    dropout on; compared with the materialising path (same Philox masks) and, with dropout off, with the oracle."""
    import subprocess, sys, os, json
    code = r"""
import json, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from nlpscratch_b200 import Transformer
from oracle.transformer_oracle import OracleTransformer, full_sequence_step
cfg = (641, 128, 4, 256, 2, 160)
rng = np.random.default_rng(0)
x = rng.integers(1, 641, (6, 160)).astype(np.int32); x[1, 100:] = 0
y = rng.integers(0, 641, (6, 160)).astype(np.int32)
out = {}
for p in (0.1, 0.0):
    np.random.seed(3)
    oracle = OracleTransformer(*cfg, dropout_p=0.0)
    m = Transformer(*cfg, dropout_p=p, precision=%r, seed=5)
    m.load_params(oracle.params)
    losses = [float(m.train_step(x, y, learning_rate=1e-3, clip_grad_norm=1.0)) for _ in range(4)]
    params = m.get_params()
    out[str(p)] = {"losses": losses, "W_vocab": params["W_vocab"].astype(np.float64).ravel()[::97].tolist(),
                   "b_vocab": params["b_vocab"].astype(np.float64).tolist(), "We": params["We"].astype(np.float64).ravel()[::131].tolist()}
    if p == 0.0:
        want = [full_sequence_step(oracle, x, y, lr=1e-3, clip=1.0)[0] for _ in range(4)]
        out["oracle"] = {"losses": want, "b_vocab": oracle.params["b_vocab"].astype(np.float64).tolist()}
    m.close()
print(json.dumps(out))
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), precision)
    res = {}
    for flag in ("0", "1"):
        env = dict(os.environ, NLPS_FUSE_VOCAB=flag, NLPS_VOCAB_CHUNK_MB="1")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-3000:]
        res[flag] = json.loads(r.stdout.strip().splitlines()[-1])
    tol = 2e-5 if precision == "fp32" else 2e-3
    for p in ("0.1", "0.0"):
        a, b = res["0"][p], res["1"][p]
        assert np.allclose(a["losses"], b["losses"], rtol=tol, atol=0), (p, a["losses"], b["losses"])
        for k in ("W_vocab", "b_vocab", "We"):
            assert np.max(np.abs(np.array(a[k]) - np.array(b[k]))) <= (1e-6 if precision == "fp32" else 2.5e-3 * 4), (p, k)
    o = res["1"]["oracle"]
    assert np.allclose(res["1"]["0.0"]["losses"], o["losses"], rtol=1e-5 if precision == "fp32" else 2e-3, atol=0)
