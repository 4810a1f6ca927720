"""Per-kernel parity on the GPU (CUDA-core FP32 mode and integer kernels): every operator of the C ABI
against a float64 NumPy statement of the reference line it replaces."""
import ctypes as C

import numpy as np
import pytest

from helpers import rel_err
from oracle.transformer_oracle import additive_mask, affine_norm, affine_norm_grad, row_softmax, sinusoid_table

pytestmark = pytest.mark.gpu

from nlpscratch_b200 import _lib  # noqa: E402
from nlpscratch_b200.array import DeviceArray, b200np  # noqa: E402
from gpu_util import S0, dev, gemm, gemm_ref  # noqa: E402

FP32_TOL = 2e-6      # CUDA-core FP32 FMA vs float64: well inside the 1e-4 bar of north_star


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (64, 64, 16), (70, 130, 37), (257, 66, 129)])
def test_gemm_fp32_layouts(M, N, K, ta, tb):
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    got = gemm(a, b, "fp32", ta, tb)
    assert rel_err(got, gemm_ref(a, b, ta, tb)) < FP32_TOL


def test_gemm_fp32_epilogue():
    rng = np.random.default_rng(0)
    M, N, K = 45, 52, 33
    a, b = rng.standard_normal((M, K)).astype(np.float32), rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    gate = rng.standard_normal((M, N)).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float32)
    acc = rng.standard_normal((M, N)).astype(np.float32)
    assert rel_err(gemm(a, b, "fp32", bias=bias, relu=True), gemm_ref(a, b, bias=bias, relu=True)) < FP32_TOL
    assert rel_err(gemm(a, b, "fp32", gate=gate, residual=res), gemm_ref(a, b, gate=gate, residual=res)) < FP32_TOL
    assert rel_err(gemm(a, b, "fp32", accumulate_into=acc), gemm_ref(a, b, acc=acc)) < FP32_TOL


def _attention_ref(qkv, x, h):
    """float64 statement of transformer.py:131-141 on packed (B,S,3E) rows."""
    B, S, E3 = qkv.shape
    E = E3 // 3
    d = E // h
    q, k, v = (qkv[..., i * E:(i + 1) * E].astype(np.float64).reshape(B, S, h, d).transpose(0, 2, 1, 3) for i in range(3))
    s = np.einsum("bhid,bhjd->bhij", q, k) / np.sqrt(d) + additive_mask(x)
    p = row_softmax(s)
    ctx = np.einsum("bhij,bhjd->bhid", p, v).transpose(0, 2, 1, 3).reshape(B, S, E)
    return q, k, v, p, ctx


def _attention_case(B, S, h, d, pad_mode, seed):
    rng = np.random.default_rng(seed)
    E = h * d
    qkv = rng.standard_normal((B, S, 3 * E)).astype(np.float32)
    x = rng.integers(1, 50, (B, S)).astype(np.int32)
    if pad_mode == "tail":
        for b in range(B):
            x[b, max(1, S - 1 - 3 * b):] = 0
    elif pad_mode == "lead":
        x[0, :min(3, S)] = 0
        if B > 1:
            x[1, S // 2:] = 0
        if B > 2:
            x[2, :] = 0                       # an all-PAD sequence: every row attends to all keys
    return qkv, x


@pytest.mark.parametrize("B,S,h,d,pad", [(2, 5, 8, 8, "none"), (3, 10, 4, 8, "tail"), (3, 8, 4, 8, "lead"),
                                         (2, 77, 2, 32, "tail"), (1, 130, 2, 64, "none"), (2, 65, 1, 16, "lead"),
                                         (1, 64, 1, 128, "none"), (2, 33, 3, 20, "tail"),
                                         # multi-block shapes for the tcgen05 kernels (head_dim 32 / 64)
                                         (2, 320, 2, 32, "tail"), (3, 300, 2, 64, "lead"), (2, 512, 2, 64, "none"),
                                         (1, 1100, 1, 64, "tail"), (3, 257, 3, 32, "lead"),
                                         # more tiles than SMs: every persistent backward CTA walks several tiles
                                         # (one-block tiles and ragged last tiles included: S = 130, 65, 320)
                                         (8, 512, 6, 64, "tail"), (40, 130, 4, 32, "lead"), (33, 65, 5, 64, "tail"),
                                         (16, 320, 8, 32, "tail"),
                                         # beyond the staged PAD-mask capacity of the forward (2048) and dQ (4096) kernels
                                         (1, 2100, 1, 64, "tail"), (1, 4200, 1, 32, "none"),
                                         # the benchmarked attention shapes: c4 (S512 h8 d64) and c5 (S2048 h12 d64)
                                         (4, 512, 8, 64, "tail"), (1, 2048, 12, 64, "tail")])
@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3"])
def test_attention_forward_backward(B, S, h, d, pad, prec):
    # fp32: CUDA-core kernels (attention.cu); tf32: tcgen05 / mma.sync tensor-core kernels (operands rounded to TF32:
    # 10 mantissa bits => ~1e-3 relative, asserted < 4e-3 / 8e-3); tf32x3: 3xTF32 split products on mma.sync
    # (attention_mma.cu, head_dim <= 64 and a multiple of 4; CUDA-core kernels otherwise): FP32-level, same bounds as fp32
    PREC = C.c_int(_lib.PREC[prec])
    tol_f, tol_b = (4e-3, 8e-3) if prec == "tf32" else (5e-6, 1e-5)
    qkv, x = _attention_case(B, S, h, d, pad, seed=S * 7 + d)
    E = h * d
    QKV, X = dev(qkv), dev(x)
    ctx = DeviceArray.empty((B, S, E))
    lse = DeviceArray.empty((B, h, S))
    _lib.call("nlps_attention_forward", S0, PREC, C.c_int(B), C.c_int(S), C.c_int(h), C.c_int(d),
              _lib.vp(QKV.ptr), _lib.vp(X.ptr), _lib.vp(ctx.ptr), _lib.vp(lse.ptr))
    q, k, v, p, ctx_ref = _attention_ref(qkv, x, h)
    assert rel_err(ctx.get(), ctx_ref) < tol_f
    # backward against the analytic formulas of transformer.py:430-437
    rng = np.random.default_rng(1)
    dctx = rng.standard_normal((B, S, E)).astype(np.float32)
    DC = dev(dctx)
    dqkv = DeviceArray.empty((B, S, 3 * E))
    _lib.call("nlps_attention_backward", S0, PREC, C.c_int(B), C.c_int(S), C.c_int(h), C.c_int(d),
              _lib.vp(QKV.ptr), _lib.vp(X.ptr), _lib.vp(ctx.ptr), _lib.vp(lse.ptr), _lib.vp(DC.ptr), _lib.vp(dqkv.ptr))
    do = dctx.astype(np.float64).reshape(B, S, h, d).transpose(0, 2, 1, 3)
    dp = np.einsum("bhid,bhjd->bhij", do, v)
    dv = np.einsum("bhij,bhid->bhjd", p, do)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True))
    dq = np.einsum("bhij,bhjd->bhid", ds, k) / np.sqrt(d)
    dk = np.einsum("bhij,bhid->bhjd", ds, q) / np.sqrt(d)
    want = np.concatenate([t.transpose(0, 2, 1, 3).reshape(B, S, E) for t in (dq, dk, dv)], axis=-1)
    assert rel_err(dqkv.get(), want) < tol_b


def test_attention_tensor_core_path_is_bitwise_deterministic_across_launches_and_shapes():
    """The persistent attention kernels draw their tiles from a global counter, so which CTA computes which tile
    differs from launch to launch -- the results must not (one owner per output, fixed summation order), and the
    self-resetting counters must survive launches of other shapes on the same stream."""
    PREC = C.c_int(_lib.PREC["tf32"])

    def run(B, S, h, d, seed):
        qkv, x = _attention_case(B, S, h, d, "tail", seed)
        E = h * d
        QKV, X = dev(qkv), dev(x)
        ctx, lse = DeviceArray.empty((B, S, E)), DeviceArray.empty((B, h, S))
        _lib.call("nlps_attention_forward", S0, PREC, C.c_int(B), C.c_int(S), C.c_int(h), C.c_int(d),
                  _lib.vp(QKV.ptr), _lib.vp(X.ptr), _lib.vp(ctx.ptr), _lib.vp(lse.ptr))
        DC = dev(np.random.default_rng(seed + 1).standard_normal((B, S, E)).astype(np.float32))
        dqkv = DeviceArray.empty((B, S, 3 * E))
        _lib.call("nlps_attention_backward", S0, PREC, C.c_int(B), C.c_int(S), C.c_int(h), C.c_int(d),
                  _lib.vp(QKV.ptr), _lib.vp(X.ptr), _lib.vp(ctx.ptr), _lib.vp(lse.ptr), _lib.vp(DC.ptr), _lib.vp(dqkv.ptr))
        return ctx.get().copy(), lse.get().copy(), dqkv.get().copy()

    first = run(24, 300, 4, 64, 5)          # 288 tiles: several per persistent CTA
    run(3, 77, 2, 32, 6)                    # other shapes in between
    run(40, 130, 4, 32, 7)
    again = run(24, 300, 4, 64, 5)
    for a, b in zip(first, again):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("rows,E", [(1, 8), (7, 32), (33, 64), (19, 128), (50, 256), (9, 512), (5, 100)])
def test_layernorm_forward_backward(rows, E):
    rng = np.random.default_rng(rows + E)
    x = (rng.standard_normal((1, rows, E)) * 2 + 0.5).astype(np.float32)
    g = rng.uniform(0.5, 1.5, E).astype(np.float32)
    b = rng.uniform(-0.2, 0.2, E).astype(np.float32)
    dy = rng.standard_normal((1, rows, E)).astype(np.float32)
    X, G, Bt, DY = dev(x), dev(g), dev(b), dev(dy)
    y, mean, rstd = DeviceArray.empty((rows, E)), DeviceArray.empty((rows,)), DeviceArray.empty((rows,))
    _lib.call("nlps_layernorm_forward", S0, C.c_int64(rows), C.c_int(E), _lib.vp(X.ptr), _lib.vp(G.ptr),
              _lib.vp(Bt.ptr), _lib.vp(y.ptr), _lib.vp(mean.ptr), _lib.vp(rstd.ptr))
    y_ref, saved = affine_norm(x.astype(np.float64), g.astype(np.float64), b.astype(np.float64))
    assert rel_err(y.get(), y_ref[0]) < 2e-6
    dx, dg, db = DeviceArray.empty((rows, E)), DeviceArray.empty((E,)), DeviceArray.empty((E,))
    _lib.call("nlps_layernorm_backward", S0, C.c_int64(rows), C.c_int(E), _lib.vp(DY.ptr), _lib.vp(X.ptr),
              _lib.vp(mean.ptr), _lib.vp(rstd.ptr), _lib.vp(G.ptr), _lib.vp(dx.ptr), _lib.vp(dg.ptr), _lib.vp(db.ptr))
    dx_ref, dg_ref, db_ref = affine_norm_grad(dy.astype(np.float64), saved)
    assert rel_err(dx.get(), dx_ref[0]) < 5e-6
    assert rel_err(dg.get(), dg_ref) < 5e-6
    assert rel_err(db.get(), db_ref) < 5e-6


def test_embedding_gather_is_bit_exact_and_scatter_matches_add_at():
    rng = np.random.default_rng(3)
    V, E, B, S = 37, 24, 4, 19
    We = rng.standard_normal((V, E)).astype(np.float32)
    pe = sinusoid_table(32, E)
    x = rng.integers(0, V, (B, S + 3)).astype(np.int32)
    x[:, 5] = 0
    xs = x[:, :S]                                           # the driver's non-contiguous slice
    Xd, Wd, Pd = dev(x), dev(We), dev(pe)              # keep the device buffers alive across the call
    out, xf, err = DeviceArray.empty((B, S, E)), DeviceArray.empty((B * S,), np.int32), b200np.zeros((1,), np.int32)
    _lib.call("nlps_embed_forward", S0, C.c_int(B), C.c_int(S), C.c_int(E), C.c_int(V), _lib.vp(Xd.ptr),
              C.c_int64(S + 3), _lib.vp(Wd.ptr), _lib.vp(Pd.ptr), _lib.vp(out.ptr), _lib.vp(xf.ptr),
              _lib.vp(err.ptr))
    assert np.array_equal(out.get(), We[xs] + pe[:S])       # bit-exact (transformer.py:114)
    assert np.array_equal(xf.get(), xs.reshape(-1))
    assert int(err.get()[0]) == 0
    dh = rng.standard_normal((B * S, E)).astype(np.float32)
    dWe, DH = DeviceArray.empty((V, E)), dev(dh)
    _lib.call("nlps_embed_backward", S0, C.c_int64(B * S), C.c_int(E), C.c_int(V), _lib.vp(xf.ptr),
              _lib.vp(DH.ptr), _lib.vp(dWe.ptr))
    want = np.zeros((V, E), np.float32)
    np.add.at(want, xs.reshape(-1), dh)                     # transformer.py:453-457, same ascending order
    assert np.array_equal(dWe.get(), want)                  # deterministic AND bit-identical to NumPy
    again = DeviceArray.empty((V, E))
    _lib.call("nlps_embed_backward", S0, C.c_int64(B * S), C.c_int(E), C.c_int(V), _lib.vp(xf.ptr),
              _lib.vp(DH.ptr), _lib.vp(again.ptr))
    assert np.array_equal(again.get(), dWe.get())


def test_embedding_scatter_heavy_collisions():
    rng = np.random.default_rng(4)
    V, E, N = 11, 128, 5000
    x = rng.choice(V, size=N, p=np.array([0.6] + [0.04] * 10)).astype(np.int32)    # Zipf-like, PAD heavy
    dh = rng.standard_normal((N, E)).astype(np.float32)
    dWe, Xd, DH = DeviceArray.empty((V, E)), dev(x), dev(dh)
    _lib.call("nlps_embed_backward", S0, C.c_int64(N), C.c_int(E), C.c_int(V), _lib.vp(Xd.ptr),
              _lib.vp(DH.ptr), _lib.vp(dWe.ptr))
    want = np.zeros((V, E), np.float32)
    np.add.at(want, x, dh)
    assert np.array_equal(dWe.get(), want)


def test_out_of_range_id_is_flagged():
    V, E = 10, 8
    x = np.array([[1, 12, 3]], np.int32)
    err = b200np.zeros((1,), np.int32)
    out, Xd, Wd, Pd = DeviceArray.empty((1, 3, E)), dev(x), b200np.zeros((V, E)), b200np.zeros((4, E))
    _lib.call("nlps_embed_forward", S0, C.c_int(1), C.c_int(3), C.c_int(E), C.c_int(V), _lib.vp(Xd.ptr),
              C.c_int64(3), _lib.vp(Wd.ptr), _lib.vp(Pd.ptr), _lib.vp(out.ptr),
              None, _lib.vp(err.ptr))
    assert int(err.get()[0]) == 1


@pytest.mark.parametrize("rows,V", [(3, 50), (6, 574), (4, 1000), (2, 8192), (1, 33000)])
def test_softmax_cross_entropy(rows, V):
    from nlpscratch_b200 import Transformer
    rng = np.random.default_rng(rows * V)
    logits = (rng.standard_normal((rows, V)) * 3).astype(np.float32)
    y = rng.integers(0, V, rows).astype(np.int32)
    m = Transformer(V, 8, 2, 8, 1, 4, precision="fp32")
    loss, d = m.calculate_loss_dlogits(dev(logits), dev(y))
    z = logits.astype(np.float64)
    z = z - z.max(-1, keepdims=True)
    p = np.exp(z) / np.exp(z).sum(-1, keepdims=True)
    want_loss = -np.mean(np.log(p[np.arange(rows), y]))
    p[np.arange(rows), y] -= 1
    assert abs(float(loss) - want_loss) < 1e-5 * max(1.0, abs(want_loss))
    assert d.shape == (rows, 1, V)
    assert rel_err(d.get()[:, 0, :], p / rows) < 5e-6
    m.close()


def test_colsum_sumsq_deterministic():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1234, 77)).astype(np.float32)
    X = dev(x)
    out = DeviceArray.empty((77,))
    _lib.call("nlps_colsum", S0, C.c_int64(1234), C.c_int(77), _lib.vp(X.ptr), C.c_int64(77), _lib.vp(out.ptr))
    assert rel_err(out.get(), x.astype(np.float64).sum(0)) < 2e-6
    s1, s2 = DeviceArray.empty((1,)), DeviceArray.empty((1,))
    for s in (s1, s2):
        _lib.call("nlps_sumsq", S0, C.c_int64(x.size), _lib.vp(X.ptr), _lib.vp(s.ptr))
    assert s1.get()[0] == s2.get()[0]
    assert abs(s1.get()[0] - np.sum(x.astype(np.float64) ** 2)) < 1e-5 * x.size


def test_array_module_surface():
    """The calls utils/train.py:46-116 and function.py:27-38 make on model.np / device arrays."""
    rng = np.random.default_rng(6)
    xh = rng.integers(0, 9, (10, 12)).astype(np.int32)
    xh[:, 8:] = 0
    x = b200np.array(xh)
    assert len(x) == 10
    perm = np.random.RandomState(0).permutation(10)
    assert np.array_equal(x[b200np.array(perm)].get(), xh[perm])
    xb = x[2:7]
    lens = b200np.sum(xb != 0, axis=1).astype(b200np.int32)
    want_lens = np.sum(xh[2:7] != 0, axis=1)
    assert np.array_equal(lens.get(), want_lens)
    assert lens.size == 5
    s_max = int(b200np.max(lens))
    assert s_max == want_lens.max()
    assert np.array_equal(xb[:, :s_max].get(), xh[2:7, :s_max])
    t = lens - 1
    assert np.array_equal(t.get(), want_lens - 1)
    Hh = rng.standard_normal((5, s_max, 16)).astype(np.float32)
    H = b200np.array(Hh)
    idx = b200np.arange(5)
    assert np.array_equal(H[idx, t, :].get(), Hh[np.arange(5), np.maximum(want_lens - 1, 0), :])
    Wh, bh = rng.standard_normal((16, 23)).astype(np.float32), rng.standard_normal(23).astype(np.float32)
    logits = H[idx, t, :] @ b200np.array(Wh) + b200np.array(bh)
    assert rel_err(logits.get(), Hh[np.arange(5), want_lens - 1] @ Wh + bh) < 2e-3
    assert int(b200np.sum(lens)) == want_lens.sum()
    nl = H[0, -1] @ b200np.array(Wh) + b200np.array(bh)
    probs = b200np.exp(nl - nl.max())
    probs /= probs.sum()
    ref = np.exp((Hh[0, -1] @ Wh + bh) - (Hh[0, -1] @ Wh + bh).max())
    assert rel_err(b200np.asnumpy(probs), ref / ref.sum()) < 2e-3
    assert np.array(probs).shape == (23,)
    # clip_grads' surface: sqrt(sum(sum(square(g)))), `>`, in-place scale (also on a strided view)
    g = b200np.array(Hh)
    total = b200np.sqrt(sum(b200np.sum(b200np.square(a)) for a in [g, g[:, 0, :]]))
    want = np.sqrt(np.sum(Hh ** 2) + np.sum(Hh[:, 0, :] ** 2))
    assert abs(float(total) - want) < 1e-4 * want
    assert bool(total > 1.0)
    view = g[:, 1, :]
    view *= 0.5
    Hh[:, 1, :] *= 0.5
    assert np.array_equal(g.get(), Hh)
    seq = b200np.array([1, 2, 3], dtype=b200np.int32)[None, :]
    assert seq.shape == (1, 3)
