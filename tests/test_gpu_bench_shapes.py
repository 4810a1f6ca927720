"""Per-kernel float64 checks at the EXACT shapes the benchmarked configurations launch:
every GEMM of the c4 step (6L d512 ff2048, 32768 tokens, V 8192) and of the c5 step (d768 ff3072, 16384 tokens) --
forward, dgrad (B stored transposed) and the split-K weight gradients (A stored transposed, K = tokens) -- and the tcgen05
attention at (B, S, h, d) = (4, 512, 8, 64) and (1, 2048, 12, 64).  The float64 reference is evaluated on a deterministic
sample of output rows (a full f64 product of 32768 x 8192 x 512 would take minutes on the host)."""
import ctypes as C

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu

from gpu_util import gemm  # noqa: E402
from nlpscratch_b200 import _lib  # noqa: E402

TF32_TOL, X3_TOL = 2e-3, 2e-6

T4, T5 = 32768, 16384
C4 = [  # (M, N, K, trans_a, trans_b, what)
    (T4, 1536, 512, False, False, "qkv"), (T4, 512, 512, False, False, "wo"), (T4, 2048, 512, False, False, "ffn1"),
    (T4, 512, 2048, False, False, "ffn2"), (8192, 8192, 512, False, False, "vocab (8192 of the 32768 rows)"),
    (T4, 512, 8192, False, True, "dH = dlogits Wv^T"), (T4, 2048, 512, False, True, "dhid = d_f W2^T"),
    (T4, 512, 2048, False, True, "dy1 = dz W1^T"), (T4, 512, 512, False, True, "dctx = d_a Wo^T"),
    (T4, 512, 1536, False, True, "dx = dqkv Wqkv^T"),
    (512, 8192, T4, True, False, "dW_vocab"), (2048, 512, T4, True, False, "dW2"), (512, 2048, T4, True, False, "dW1"),
    (512, 512, T4, True, False, "dWo"), (512, 1536, T4, True, False, "dWqkv")]
C5 = [(T5, 2304, 768, False, False, "qkv"), (T5, 768, 768, False, False, "wo"), (T5, 3072, 768, False, False, "ffn1"),
      (T5, 768, 3072, False, False, "ffn2"), (T5, 768, 3072, False, True, "dy1"), (T5, 768, 2304, False, True, "dx"),
      (768, 8192, T5, True, False, "dW_vocab"), (3072, 768, T5, True, False, "dW2"), (768, 3072, T5, True, False, "dW1"),
      (768, 768, T5, True, False, "dWo"), (768, 2304, T5, True, False, "dWqkv")]


def sampled_rows_ref(a, b, ta, tb, rows):
    A = (a[:, rows].T if ta else a[rows, :]).astype(np.float64)
    B = (b.T if tb else b).astype(np.float64)
    return A @ B


@pytest.mark.parametrize("M,N,K,ta,tb,what", C4 + C5)
def test_tf32_gemm_at_benchmark_shapes(M, N, K, ta, tb, what):
    rng = np.random.default_rng(M + 3 * N + 7 * K + ta + 2 * tb)
    a = rng.standard_normal((K, M) if ta else (M, K), dtype=np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N), dtype=np.float32)
    before = (C.c_int64 * 2)()
    _lib.call("nlps_kernel_stats", before, C.c_int(2), C.c_int(1))
    got = gemm(a, b, "tf32", ta, tb)
    after = (C.c_int64 * 2)()
    _lib.call("nlps_kernel_stats", after, C.c_int(2), C.c_int(0))
    assert after[0] == 1, "expected the CTA-pair tcgen05 kernel for %s" % what
    rows = np.unique(np.linspace(0, M - 1, 48).astype(np.int64))
    assert rel_err(got[rows], sampled_rows_ref(a, b, ta, tb, rows)) < TF32_TOL, what
    if ta:                                         # split-K weight gradients: fixed-order reduction, bitwise repeatable
        assert np.array_equal(got, gemm(a, b, "tf32", ta, tb)), what


@pytest.mark.parametrize("M,N,K,ta,tb,what", [C4[2], C4[6], C4[11], C5[3]])
def test_tf32x3_gemm_at_benchmark_shapes(M, N, K, ta, tb, what):
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((K, M) if ta else (M, K), dtype=np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N), dtype=np.float32)
    got = gemm(a, b, "tf32x3", ta, tb)
    rows = np.unique(np.linspace(0, M - 1, 48).astype(np.int64))
    # FP32 accumulation over K terms (TMEM accumulators, split-K partials added in FP32): the error grows ~ sqrt(K) * 2^-24;
    # K = 32768 measures 8.6e-6 -- FP32-level, still an order of magnitude inside the 1e-4 bar of the FP32-faithful mode
    assert rel_err(got[rows], sampled_rows_ref(a, b, ta, tb, rows)) < X3_TOL * max(1.0, (K / 512.0) ** 0.5), what
