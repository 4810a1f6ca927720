"""tcgen05 / TMEM / TMA GEMM (kind::tf32) against float64 NumPy: all operand majors, ragged edges,
split-K, fused epilogue, and the 3xTF32 split.

Tolerance: TF32 keeps 10 mantissa bits per operand, FP32 accumulation in TMEM: norm-wise relative
error ~ 2^-11 * sqrt(2) ~ 7e-4; asserted < 2e-3.  3xTF32: < 2e-6 (FP32 level)."""
import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu

from gpu_util import gemm, gemm_ref  # noqa: E402

TF32_TOL, X3_TOL = 2e-3, 2e-6
SHAPES = [(128, 128, 32), (128, 256, 64), (256, 512, 128), (200, 136, 72), (130, 260, 36), (1, 8, 4),
          (4096, 512, 512), (16, 576, 64), (300, 1536, 256)]


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tf32_layouts(M, N, K, ta, tb):
    rng = np.random.default_rng(M + 3 * N + 7 * K)
    # stored leading dims must be multiples of 4 floats for TMA
    if (ta and M % 4) or (not ta and K % 4) or (tb and K % 4) or (not tb and N % 4):
        pytest.skip("operand not TMA-addressable; covered by the FP32 path")
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    got = gemm(a, b, "tf32", ta, tb)
    assert rel_err(got, gemm_ref(a, b, ta, tb)) < TF32_TOL


def test_tf32_wgrad_shape_uses_split_k_and_is_deterministic():
    rng = np.random.default_rng(11)
    K, M, N = 8192, 256, 384                     # X^T dY with few output tiles and a long K
    a = rng.standard_normal((K, M)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    g1 = gemm(a, b, "tf32", True, False)
    g2 = gemm(a, b, "tf32", True, False)
    assert np.array_equal(g1, g2)
    assert rel_err(g1, gemm_ref(a, b, True, False)) < TF32_TOL


def test_tf32_epilogue_fusions():
    rng = np.random.default_rng(12)
    M, N, K = 300, 264, 128
    a, b = rng.standard_normal((M, K)).astype(np.float32), rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    gate = rng.standard_normal((M, N)).astype(np.float32)
    res = rng.standard_normal((M, N)).astype(np.float32)
    acc = rng.standard_normal((M, N)).astype(np.float32)
    assert rel_err(gemm(a, b, "tf32", bias=bias, relu=True), gemm_ref(a, b, bias=bias, relu=True)) < TF32_TOL
    assert rel_err(gemm(a, b, "tf32", gate=gate, residual=res), gemm_ref(a, b, gate=gate, residual=res)) < TF32_TOL
    assert rel_err(gemm(a, b, "tf32", accumulate_into=acc), gemm_ref(a, b, acc=acc)) < TF32_TOL


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False)])
def test_tf32x3_reaches_fp32_accuracy(ta, tb):
    rng = np.random.default_rng(13)
    M, N, K = 260, 200, 516
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    got = gemm(a, b, "tf32x3", ta, tb, bias=bias, relu=True)
    assert rel_err(got, gemm_ref(a, b, ta, tb, bias=bias, relu=True)) < X3_TOL
