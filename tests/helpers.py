"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("We", "Wq", "Wk", "Wv", "Wo", "W1", "W2", "b1", "b2", "W_vocab", "b_vocab",
         "ln1_gamma", "ln1_beta", "ln2_gamma", "ln2_beta")
FULL_SEQ_CASES = ("full_seq_small", "full_seq_sgd", "full_seq_padded", "full_seq_leading_pad", "full_seq_v574")


def load_case(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        return json.load(fh)


def rel_err(got, want):
    """Norm-wise relative error ||got-want|| / max(||want||, tiny)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    den = max(float(np.linalg.norm(want.ravel())), 1e-30)
    return float(np.linalg.norm((got - want).ravel())) / den


def params_from_case(case, step=0):
    return {k: case["p%d_%s" % (step, k)] for k in NAMES}


# ---- large-model fixtures (tests/golden/make_golden_large.py) -------------------------------------------
# Parameters are NOT stored: the reference draws them from NumPy's global generator in a fixed order
# (transformer.py:34-55), so (seed, perturbation recipe) reproduces them bit for bit on either side; the
# fixture pins that with a SHA-256 per tensor.
LARGE_CASES = ("default_full_seq", "default_last_token", "c3_full_seq", "c4_full_seq", "c5_full_seq")


def perturb_params(params, seed):
    """Make biases / LayerNorm parameters non-trivial (they start at 0 / 1): same recipe in the golden
    generator (applied to the real reference model) and in the tests (applied to the model under test)."""
    rng = np.random.default_rng(seed)
    for n in ("b1", "b2", "b_vocab", "ln1_beta", "ln2_beta"):
        params[n][...] = rng.uniform(-0.2, 0.2, params[n].shape).astype(np.float32)
    for n in ("ln1_gamma", "ln2_gamma"):
        params[n][...] = rng.uniform(0.5, 1.5, params[n].shape).astype(np.float32)
    return params


def sample_idx(size, n=4096):
    """Deterministic spread of flat indices used to keep large tensors out of the fixtures."""
    size = int(size)
    if size <= n:
        return np.arange(size, dtype=np.int64)
    return np.unique(np.linspace(0, size - 1, n).astype(np.int64))


def sampled(a, n=4096):
    a = np.ascontiguousarray(a)
    return a.reshape(-1)[sample_idx(a.size, n)]


def sha16(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
