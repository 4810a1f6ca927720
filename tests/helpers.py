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
