"""Generate the golden fixtures in this directory from the REAL reference.

Run once in the build container (the reference is mounted read-only at
/root/reference and is pure Python/NumPy, so it imports here):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The GPU box has no /root/reference; tests only read the committed .npz/.json
files.  numpy version used is recorded in manifest.json.
"""
import hashlib
import json
import os
import sys

import numpy as np

REF = os.environ.get("NLPS_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from Transformer.transformer import Transformer, pad_sequence  # noqa: E402
from utils.function import clip_grads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["We", "Wq", "Wk", "Wv", "Wo", "W1", "W2", "b1", "b2", "W_vocab", "b_vocab",
         "ln1_gamma", "ln1_beta", "ln2_gamma", "ln2_beta"]


def params_of(model):
    return {n: np.array(getattr(model, n)) for n in NAMES}


def perturb(model, rng):
    """Make biases / LN params non-trivial so their paths are exercised."""
    for n in ("b1", "b2", "b_vocab", "ln1_beta", "ln2_beta"):
        a = getattr(model, n)
        a[...] = rng.uniform(-0.2, 0.2, a.shape).astype(np.float32)
    for n in ("ln1_gamma", "ln2_gamma"):
        a = getattr(model, n)
        a[...] = rng.uniform(0.5, 1.5, a.shape).astype(np.float32)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    return os.path.getsize(path)


def case_full_seq(name, cfg, x, y, steps=2, lr=1e-3, use_adam=True, seed=0):
    np.random.seed(seed)
    model = Transformer(*cfg, dropout_p=0.0, use_adam=use_adam)
    perturb(model, np.random.default_rng(seed + 1))
    out = {"cfg": np.array(cfg), "x": x, "y": y, "lr": np.array(lr), "use_adam": np.array(use_adam)}
    out.update({"p0_" + k: v for k, v in params_of(model).items()})
    for s in range(steps):
        logits, cache = model.forward(x)
        loss, dlogits = model.calculate_loss_dlogits(logits, y)
        grads = model.backward(cache, dlogits)
        if s == 0:
            out["logits"] = np.array(logits)
            out["dlogits"] = np.array(dlogits)
            out["H_final"] = np.array(cache["H_final"])
            out["attn0"] = np.array(cache["layers"][0]["attn_weights"])
            out.update({"g_" + k: np.array(v) for k, v in grads.items()})
        out["loss%d" % s] = np.array(loss)
        model.step(grads, lr)
        out.update({"p%d_%s" % (s + 1, k): v for k, v in params_of(model).items()})
    return save(name, **out)


def case_last_token(name, cfg, xfull, y, steps=2, lr=1e-3, clip=1.0, seed=0):
    """Replays the body of utils/train.py:104-131 against the reference model."""
    np.random.seed(seed)
    model = Transformer(*cfg, dropout_p=0.0)
    perturb(model, np.random.default_rng(seed + 1))
    out = {"cfg": np.array(cfg), "x": xfull, "y": y, "lr": np.array(lr), "clip": np.array(clip)}
    out.update({"p0_" + k: v for k, v in params_of(model).items()})
    npb = model.np
    for s in range(steps):
        xb = xfull
        lens = npb.sum(xb != 0, axis=1).astype(npb.int32)
        s_max = int(npb.max(lens))
        xb = xb[:, :s_max]
        H, cache = model.forward(xb, return_hidden_only=True)
        idx = npb.arange(len(y))
        t = lens - 1
        logits_last = H[idx, t, :] @ model.W_vocab + model.b_vocab
        loss, dlog = model.calculate_loss_dlogits(logits_last, y)
        grads = model.backward_last_token(cache, t, dlog[:, 0, :])
        if s == 0:
            out["H"] = np.array(H)
            out["logits_last"] = np.array(logits_last)
            out["dlog"] = np.array(dlog)
            out.update({"graw_" + k: np.array(v) for k, v in grads.items()})
        clip_grads(list(grads.values()), clip, lib=npb)
        if s == 0:
            out.update({"g_" + k: np.array(v) for k, v in grads.items()})
        out["loss%d" % s] = np.array(loss)
        model.step(grads, lr)
        out.update({"p%d_%s" % (s + 1, k): v for k, v in params_of(model).items()})
    return save(name, **out)


def case_smoke_main():
    """The reference's own __main__ smoke (transformer.py:618-641)."""
    res = {}
    for p in (0.1, 0.0):
        np.random.seed(42)
        model = Transformer(1000, 64, 8, 256, 2, 50, dropout_p=p)
        x = np.array([[1, 2, 3, 4, 5], [5, 4, 3, 2, 1]], dtype=np.int32)
        y = np.array([[2, 3, 4, 5, 6], [1, 1, 1, 1, 1]], dtype=np.int32)
        logits, cache = model.forward(x)
        loss, dlogits = model.calculate_loss_dlogits(logits, y)
        grads = model.backward(cache, dlogits)
        res["loss_p%g" % p] = float(loss)
        if p == 0.0:
            res["logits_0_0_3"] = [float(v) for v in logits[0, 0, :3]]
            res["grad_norms"] = {k: float(np.sqrt(np.sum(np.square(v.astype(np.float64)))))
                                 for k, v in grads.items()}
            res["param_sha256_16"] = {k: hashlib.sha256(np.ascontiguousarray(v).tobytes()).hexdigest()[:16]
                                      for k, v in params_of(model).items()}
            res["pe_1_4"] = [float(v) for v in model.pe[1, :4]]
    res["pad_sequence"] = {"short": pad_sequence([1, 2, 3], 5), "long": pad_sequence([1, 2, 3, 4, 5, 6], 4)}
    return res


def case_dataset_hashes():
    """Word-level and BPE prefix datasets on the mini corpus (transformer_train.py:47-60)."""
    from utils.utils import read_file_to_list, reader, count_word_POS, build_vocab
    res = {}
    lines = read_file_to_list(os.path.join(REF, "dataset", "tagged_train_mini.txt"))
    proc = reader(lines)
    pos_cnt, word_cnt = count_word_POS(proc)
    w2i, _ = build_vocab(word_cnt, pos_cnt)
    X, Y = [], []
    for sent in proc:
        ids = [w2i.get(w, w2i["<UNKNOWN>"]) for w, _t in sent]
        for i in range(1, len(ids)):
            X.append(pad_sequence(ids[:i], 320))
            Y.append(ids[i])
    X = np.array(X, dtype=np.int32)
    Y = np.array(Y, dtype=np.int32)
    res["word"] = {"vocab": len(w2i), "X_shape": list(X.shape),
                   "X_sha": hashlib.sha256(X.tobytes()).hexdigest()[:16],
                   "Y_sha": hashlib.sha256(Y.tobytes()).hexdigest()[:16],
                   "nonpad_tokens": int((X != 0).sum())}
    return res


def main():
    rng = np.random.default_rng(7)
    sizes = {}
    cfg = (50, 32, 4, 64, 2, 16)           # V E h F L max_len
    x = rng.integers(1, 50, (3, 10)).astype(np.int32)
    y = rng.integers(0, 50, (3, 10)).astype(np.int32)
    sizes["full_seq_small"] = case_full_seq("full_seq_small", cfg, x, y)
    sizes["full_seq_sgd"] = case_full_seq("full_seq_sgd", cfg, x, y, steps=1, lr=1e-2, use_adam=False, seed=3)

    # tail padding with ragged lengths, incl. PAD targets in y
    xp = rng.integers(1, 50, (4, 12)).astype(np.int32)
    for b, n in enumerate((12, 7, 3, 1)):
        xp[b, n:] = 0
    yp = rng.integers(0, 50, (4, 12)).astype(np.int32)
    sizes["full_seq_padded"] = case_full_seq("full_seq_padded", cfg, xp, yp, steps=1, seed=5)

    # leading PAD => fully-masked query rows (softmax over raw scores of ALL keys)
    xd = rng.integers(1, 50, (3, 8)).astype(np.int32)
    xd[0, 0] = 0
    xd[1, :3] = 0
    xd[2, 5:] = 0
    yd = rng.integers(0, 50, (3, 8)).astype(np.int32)
    sizes["full_seq_leading_pad"] = case_full_seq("full_seq_leading_pad", cfg, xd, yd, steps=1, seed=9)

    # the real driver's path: trim to s_max, last-token loss, clip, Adam
    xl = np.zeros((5, 16), np.int32)
    for b, n in enumerate((9, 4, 11, 1, 6)):
        xl[b, :n] = rng.integers(1, 50, n)
    yl = rng.integers(0, 50, (5,)).astype(np.int32)
    sizes["last_token_clip"] = case_last_token("last_token_clip", cfg, xl, yl, clip=0.05)

    # head_dim 32 / odd vocab (574 like the word-level mini corpus), single layer
    cfg2 = (574, 32, 1, 64, 1, 12)
    x2 = rng.integers(1, 574, (2, 12)).astype(np.int32)
    y2 = rng.integers(0, 574, (2, 12)).astype(np.int32)
    sizes["full_seq_v574"] = case_full_seq("full_seq_v574", cfg2, x2, y2, steps=1, seed=11)

    manifest = {"numpy": np.__version__, "reference": "hydrogendeuteride/NLPScratch @ /root/reference",
                "sizes_bytes": sizes, "smoke_main": case_smoke_main(), "datasets": case_dataset_hashes()}
    with open(os.path.join(HERE, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
