"""Golden trajectory of the reference's OWN training loop on its OWN dataset (BASELINE.json configs[0]).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_driver_golden.py

Runs utils/train.py::train_transformer (unmodified) with the reference Transformer and the default hyper-parameters of
Transformer/transformer_train.py:28-44 (V574 E256 h8 F1024 L3 max_len 320, B16, Adam lr 1e-4, clip 1.0) on the word-level
prefix dataset of dataset/tagged_train_mini.txt (tests/golden/dataset_mini.npz holds X, Y), dropout 0 so that the GPU
path can be compared step for step, 2 epochs, evaluation every epoch.  Stores the printed evaluation losses, a further
evaluation after the last epoch, and samples of the final parameters.
"""
import contextlib
import io
import json
import os
import re
import sys
import time

import numpy as np

REF = os.environ.get("NLPS_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(1, os.path.dirname(HERE))

from Transformer.transformer import Transformer  # noqa: E402
from utils.train import train_transformer  # noqa: E402
from helpers import NAMES, sampled  # noqa: E402

SEED, NEPOCH = 2024, 2


def final_eval(model, X, Y, batch_size=16):
    """compute_eval_loss of utils/train.py:51-75 (it is a closure there), for the state after the last epoch."""
    npb = model.np
    tot, cnt = 0.0, 0
    model.eval()
    for i in range(0, len(Y), batch_size):
        xb, yb = X[i:i + batch_size], Y[i:i + batch_size]
        lens = npb.sum(xb != 0, axis=1).astype(npb.int32)
        xb = xb[:, :int(npb.max(lens))]
        H, _ = model.forward(xb, return_hidden_only=True)
        logits = H[npb.arange(len(yb)), lens - 1, :] @ model.W_vocab + model.b_vocab
        loss, _ = model.calculate_loss_dlogits(logits, yb)
        tot += float(loss) * len(yb)
        cnt += len(yb)
    model.train()
    return tot / cnt


def main():
    with np.load(os.path.join(HERE, "dataset_mini.npz")) as z:
        X, Y, V = z["X"], z["Y"], int(z["vocab"])
    np.random.seed(SEED)
    model = Transformer(V, 256, 8, 1024, 3, 320, dropout_p=0.0)
    buf = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
        train_transformer(model, X, Y, learning_rate=1e-4, nepoch=NEPOCH, evaluation_loss_after=1, batch_size=16,
                          print_every=10, clip_grad_norm=1.0)
    evals = [float(v) for v in re.findall(r"Eval loss=([0-9.]+)", buf.getvalue())]
    out = {"seed": SEED, "nepoch": NEPOCH, "eval_losses": evals, "final_eval": final_eval(model, X, Y),
           "seconds": round(time.time() - t0, 1), "numpy": np.__version__,
           "params_sampled": {k: [float(v) for v in sampled(getattr(model, k), 64)] for k in NAMES}}
    with open(os.path.join(HERE, "driver_mini.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps({k: out[k] for k in ("eval_losses", "final_eval", "seconds")}))


if __name__ == "__main__":
    main()
