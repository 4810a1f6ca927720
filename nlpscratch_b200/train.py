"""Device-resident mirror of the reference training loop ``utils/train.py::train_transformer``
(utils/train.py:30-174): same signature, same batching / per-batch trimming / last-token objective /
clip / evaluation-and-LR-halving schedule / progress prints, but each training iteration is ONE fused
library call (``Transformer.train_step_last_token``) and the loop performs no device synchronisation
per step: sequence lengths are computed once on the host, the shuffle is a device gather, and the
running loss is read back only when it is printed.

The unmodified reference loop also runs against ``nlpscratch_b200.Transformer`` (that is the drop-in
contract); this module is the fast path for the same job.
"""
from __future__ import annotations

import time

import numpy as np

from .array import DeviceArray, b200np


def train_transformer(model, x_train, y_train, learning_rate=0.01, nepoch=100, evaluation_loss_after=5,
                      batch_size=32, print_every=1, clip_grad_norm=None, idx_to_word=None, sample_prompts=None,
                      topk=5, log=print):
    npb = model.np
    losses = []
    seen_examples = 0
    start_time = time.time()
    x_host = np.asarray(x_train.get() if isinstance(x_train, DeviceArray) else x_train).astype(np.int32)
    y_host = np.asarray(y_train.get() if isinstance(y_train, DeviceArray) else y_train).astype(np.int32)
    lens_host = (x_host != 0).sum(axis=1).astype(np.int32)        # once; rows are only permuted afterwards
    x_dev, y_dev = npb.array(x_host), npb.array(y_host)
    N = len(y_host)
    stats = {"steps": 0, "tokens": 0, "train_seconds": 0.0}

    def compute_eval_loss():                                      # utils/train.py:51-75
        total_loss, total_count = 0.0, 0
        model.eval()
        for i in range(0, N, batch_size):
            lens = lens_host[i:i + batch_size]
            if lens.size == 0:
                continue
            s_max = int(lens.max())
            xb, yb = x_dev[i:i + batch_size][:, :s_max], y_dev[i:i + batch_size]
            H, _ = model.forward(xb, return_hidden_only=True)
            H_last = H[npb.arange(len(lens)), npb.array(lens - 1), :]
            logits_last = H_last @ model.W_vocab + model.b_vocab
            loss, _ = model.calculate_loss_dlogits(logits_last, yb)
            total_loss += float(loss) * len(lens)
            total_count += len(lens)
        model.train()
        return total_loss / max(1, total_count)

    for epoch in range(nepoch):
        if epoch % evaluation_loss_after == 0:
            eval_loss = compute_eval_loss()
            losses.append((seen_examples, eval_loss))
            log(f"Eval loss={eval_loss:.6f} | ppl={np.exp(eval_loss):.2f} | seen={seen_examples} | epoch={epoch} | "
                f"{time.time() - start_time:.1f}s")
            if len(losses) > 1 and eval_loss > losses[-2][1]:
                learning_rate *= 0.5
                log(f"Learning rate decreased to: {learning_rate}")
        perm = np.random.permutation(N)                           # same generator as utils/train.py:98
        x_dev, y_dev = x_dev[npb.array(perm)], y_dev[npb.array(perm)]
        lens_host = lens_host[perm]
        t0 = time.time()
        token_done, last_loss = 0, None
        for bi, i in enumerate(range(0, N, batch_size)):
            lens = lens_host[i:i + batch_size]
            s_max = int(lens.max())
            xb, yb = x_dev[i:i + batch_size][:, :s_max], y_dev[i:i + batch_size]
            last_loss = model.train_step_last_token(xb, yb, lens=lens, learning_rate=learning_rate,
                                                    clip_grad_norm=clip_grad_norm)
            seen_examples += len(lens)
            token_done += int(lens.sum())
            stats["steps"] += 1
            if print_every and (bi % print_every == 0) and log is not print:
                log(f"epoch {epoch + 1} batch {bi} loss {float(last_loss):.4f} tokens {token_done}")
        model.synchronize()
        stats["tokens"] += token_done
        stats["train_seconds"] += time.time() - t0
        if idx_to_word is not None and sample_prompts:            # utils/train.py:143-163
            log("\nSamples:")
            for prompt in sample_prompts:
                seq = npb.array(prompt, dtype=npb.int32)[None, :]
                H, _ = model.forward(seq, return_hidden_only=True)
                next_logits = H[0, -1] @ model.W_vocab + model.b_vocab
                probs = npb.exp(next_logits - next_logits.max())
                probs /= probs.sum()
                top_idx = np.array(npb.asnumpy(probs)).argsort()[-topk:][::-1]
                log("  -> " + " ".join(idx_to_word.get(int(j), "<UNK>") for j in top_idx))
    log(f"Training completed in {time.time() - start_time:.2f} sec")
    stats["losses"] = losses
    return stats
