"""Device-resident mirror of the reference training loop ``utils/train.py::train_transformer``
(utils/train.py:30-174): same signature, same batching / per-batch trimming / last-token objective /
clip / evaluation-and-LR-halving schedule / progress prints, but each training iteration is ONE fused
library call (``Transformer.train_step_last_token``) and the loop performs no device synchronisation
per step: sequence lengths are computed once on the host, the shuffle is a device gather, and the
running loss is read back only when it is printed.  The evaluation pass accumulates its loss on the
device (``eval_batch_last_token``) and is read back ONCE per pass together with the learning rate the
device-side schedule decided (``eval_finish``: halve when the loss went up); the top-k sampling block
runs its softmax and selection on the device (``topk_next``).

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

    def bucket(lens):
        # utils/train.py:108-110 trims a batch to its longest sequence; here the trim is rounded up to the next multiple of
        # 8 columns so that an epoch has ~10 batch shapes (one activation cache and one CUDA graph each) instead of ~60.
        # The extra columns hold PAD ids: masked keys contribute exact zeros and their (zero) gradients add exact zeros.
        # (An all-PAD row -- never produced by the prefix datasets -- reads position s_max-1 through t = -1, so a batch
        # that holds one keeps the exact trim.)
        s_max = int(lens.max())
        if int(lens.min()) >= 1:
            s_max = min(x_host.shape[1], (s_max + 7) // 8 * 8)
        return s_max

    accum = DeviceArray.from_host(np.zeros(2, np.float32))
    sched = DeviceArray.from_host(np.array([0.0, 0.0, learning_rate, 0.0, 0.0], np.float32))   # prev, passes, lr, halved, eval

    def compute_eval_loss():                                      # utils/train.py:51-75: one read-back per pass
        for i in range(0, N, batch_size):
            lens = lens_host[i:i + batch_size]
            if lens.size == 0:
                continue
            s_max = bucket(lens)
            model.eval_batch_last_token(x_dev[i:i + batch_size][:, :s_max], y_dev[i:i + batch_size], accum, lens=lens)
        return model.eval_finish(accum, sched)                    # (loss, lr after the schedule, halved)

    for epoch in range(nepoch):
        if epoch % evaluation_loss_after == 0:
            eval_loss, learning_rate, halved = compute_eval_loss()
            losses.append((seen_examples, eval_loss))
            log(f"Eval loss={eval_loss:.6f} | ppl={np.exp(eval_loss):.2f} | seen={seen_examples} | epoch={epoch} | "
                f"{time.time() - start_time:.1f}s")
            if halved:                                             # utils/train.py:86-88, decided on the device
                log(f"Learning rate decreased to: {learning_rate}")
        perm = np.random.permutation(N)                           # same generator as utils/train.py:98
        x_dev, y_dev = x_dev[npb.array(perm)], y_dev[npb.array(perm)]
        lens_host = lens_host[perm]
        t_dev = npb.array(lens_host - 1)                          # the epoch's last-token indices: one copy, sliced below
        t0 = time.time()
        token_done, last_loss = 0, None
        for bi, i in enumerate(range(0, N, batch_size)):
            lens = lens_host[i:i + batch_size]
            s_max = bucket(lens)
            xb, yb = x_dev[i:i + batch_size][:, :s_max], y_dev[i:i + batch_size]
            last_loss = model.train_step_last_token(xb, yb, t_index=t_dev[i:i + batch_size], learning_rate=learning_rate,
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
                top_idx, _ = model.topk_next(prompt, topk)
                log("  -> " + " ".join(idx_to_word.get(int(j), "<UNK>") for j in top_idx))
    log(f"Training completed in {time.time() - start_time:.2f} sec")
    stats["losses"] = losses
    return stats
