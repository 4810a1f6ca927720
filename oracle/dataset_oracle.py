"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's dataset construction.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product
path (nlpscratch_b200/data.py) never does.

Restates, in plain Python/NumPy:
  * pad_sequence                      Transformer/transformer.py:575-582
  * generate_sequence_data            Transformer/transformer_train.py:47-60
  * generate_bpe_sequence_data        Transformer/transformer_train_bpe.py:52-65
  * BPEEncoder.encode_ids (table form) Transformer/bpe_encoder.py:225-249
Pinned to the real reference by tests/golden/dataset_mini.npz (made by tests/golden/make_dataset_golden.py,
which imports /root/reference) and by the hashes recorded in SURVEY.md 8c.
"""
import numpy as np


def pad_sequence(seq, max_len, pad_token=0):
    seq = list(seq)
    if len(seq) >= max_len:
        return seq[len(seq) - max_len:]          # keep the most recent context
    return seq + [pad_token] * (max_len - len(seq))


def prefix_dataset(id_seqs, max_len, pad_token=0):
    X, Y = [], []
    for ids in id_seqs:
        for i in range(1, len(ids)):
            X.append(pad_sequence(ids[:i], max_len, pad_token))
            Y.append(ids[i])
    return (np.array(X, dtype=np.int32).reshape(len(X), max_len), np.array(Y, dtype=np.int32))


def expand_tokens(word_id_seqs, tab_off, tab_ids, bos=-1, eos=-1):
    out = []
    for words in word_id_seqs:
        ids = [bos] if bos >= 0 else []
        for w in words:
            ids.extend(int(v) for v in tab_ids[tab_off[w]:tab_off[w + 1]])
        if eos >= 0:
            ids.append(eos)
        out.append(ids)
    return out
