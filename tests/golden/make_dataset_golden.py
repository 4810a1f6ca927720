"""Golden fixture for the dataset-construction kernels, generated from the REAL reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_dataset_golden.py

Runs reader / build_vocab / generate_sequence_data (word level) and BPEEncoder.train(8000, min_frequency 2) /
generate_bpe_sequence_data on dataset/tagged_train_mini.txt and stores the integer inputs our kernels take
(ragged ids, the per-word-type BPE table) next to the reference's X / Y and their SHA-256.
"""
import hashlib
import os
import sys

import numpy as np

REF = os.environ.get("NLPS_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from utils.utils import read_file_to_list, reader, count_word_POS, build_vocab  # noqa: E402
from Transformer.transformer_train import generate_sequence_data  # noqa: E402
from Transformer.transformer_train_bpe import generate_bpe_sequence_data  # noqa: E402
from Transformer.bpe_encoder import BPEEncoder  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEN = 320


def ragged(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    return np.concatenate([np.asarray(s, dtype=np.int32) for s in seqs]), off


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    data = reader(read_file_to_list(os.path.join(REF, "dataset", "tagged_train_mini.txt")))
    pos_cnt, word_cnt = count_word_POS(data)
    word_to_idx, _ = build_vocab(word_cnt, pos_cnt)
    X, Y = generate_sequence_data(data, word_to_idx, MAX_LEN)
    unk = word_to_idx["<UNKNOWN>"]
    w_ids, w_off = ragged([[word_to_idx.get(w, unk) for w, _t in s] for s in data])

    token_seqs = [[w for (w, _t) in sent[1:-1]] for sent in data]
    bpe = BPEEncoder().train(token_seqs, vocab_size=8000, min_frequency=2)
    Xb, Yb = generate_bpe_sequence_data(token_seqs, bpe, MAX_LEN)
    types = {}
    b_words, b_off = ragged([[types.setdefault(w, len(types)) for w in toks] for toks in token_seqs])
    unk_id = bpe.token_to_id.get(bpe.special.unk, 1)
    tab = [[bpe.token_to_id.get(t, unk_id) for t in bpe.encode_word(w)] for w in types]
    tab_ids, tab_off = ragged(tab)
    out = dict(max_len=np.array(MAX_LEN), word_ids=w_ids, word_off=w_off, X=X, Y=Y, X_sha=np.array(sha(X)), Y_sha=np.array(sha(Y)),
               vocab=np.array(len(word_to_idx)),
               bpe_words=b_words, bpe_off=b_off, bpe_tab_ids=tab_ids, bpe_tab_off=tab_off.astype(np.int32),
               bpe_special=np.array([bpe.pad_id, bpe.unk_id, bpe.token_to_id[bpe.special.bos], bpe.token_to_id[bpe.special.eos]]),
               bpe_vocab=np.array(bpe.vocab_size), bpe_merges=np.array(len(bpe.merges)),
               Xb=Xb, Yb=Yb, Xb_sha=np.array(sha(Xb)), Yb_sha=np.array(sha(Yb)))
    path = os.path.join(HERE, "dataset_mini.npz")
    np.savez_compressed(path, **out)
    print("word level: vocab", len(word_to_idx), "X", X.shape, sha(X)[:16], "Y", sha(Y)[:16])
    print("bpe: vocab", bpe.vocab_size, "merges", len(bpe.merges), "X", Xb.shape, sha(Xb)[:16], "Y", sha(Yb)[:16])
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
