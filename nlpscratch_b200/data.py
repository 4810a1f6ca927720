"""Dataset construction on the device: the step in front of ``train_transformer``.

Mirrors ``generate_sequence_data`` (Transformer/transformer_train.py:47-60) and
``generate_bpe_sequence_data`` (Transformer/transformer_train_bpe.py:52-65): same argument meaning,
same (X, Y) contents bit for bit, but X and Y are built by CUDA kernels and stay in HBM
(``DeviceArray``; ``.get()`` gives the reference's NumPy arrays).  Host work left: the string -> id
dictionary lookups, once per token (word level) or once per word TYPE (BPE, the reference's own
``encode_word`` cache).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .array import DeviceArray


def _dev_i64(a) -> DeviceArray:
    """int64 host vector -> device (stored as int32 pairs; DeviceArray has no 64-bit dtype)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    return DeviceArray.from_host(a.view(np.int32).reshape(-1, 2))


def _ragged(seqs):
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.zeros(max(int(off[-1]), 1), dtype=np.int32)
    if off[-1]:
        flat[:off[-1]] = np.concatenate([np.asarray(s, dtype=np.int32) for s in seqs if len(s)])
    return flat, off, lens


def prefix_dataset(ids_flat: DeviceArray, sent_off: np.ndarray, max_len: int, pad_token: int = 0):
    """All proper prefixes of ragged device sentences -> (X (n, max_len) int32, Y (n,) int32) on the device."""
    _lib.require_device()
    sent_off = np.asarray(sent_off, dtype=np.int64)
    lens = np.diff(sent_off)
    row_off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(np.maximum(lens - 1, 0), out=row_off[1:])
    n_rows = int(row_off[-1])
    X = DeviceArray.empty((n_rows, int(max_len)), np.int32)
    Y = DeviceArray.empty((n_rows,), np.int32)
    if n_rows:
        d_so, d_ro = _dev_i64(sent_off), _dev_i64(row_off)
        _lib.call("nlps_prefix_dataset", _lib.stream_ptr(), C.c_int(len(lens)), C.c_int64(n_rows), C.c_int(int(max_len)),
                  C.c_int32(int(pad_token)), _lib.vp(ids_flat.ptr), _lib.vp(d_so.ptr), _lib.vp(d_ro.ptr),
                  _lib.vp(X.ptr), _lib.vp(Y.ptr))
    return X, Y


def generate_sequence_data(processed_data, word_to_index, max_len):
    """transformer_train.py:47-60 -- ``processed_data`` is the reader()'s list of [(word, tag), ...]."""
    unk = word_to_index["<UNKNOWN>"]
    seqs = [[word_to_index.get(word, unk) for word, _tag in sentence] for sentence in processed_data]
    flat, off, _ = _ragged(seqs)
    return prefix_dataset(DeviceArray.from_host(flat), off, max_len, 0)


def build_bpe_table(words, bpe):
    """Per word TYPE: the subword ids ``encode_ids`` would emit for it (bpe_encoder.py:188-249)."""
    unk_id = bpe.token_to_id.get(bpe.special.unk, 1)
    tab = [[bpe.token_to_id.get(t, unk_id) for t in bpe.encode_word(w)] for w in words]
    flat, off, _ = _ragged(tab)
    return off.astype(np.int32), flat


def expand_tokens(word_ids_flat: DeviceArray, sent_off: np.ndarray, tab_off: np.ndarray, tab_ids: np.ndarray,
                  bos: int = -1, eos: int = -1):
    """encode_ids for a batch of sentences given as word-type ids: returns (ids_flat DeviceArray, offsets int64)."""
    _lib.require_device()
    n_sent = len(sent_off) - 1
    d_so = _dev_i64(sent_off)
    d_to, d_ti = DeviceArray.from_host(np.asarray(tab_off, np.int32)), DeviceArray.from_host(np.asarray(tab_ids, np.int32))
    scratch = DeviceArray.empty((max(n_sent, 1),), np.int32)
    d_oo = DeviceArray.empty((n_sent + 1, 2), np.int32)
    st = _lib.stream_ptr()
    _lib.call("nlps_expand_tokens_offsets", st, C.c_int(n_sent), _lib.vp(word_ids_flat.ptr), _lib.vp(d_so.ptr),
              _lib.vp(d_to.ptr), C.c_int32(bos), C.c_int32(eos), _lib.vp(scratch.ptr), _lib.vp(d_oo.ptr))
    out_off = np.ascontiguousarray(d_oo.get()).view(np.int64).reshape(-1)
    out = DeviceArray.empty((max(int(out_off[-1]), 1),), np.int32)
    _lib.call("nlps_expand_tokens_fill", st, C.c_int(n_sent), _lib.vp(word_ids_flat.ptr), _lib.vp(d_so.ptr),
              _lib.vp(d_to.ptr), _lib.vp(d_ti.ptr), C.c_int32(bos), C.c_int32(eos), _lib.vp(d_oo.ptr), _lib.vp(out.ptr))
    return out, out_off


def generate_bpe_sequence_data(token_seqs, bpe, max_len):
    """transformer_train_bpe.py:52-65 -- ``bpe`` is the reference's BPEEncoder (or any object with
    encode_word / token_to_id / special); BOS and EOS are added like ``encode_ids(add_bos=True, add_eos=True)``."""
    types = {}
    seqs = [[types.setdefault(bpe.lowercase and w.lower() or w, len(types)) for w in tokens] for tokens in token_seqs]
    tab_off, tab_ids = build_bpe_table(list(types.keys()), bpe)
    flat, off, _ = _ragged(seqs)
    bos = bpe.token_to_id.get(bpe.special.bos, 2)
    eos = bpe.token_to_id.get(bpe.special.eos, 3)
    ids, ids_off = expand_tokens(DeviceArray.from_host(flat), off, tab_off, tab_ids, bos, eos)
    return prefix_dataset(ids, ids_off, max_len, bpe.token_to_id.get(bpe.special.pad, 0))
