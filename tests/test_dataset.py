"""Dataset construction: oracle vs the reference's golden fixture on CPU, CUDA kernels vs
both on the GPU -- bit-exact (integer work)."""
import hashlib
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import dataset_oracle as O


def _gold():
    with np.load(os.path.join(GOLDEN, "dataset_mini.npz")) as z:
        return {k: z[k] for k in z.files}


def _split(flat, off):
    return [flat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# hashes recorded from the reference in SURVEY.md 8c
SURVEY_SHA16 = {"X": "089bfd771fb38525", "Y": "0b9d2b058ba3b9f2", "Xb": "062ea96e2a3af2ef", "Yb": "96760d3f4ad8cd05"}


def test_fixture_matches_survey_hashes():
    g = _gold()
    for k, want in SURVEY_SHA16.items():
        assert _sha(g[k])[:16] == want
    assert g["X"].shape == (1434, 320) and g["Xb"].shape == (2188, 320)
    assert int(g["vocab"]) == 574 and int(g["bpe_vocab"]) == 540 and int(g["bpe_merges"]) == 573


def test_oracle_pad_sequence_known_answers():
    assert O.pad_sequence([1, 2, 3], 5) == [1, 2, 3, 0, 0]
    assert O.pad_sequence([1, 2, 3, 4, 5, 6], 4) == [3, 4, 5, 6]
    assert O.pad_sequence([], 3, 7) == [7, 7, 7]


def test_oracle_reproduces_reference_datasets():
    g = _gold()
    X, Y = O.prefix_dataset(_split(g["word_ids"], g["word_off"]), int(g["max_len"]))
    assert np.array_equal(X, g["X"]) and np.array_equal(Y, g["Y"])
    pad, _unk, bos, eos = (int(v) for v in g["bpe_special"])
    ids = O.expand_tokens(_split(g["bpe_words"], g["bpe_off"]), g["bpe_tab_off"], g["bpe_tab_ids"], bos, eos)
    assert ids[0][:8] == [2, 250, 126, 252, 242, 126, 448, 4]      # sentence 0 of the BPE corpus
    Xb, Yb = O.prefix_dataset(ids, int(g["max_len"]), pad)
    assert np.array_equal(Xb, g["Xb"]) and np.array_equal(Yb, g["Yb"])


@pytest.mark.gpu
def test_gpu_prefix_dataset_matches_reference_bit_for_bit():
    from nlpscratch_b200 import data
    from nlpscratch_b200.array import DeviceArray
    g = _gold()
    X, Y = data.prefix_dataset(DeviceArray.from_host(g["word_ids"]), g["word_off"], int(g["max_len"]), 0)
    assert _sha(X.get()) == str(g["X_sha"]) and _sha(Y.get()) == str(g["Y_sha"])


@pytest.mark.gpu
def test_gpu_bpe_expand_and_prefix_match_reference_bit_for_bit():
    from nlpscratch_b200 import data
    from nlpscratch_b200.array import DeviceArray
    g = _gold()
    pad, _unk, bos, eos = (int(v) for v in g["bpe_special"])
    ids, off = data.expand_tokens(DeviceArray.from_host(g["bpe_words"]), g["bpe_off"], g["bpe_tab_off"], g["bpe_tab_ids"], bos, eos)
    want = O.expand_tokens(_split(g["bpe_words"], g["bpe_off"]), g["bpe_tab_off"], g["bpe_tab_ids"], bos, eos)
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(s) for s in want])]).tolist()
    assert ids.get()[:off[-1]].tolist() == [v for s in want for v in s]
    Xb, Yb = data.prefix_dataset(ids, off, int(g["max_len"]), pad)
    assert _sha(Xb.get()) == str(g["Xb_sha"]) and _sha(Yb.get()) == str(g["Yb_sha"])


@pytest.mark.gpu
@pytest.mark.parametrize("max_len", [4, 7, 33, 320])
def test_gpu_prefix_dataset_ragged_edges(max_len):
    # empty and one-id sentences contribute no rows; prefixes longer than max_len keep their last max_len ids;
    # odd max_len takes the scalar store path
    from nlpscratch_b200 import data
    from nlpscratch_b200.array import DeviceArray
    rng = np.random.default_rng(max_len)
    seqs = [[], [5], rng.integers(1, 99, 2).tolist(), rng.integers(1, 99, 40).tolist(), [], rng.integers(1, 99, 9).tolist(),
            rng.integers(1, 99, 700).tolist(), [3]]
    off = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.int64)
    flat = np.array([v for s in seqs for v in s], dtype=np.int32)
    X, Y = data.prefix_dataset(DeviceArray.from_host(flat), off, max_len, 0)
    Xw, Yw = O.prefix_dataset(seqs, max_len, 0)
    assert np.array_equal(X.get(), Xw) and np.array_equal(Y.get(), Yw)


@pytest.mark.gpu
def test_gpu_generate_bpe_sequence_data_with_a_stand_in_encoder():
    # the reference-facing call: any object with encode_word / token_to_id / special / lowercase works
    from nlpscratch_b200 import data

    class Special:
        pad, unk, bos, eos = "<PAD>", "<UNK>", "<BOS>", "<EOS>"

    class Enc:
        special, lowercase = Special(), False
        token_to_id = {"<PAD>": 0, "<UNK>": 1, "<BOS>": 2, "<EOS>": 3, "ab": 4, "c</w>": 5, "d": 6, "e</w>": 7}

        def encode_word(self, w):
            return {"abc": ["ab", "c</w>"], "de": ["d", "e</w>"], "zz": ["z", "z</w>"]}[w]

    toks = [["abc", "de"], ["zz"], ["de", "de", "abc"]]
    X, Y = data.generate_bpe_sequence_data(toks, Enc(), 6)
    ids = [[2, 4, 5, 6, 7, 3], [2, 1, 1, 3], [2, 6, 7, 6, 7, 4, 5, 3]]
    Xw, Yw = O.prefix_dataset(ids, 6, 0)
    assert np.array_equal(X.get(), Xw) and np.array_equal(Y.get(), Yw)
