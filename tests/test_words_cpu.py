"""Host-side containers of the reference's Python module (granne_b200/words.py): WordDict (py/src/variants/mod.rs:8-78),
Embeddings (py/src/embeddings.rs:8-144) and the file readers behind string queries.  CPU only."""
import numpy as np
import pytest

import granne_b200
from granne_b200 import words as W
from helpers.data import random_sum_embeddings, random_vectors


def test_word_dict_round_trip_and_lookup(tmp_path):
    d = W.WordDict()
    for w in ["hello", "wörld", 'quo"te', "hello", "tab\tword"]:
        d.push(w)
    assert len(d) == 4 and d.push("new") and not d.push("new")
    path = str(tmp_path / "words.txt")
    d.write(path)
    e = W.WordDict(path)
    assert e.id_to_word == d.id_to_word
    assert e.get_word_ids("hello unknown  wörld\nnew") == [0, 1, 4]      # unknown words are dropped (:53-58)
    assert e.get_words([1, 0]) == "wörld hello" and e.get_words([]) == ""
    with open(path, "a", encoding="utf-8") as f:
        f.write('"hello"\n')                                              # duplicate: the last id wins (:26)
    assert W.WordDict(path).get_word_ids("hello") == [5]


def test_embeddings_container_matches_the_oracle(oracle, tmp_path):
    emb = random_vectors(40, 12, seed=5)
    e = granne_b200.Embeddings()
    for i, row in enumerate(emb):
        assert e.append(row.tolist(), "w%d" % i)
    assert not e.append(emb[0].tolist(), "w3") and len(e) == 40           # existing word: nothing stored (:109-116)
    ids = [3, 17, 3, 39, 0]
    want = emb[3].copy()
    for w in ids[1:]:
        want = oracle.sum_into_f32(want, emb[w])                          # math.rs:92-116, in order
    assert np.array_equal(np.asarray(e.get_embedding(ids), dtype=np.float32), want)
    assert e.get_embedding("w3 nope w17 w3 w39 w0") == e.get_embedding(ids)
    assert e.get_embedding(7) == emb[7].tolist() and e.get_embedding([]) == [0.0] * 12
    ep, wp = str(tmp_path / "emb.bin"), str(tmp_path / "words.txt")
    e.save(ep, wp)
    el = oracle.Elements.sum_embeddings(emb, [[0, 1]])
    assert open(ep, "rb").read() == el.to_bytes(1)                        # same file as the oracle's writer
    f = granne_b200.Embeddings(ep, wp)
    assert len(f) == 40 and f.get_embedding("w5 w6") == e.get_embedding([5, 6])
    with pytest.raises(ValueError):
        granne_b200.Embeddings(ep, None)


def test_file_readers_match_the_oracle(oracle):
    el = random_sum_embeddings(oracle, 10, 120, 300, seed=2)
    table = W.read_dense_f32(el.to_bytes(1))
    assert np.array_equal(table, el.rows())
    terms = W.read_sum_terms(el.to_bytes(0))
    assert len(terms) == 300 and all(terms[i] == el.terms(i) for i in range(300))
    for i in (0, 7, 299):  # ElementContainer::get = normalised create_embedding (embeddings/mod.rs:164-166)
        assert np.array_equal(oracle.normalize_f32(W.create_embedding(table, terms[i])), el.get(i))
    with pytest.raises(ValueError):
        W.read_sum_terms(el.to_bytes(0)[:20])
    with pytest.raises(ValueError):
        W.read_dense_f32(b"\x03\x00\x00\x00\x00\x00\x00\x00" + b"\x00" * 16)
