"""Host side of Granne::reorder in the product library (granne_b200/csrc/reorder.hpp) against the oracle; CPU only.
The GPU half (the trail searches of compute_order) is covered by tests/test_reorder_gpu.py."""
import numpy as np
import pytest

import granne_b200
from granne_b200 import api
from helpers.data import build_fixture


@pytest.fixture(scope="module")
def lib():
    return api.load_library()


def _all_trails(index):
    n = len(index)
    trails = np.zeros((n, 8), dtype=np.uint32)
    for layer in range(1, index.num_layers()):
        for idx in range(index.layer_len(layer - 1), index.layer_len(layer)):
            trails[idx] = index.entrypoint_trail(idx, layer)
    return trails


@pytest.mark.parametrize("kind,n,dim,mult", [("angular", 3000, 6, 5.0), ("angular_int", 2000, 16, 8.0),
                                             ("embeddings", 900, 10, 4.0), ("angular", 40, 4, 15.0)])
def test_order_from_trails_and_apply_order_match_the_oracle(lib, oracle, kind, n, dim, mult):
    el, index, index_bytes, eb, mb = build_fixture(oracle, kind, n, dim, seed=n, num_neighbors=12, max_search=15,
                                                   layer_multiplier=mult)
    lens = [index.layer_len(l) for l in range(index.num_layers())]
    want_order = index.compute_order()
    got_order = api.order_from_trails(lens, _all_trails(index))            # compute_order, reorder.rs:126-174
    assert np.array_equal(want_order, got_order)
    new_index, new_elements = api.apply_order(index_bytes, kind, eb, got_order)
    ref = oracle.Granne.from_bytes(index_bytes, oracle.Elements.from_bytes(kind, eb, mb))
    ref.reorder()
    assert new_index == ref.to_bytes()                                      # reorder_layers, reorder.rs:209-292
    assert new_elements == ref.elements.to_bytes(0)                         # Permutable::permute


def test_order_by_keys_and_embedding_keys_match_the_oracle(lib, oracle):
    # reorder_sum_embeddings (embeddings/reorder.rs:79-102)
    el, index, index_bytes, eb, mb = build_fixture(oracle, "embeddings", 500, 5, seed=6, num_neighbors=30, max_search=5,
                                                   layer_multiplier=5.0, num_embeddings=277)
    keys = api.compute_keys_for_reordering(eb, mb)
    assert np.array_equal(keys, el.reorder_keys())
    order = api.order_by_keys(index_bytes, keys)
    assert np.array_equal(order, index.order_by_keys(keys))
    rng = np.random.default_rng(3)
    scalar = rng.integers(0, 20, size=500).astype(np.uint64)               # many ties: (key, idx) order
    assert np.array_equal(api.order_by_keys(index_bytes, scalar), index.order_by_keys(scalar))
    new_index, new_elements = api.apply_order(index_bytes, "embeddings", eb, order)
    ref = oracle.Granne.from_bytes(index_bytes, oracle.Elements.from_bytes("embeddings", eb, mb))
    ref.reorder_by_keys(keys)
    assert new_index == ref.to_bytes() and new_elements == ref.elements.to_bytes(0)


def test_identity_and_error_cases(lib, oracle):
    el, index, index_bytes, eb, _ = build_fixture(oracle, "angular", 300, 8, seed=2, num_neighbors=8, max_search=10,
                                                  layer_multiplier=6.0)
    ident = np.arange(300, dtype=np.uint64)
    new_index, new_elements = api.apply_order(index_bytes, "angular", eb, ident)
    assert new_index == index_bytes and new_elements == eb
    with pytest.raises(granne_b200.GranneError):                            # not a permutation
        api.apply_order(index_bytes, "angular", eb, np.zeros(300, dtype=np.uint64))
    with pytest.raises(granne_b200.GranneError):                            # wrong length (assert_eq! in permute)
        api.apply_order(index_bytes, "angular", eb, ident[:299])
    swapped = ident.copy()
    swapped[0], swapped[299] = 299, 0                                       # moves a last-layer node into layer 0
    assert index.layer_len(0) < 300
    with pytest.raises(granne_b200.GranneError):
        api.apply_order(index_bytes, "angular", eb, swapped)
    with pytest.raises(granne_b200.GranneError):                            # assert_eq!(self.len(), keys.len())
        api.order_by_keys(index_bytes, np.zeros(10, dtype=np.uint64))
