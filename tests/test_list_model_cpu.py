"""The kernel's candidate-list algorithm, modelled step by step on the CPU (tests/helpers/list_model.py: one sorted list
with 'expanded' flags, incremental threshold position, cursor, batched rank merge, strict-drop rule), reproduces the
reference's two-heap search (src/index/mod.rs:962-1037) exactly — ids, distance bits and the n_dist / n_expand counters
— and reports an overflow instead of a wrong answer when equal distances make a bounded list insufficient."""
import numpy as np
import pytest

from helpers.data import build_fixture, random_vectors
from helpers.list_model import Overflow, search_layer_model, search_layer_model_v2


def _bits(x):
    return int(np.float32(x).view(np.uint32))


def _model_search(model, g, el, query, ef, cap, **kw):
    """Granne::search (find_entrypoint + bottom layer) with `model` as search_for_neighbors."""
    memo = {}

    def dist_bits(i):
        if i not in memo:
            memo[i] = _bits(el.dist_to_element(i, query))
        return memo[i]

    entry, n_dist, n_expand = 0, 0, 0
    last = g.num_layers() - 1
    for layer in range(g.num_layers()):
        memo_layer = {}
        nbrs = lambda i, layer=layer: memo_layer.setdefault(i, g.get_neighbors(i, layer))  # noqa: E731
        res, nd, ne = model(nbrs, dist_bits, entry, ef if layer == last else 1, cap if layer == last else 32, **kw)
        n_dist, n_expand = n_dist + nd, n_expand + ne
        if layer < last:
            entry = res[0][0]
    return res, n_dist, n_expand


@pytest.mark.parametrize("kind,n,dim,m", [("angular", 600, 8, 8), ("angular_int", 500, 12, 12), ("angular", 300, 3, 20)])
@pytest.mark.parametrize("ef,cap", [(1, 32), (5, 32), (20, 96), (80, 96), (40, 224)])
def test_model_equals_the_two_heap_search(oracle, kind, n, dim, m, ef, cap):
    el, g, _, _, _ = build_fixture(oracle, kind, n, dim, seed=n, num_neighbors=m, max_search=20, layer_multiplier=6.0)
    queries = random_vectors(12, dim, seed=ef)
    ref = g.search_batch(queries, ef, ef, with_stats=True)
    for qi in range(queries.shape[0]):
        for model, kw in ((search_layer_model, {}), (search_layer_model_v2, {"batch": 32}), (search_layer_model_v2, {"batch": 5})):
            try:
                res, n_dist, n_expand = _model_search(model, g, el, queries[qi], ef, cap, **kw)
            except Overflow:
                continue  # legal: the kernel hands such a query to the slow path
            c = int(ref[2][qi])
            assert [r[0] for r in res[:c]] == ref[0][qi, :c].tolist()
            assert [r[1] for r in res[:c]] == [_bits(x) for x in ref[1][qi, :c]]
            assert (n_dist, n_expand) == (int(ref[3][qi, 0]), int(ref[3][qi, 1]))


def test_ties_are_exact_or_reported(oracle):
    # many duplicated vectors: equal distances everywhere.  The model may only answer exactly or raise Overflow.
    base = random_vectors(40, 6, seed=3)
    raw = np.repeat(base, 10, axis=0)                                  # 400 vectors, 10 copies of each
    el = oracle.Elements.angular(raw)
    g = oracle.GranneBuilder(el, num_neighbors=10, max_search=20, layer_multiplier=6.0).build()
    queries = np.concatenate([base[:6], random_vectors(6, 6, seed=4)])
    exact = overflow = 0
    for ef, cap in ((5, 32), (12, 32), (25, 96)):
        ref = g.search_batch(queries, ef, ef, with_stats=True)
        for qi in range(queries.shape[0]):
            try:
                res, n_dist, n_expand = _model_search(search_layer_model_v2, g, el, queries[qi], ef, cap, batch=32)
            except Overflow:
                overflow += 1
                continue
            exact += 1
            c = int(ref[2][qi])
            assert [r[0] for r in res[:c]] == ref[0][qi, :c].tolist()
            assert (n_dist, n_expand) == (int(ref[3][qi, 0]), int(ref[3][qi, 1]))
    assert exact > 0
