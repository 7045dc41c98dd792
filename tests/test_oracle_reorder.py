"""Oracle vs the reference's own reorder / permute tests (src/index/reorder.rs:294-334,
src/slice_vector/mod.rs:1028-1092, src/elements/embeddings/reorder.rs:60-102).  CPU only."""
import numpy as np
import pytest

from helpers.data import build_fixture, index_from_lists, random_sum_embeddings, random_vectors


def _element_queries(el, ids):
    return np.stack([el.get(i) for i in ids])


def _check_results_map_through_permutation(index, reordered, permutation, queries, max_search=10, k=10):
    exp = index.search_batch(queries, max_search, k, already_element=True)
    res = reordered.search_batch(queries, max_search, k, already_element=True)
    assert np.array_equal(exp[2], res[2])
    for qi in range(queries.shape[0]):
        c = int(exp[2][qi])
        assert np.array_equal(exp[0][qi, :c], permutation[res[0][qi, :c]])       # reorder.rs:317-321
        assert np.array_equal(exp[1][qi, :c].view(np.uint32), res[1][qi, :c].view(np.uint32))


def test_reorder_index(oracle):
    # reorder_index (reorder.rs:298-323): 5000 x 5, max_search 5, layer_multiplier 5
    el, index, index_bytes, eb, _ = build_fixture(oracle, "angular", 5000, 5, seed=1, num_neighbors=30, max_search=5,
                                                  layer_multiplier=5.0)
    reordered = oracle.Granne.from_bytes(index_bytes, oracle.Elements.from_bytes("angular", eb))
    permutation = reordered.reorder()
    assert sorted(permutation.tolist()) == list(range(5000))
    queries = _element_queries(el, [0, 10, 123, 99, 499])
    _check_results_map_through_permutation(index, reordered, permutation, queries)
    # layer-preserving: the nodes of layer l stay inside [0, layer_len(l))
    for l in range(index.num_layers()):
        n = index.layer_len(l)
        assert reordered.layer_len(l) == n and sorted(permutation[:n].tolist()) == list(range(n))
    # the first layer keeps its order (reorder.rs:127), elements moved with their nodes
    assert permutation[:index.layer_len(0)].tolist() == list(range(index.layer_len(0)))
    for i in (0, 7, 4999):
        assert np.array_equal(reordered.elements.get(i), el.get(int(permutation[i])))
    # the reordered graph is the old graph renamed (reorder.rs:227-276)
    rev = np.argsort(permutation)
    for l in range(index.num_layers()):
        for i in range(0, index.layer_len(l), 97):
            assert reordered.get_neighbors(i, l) == sorted(rev[index.get_neighbors(int(permutation[i]), l)].tolist())


def test_trail_searches_every_layer_from_node_zero(oracle):
    # find_entrypoint_trail (reorder.rs:180-207) seeds layer i with eps[i], still 0: trail[i] is the max_search = 1
    # result of layer i searched from node 0
    el, index, index_bytes, eb, _ = build_fixture(oracle, "angular", 3000, 8, seed=2, num_neighbors=10, max_search=20,
                                                  layer_multiplier=6.0)
    assert index.num_layers() >= 4
    for idx in (index.layer_len(1), index.layer_len(2) + 5, 2999):
        first = next(l for l in range(index.num_layers()) if idx < index.layer_len(l))
        trail = index.entrypoint_trail(idx, first)
        assert not trail[first:].any()
        for l in range(min(first, 8)):
            head = oracle.Granne.from_bytes(_truncate_to_single_layer(oracle, index, l), el)
            want = head.search_batch(el.get(idx)[None, :], 1, 1, already_element=True)[0][0, 0]
            assert trail[l] == want


def _truncate_to_single_layer(oracle, index, layer):
    """An index image whose only layer is `layer` of `index` (so Granne::search runs search_for_neighbors from 0)."""
    return index_from_lists(oracle, [[index.get_neighbors(i, layer) for i in range(index.layer_len(layer))]])


def test_reverse_mapping_and_permute(oracle):
    # test_reverse_mapping (reorder.rs:326-334), permute_fixed_width_{identity,reverse,rand_shuffle}
    # (slice_vector/mod.rs:1028-1092)
    raw = np.arange(522 * 7, dtype=np.float32).reshape(522, 7)
    rng = np.random.default_rng(5)
    for perm in (np.arange(522), np.arange(522)[::-1].copy(), rng.permutation(522)):
        el = oracle.Elements.angular(raw, as_is=True)
        el.permute(perm)
        assert np.array_equal(el.rows(), raw[perm])
    q = random_vectors(300, 12, seed=3)
    eli = oracle.Elements.angular_int(q)
    before = eli.rows()
    perm = rng.permutation(300)
    eli.permute(perm)
    assert np.array_equal(eli.rows(), before[perm])


def test_reorder_sum_embeddings_reverse(oracle):
    # reorder_sum_embeddings_reverse (embeddings/reorder.rs:64-77)
    el = random_sum_embeddings(oracle, 25, 225, 200, seed=4)
    terms = [el.terms(i) for i in range(200)]
    vecs = [el.get(i) for i in range(200)]
    el.permute(np.arange(200)[::-1].copy())
    for i in range(200):
        assert el.terms(200 - i - 1) == terms[i]
        assert np.array_equal(el.get(200 - i - 1), vecs[i])


def test_reorder_sum_embeddings_by_keys(oracle):
    # reorder_sum_embeddings (embeddings/reorder.rs:79-102): 500 elements over 277 embeddings of dim 5
    el, index, index_bytes, eb, mb = build_fixture(oracle, "embeddings", 500, 5, seed=6, num_neighbors=30, max_search=5,
                                                   layer_multiplier=5.0, num_embeddings=277)
    keys = el.reorder_keys()
    # keys: embedding ids by decreasing norm (embeddings/reorder.rs:31-58)
    emb = el.rows()
    norms = np.sqrt(np.array([np.float32(sum(np.float32(x) * np.float32(x) for x in r)) for r in emb],
                             dtype=np.float32))
    for i in (0, 17, 499):
        t = el.terms(i)
        ranked = sorted(range(len(t)), key=lambda j: norms[t[j]])[::-1]
        want = [t[j] for j in ranked][:8]
        assert keys[i, :len(want)].tolist() == want and not keys[i, len(want):].any()
    reordered = oracle.Granne.from_bytes(index_bytes, oracle.Elements.from_bytes("embeddings", eb, mb))
    permutation = reordered.reorder_by_keys(keys)
    assert sorted(permutation.tolist()) == list(range(500))
    for l in range(index.num_layers()):
        b, e = (index.layer_len(l - 1) if l else 0), index.layer_len(l)
        seg = permutation[b:e]
        assert sorted(seg.tolist()) == list(range(b, e))
        ks = [tuple(keys[i].tolist()) + (int(i),) for i in seg]
        assert ks == sorted(ks)
    queries = _element_queries(el, [0, 10, 123, 99, 499])
    _check_results_map_through_permutation(index, reordered, permutation, queries)


@pytest.mark.parametrize("threads", [1, 4])
def test_compute_order_is_deterministic(oracle, threads):
    el, index, index_bytes, eb, _ = build_fixture(oracle, "angular_int", 2500, 16, seed=8, num_neighbors=12,
                                                  max_search=20, layer_multiplier=8.0)
    a = index.compute_order(threads=1)
    b = index.compute_order(threads=threads)
    assert np.array_equal(a, b)
