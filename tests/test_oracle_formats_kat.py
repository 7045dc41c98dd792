"""More of the reference's own tests restated against the oracle (CPU only): Offsets (src/slice_vector/offsets.rs:
296-351), MultiSetVector writers (src/slice_vector/set_vector.rs:318-425), the SumEmbeddings element container format
(src/slice_vector/mod.rs variable_width_*), and the index tests that touch the search path's files
(src/index/tests.rs: select_neighbors, empty_build, write_and_load, write_and_load_compressed, incremental builds)."""
import numpy as np
import pytest

from helpers.data import random_sum_embeddings, random_vectors

DIST_EPSILON = 10.0 * float(np.finfo(np.float32).eps)  # src/index/tests.rs:9


# ---- offsets.rs:302-351 ----
def test_offsets_general_behavior(oracle):
    reference = [9]
    for i in range(255):
        reference.append(reference[-1] + i)
    n, values, _ = oracle.offsets_roundtrip(reference)
    assert n == len(reference) and values == reference


def test_offsets_last(oracle):
    rng = np.random.default_rng(1)
    offsets = np.cumsum(rng.integers(0, 65535, size=1000)).tolist()   # test_helper::random_offsets(u16::MAX)
    n, values, last = oracle.offsets_roundtrip(offsets)
    assert last == offsets and values == offsets and n == 1000


def test_offsets_empty(oracle):
    assert oracle.offsets_roundtrip([]) == (0, [], [])
    n, values, _ = oracle.offsets_roundtrip([0] * 101)
    assert n == 101 and values == [0] * 101


def test_offsets_not_offset_panics(oracle):
    assert oracle.offsets_roundtrip([14, 5]) is None                  # #[should_panic] not_offset
    assert oracle.offsets_roundtrip([0, 70000]) is None               # delta above u16::MAX (offsets.rs:203)


# ---- set_vector.rs:318-425 ----
def test_push_unsorted(oracle):
    lists = [[(7 * j) % 15 for j in range(i, 20)] for i in range(20)]
    vec = oracle.MultiSetVector(lists)
    assert len(vec) == 20
    for i, l in enumerate(lists):
        assert vec.get(i) == sorted(l)


def test_write_fixed_width_vector_as_multi_set_vector(oracle):
    rows = [list(range(2 * i + 3, 2 * i + 3 + 7)) for i in range(120)]
    vec = oracle.MultiSetVector(rows)
    assert len(vec) == 120 and all(vec.get(i) == rows[i] for i in range(120))


def test_write_fixed_width_vector_as_multi_set_vector_predicate(oracle):
    rows = [list(range(2 * i + 3, 2 * i + 3 + 7)) for i in range(522)]
    kept = [[x for x in r if x % 3 == 0] for r in rows]               # the writer's predicate (set_vector.rs:196-199)
    vec = oracle.MultiSetVector(kept)
    assert len(vec) == 522 and all(vec.get(i) == kept[i] for i in range(522))


def test_write_fixed_width_vector_as_multi_set_vector_empty(oracle):
    assert len(oracle.MultiSetVector([])) == 0
    vec = oracle.MultiSetVector([[] for _ in range(1000)])            # ..._empty_slices: predicate drops everything
    assert len(vec) == 1000 and all(vec.get(i) == [] for i in range(0, 1000, 37))


# ---- slice_vector/mod.rs variable_width_push / write_and_load: the SumEmbeddings elements file ----
def test_variable_width_write_and_load(oracle):
    el = random_sum_embeddings(oracle, 6, 50, 300, seed=1)
    image = el.to_bytes(0)
    n = int.from_bytes(image[:8], "little")
    assert n == 300
    offsets = [int.from_bytes(image[8 + 5 * i:13 + 5 * i], "little") for i in range(n + 1)]  # FiveByteInt offsets
    assert offsets[0] == 0 and offsets[-1] == sum(2 + i % 8 for i in range(300))
    assert len(image) == 8 + 5 * (n + 1) + 3 * offsets[-1]                                    # ThreeByteInt ids
    back = oracle.Elements.from_bytes("embeddings", image, el.to_bytes(1))
    assert len(back) == 300
    for i in range(300):
        assert back.terms(i) == el.terms(i) == [j % 50 for j in range(i, i + 2 + i % 8)]
    empty = oracle.Elements.sum_embeddings(random_vectors(5, 6, seed=2), [[], [1], []])      # *_empty_left/right
    again = oracle.Elements.from_bytes("embeddings", empty.to_bytes(0), empty.to_bytes(1))
    assert [again.terms(i) for i in range(3)] == [[], [1], []]
    assert not again.get(0).any()                                                               # zeros (mod.rs:126-131)


# ---- src/index/tests.rs ----
def test_select_neighbors(oracle):
    # tests.rs:11-40
    vecs = random_vectors(51, 50, seed=3)
    el = oracle.Elements.angular(vecs[1:])
    element = oracle.normalize_f32(vecs[0])
    cands = sorted(((i, float(el.dist_to_element(i, element, already_element=True))) for i in range(50)),
                   key=lambda c: c[1])
    nb = oracle.select_neighbors(el, cands, 10)
    assert 0 < len(nb) <= 10 and all(nb[i - 1][1] <= nb[i][1] for i in range(1, len(nb)))
    assert nb[0] == cands[0]                                          # the closest candidate is always selected
    everything = oracle.select_neighbors(el, cands, 60)
    assert everything == cands


def test_empty_build(oracle):
    # tests.rs:292-301: build_partial(0) on a fresh builder leaves an index without layers
    el = oracle.Elements.angular(random_vectors(100, 25, seed=4))
    b = oracle.GranneBuilder(el)
    image = b.build_partial(0).to_bytes()
    g = oracle.Granne.from_bytes(image, el)
    assert len(g) == 0 and g.num_layers() == 0
    assert g.search(random_vectors(1, 25, seed=5)[0], 10, 5) == []    # search_internal returns Vec::new() (:978-980)


@pytest.mark.parametrize("max_search", [5, 10])
def test_write_and_load(oracle, max_search):
    # write_and_load / write_and_load_compressed (tests.rs:336-417): 100 x 50, num_neighbors 20
    el = oracle.Elements.angular(random_vectors(100, 50, seed=6))
    built = oracle.GranneBuilder(el, num_neighbors=20, max_search=max_search).build()
    index = oracle.Granne.from_bytes(built.to_bytes(), oracle.Elements.from_bytes("angular", el.to_bytes()))
    assert built.num_layers() == index.num_layers() and len(built) == len(index) == 100
    for layer in range(built.num_layers()):
        assert built.layer_len(layer) == index.layer_len(layer)
        for i in range(built.layer_len(layer)):
            assert sorted(built.get_neighbors(i, layer)) == sorted(index.get_neighbors(i, layer))
    for i in range(100):
        assert float(el.dist_to_element(i, index.elements.get(i), already_element=True)) < DIST_EPSILON


def test_incremental_build_with_write_and_read(oracle):
    # incremental_build_0/1 + incremental_build_with_write_and_read (tests.rs:134-242): build_partial in steps on ONE
    # builder; every intermediate index is a valid file whose elements find themselves
    raw = random_vectors(1000, 25, seed=7)
    el = oracle.Elements.angular(raw)
    b = oracle.GranneBuilder(el, num_neighbors=20, max_search=20)
    rows = el.rows()
    previous_layers = 0
    for num in (100, 101, 400, 1000):
        g = oracle.Granne.from_bytes(b.build_partial(num).to_bytes(), el)
        assert len(g) == num and g.num_layers() >= previous_layers
        previous_layers = g.num_layers()
        ids = g.search_batch(rows[:num], 10, 1, already_element=True)[0][:, 0]
        assert (ids == np.arange(num)).mean() > 0.95
    assert [g.layer_len(l) for l in range(g.num_layers())] == [oracle.num_elements_in_layer(1000, 15.0, l)
                                                                for l in range(g.num_layers())]
    with pytest.raises(ValueError):
        b.build_partial(10)                                           # tests.rs / mod.rs:379-382
    # a one-shot build of the same elements has the same shape
    one = oracle.GranneBuilder(el, num_neighbors=20, max_search=20).build()
    assert [one.layer_len(l) for l in range(one.num_layers())] == [g.layer_len(l) for l in range(g.num_layers())]


def _findable(g, element, idx, max_search):
    return any(i == idx for i, _ in g.search(element, max_search, 1))


def test_append_elements(oracle):
    # append_elements (tests.rs:503-567): push 500, build, push 500 more, build; expected_num_elements(1000),
    # layer_multiplier(10), num_neighbors(20), max_search(50) -> three layers both times
    first, second = random_vectors(500, 50, seed=8), random_vectors(500, 50, seed=9)
    el = oracle.Elements("angular", 50)
    b = oracle.GranneBuilder(el, num_neighbors=20, max_search=50, layer_multiplier=10.0, expected_num_elements=1000)
    el.push(first)
    g = b.build_partial(len(el))
    assert g.num_layers() == 3 and g.layer_len(2) == 500
    assert _findable(g, first[123], 123, 50)
    el.push(second)
    g = b.build_partial(len(el))
    assert g.num_layers() == 3 and g.layer_len(2) == 1000
    assert _findable(g, first[123], 123, 50) and _findable(g, second[123], 500 + 123, 50)


def test_with_elements_and_add_and_borrowed(oracle):
    # with_borrowed_elements / with_elements_and_add (tests.rs:64-112): verify_search(index, 0.95, 40)
    raw = random_vectors(600, 25, seed=10)
    el = oracle.Elements.angular(raw[:500])
    g = oracle.GranneBuilder(el, max_search=5, reinsert_elements=False).build()
    rows = el.rows()
    assert len(g) == 500
    assert (g.search_batch(rows, 40, 1, already_element=True)[0][:, 0] == np.arange(500)).mean() > 0.95
    b = oracle.GranneBuilder(el, num_neighbors=20, max_search=5, expected_num_elements=600)
    el.push(raw[500:])
    assert len(el) == 600
    g = b.build_partial(600)
    rows = el.rows()
    assert (g.search_batch(rows, 40, 1, already_element=True)[0][:, 0] == np.arange(600)).mean() > 0.95


def test_read_index_reduce_num_neighbors(oracle):
    # read_index_reduce_num_neighbors (tests.rs:243-290): build half with num_neighbors 20, reload the file into a
    # builder with num_neighbors 5, finish the build
    el = oracle.Elements.angular(random_vectors(1000, 5, seed=11))
    half = oracle.GranneBuilder(el, num_neighbors=20, max_search=10).build_partial(500)
    assert len(half.get_neighbors(0)) > 5                             # "not necessarily true, but should be valid"
    loaded = oracle.Granne.from_bytes(half.to_bytes(), el)
    b = oracle.GranneBuilder.from_index(loaded, el, num_neighbors=5, max_search=10)
    g = b.build_partial(500)
    assert len(g) == 500 and g.num_layers() == half.num_layers()
    g = b.build_partial(1000)
    assert len(g) == 1000 and len(g.get_neighbors(0)) <= 5
    for layer in range(g.num_layers()):
        for i in range(0, g.layer_len(layer), 41):
            assert len(g.get_neighbors(i, layer)) <= 5
