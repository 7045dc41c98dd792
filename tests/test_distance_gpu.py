"""compute_distance (py/src/lib.rs:71-89) through the C ABI: Vector::from(a).dist(&Vector::from(b)) for raw f32
vectors, bit-identical to the oracle's angular (angular.rs:55-74) and angular_int (angular_int.rs:19-59) distances."""
import numpy as np
import pytest

import granne_b200
from helpers.data import random_vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _lib():
    from granne_b200 import build

    build.build()
    granne_b200.load_library()


def _oracle_pairs(oracle, kind, a, b):
    out = np.empty(a.shape[0], dtype=np.float32)
    for i in range(a.shape[0]):
        if kind == "angular":
            out[i] = oracle.dist_f32(oracle.normalize_f32(a[i]), oracle.normalize_f32(b[i]))
        else:
            out[i] = oracle.dist_i8(oracle.quantize_i8(a[i]), oracle.quantize_i8(b[i]))
    return out


@pytest.mark.parametrize("kind", ["angular", "angular_int"])
@pytest.mark.parametrize("dim", [1, 3, 25, 32, 33, 64, 100, 128, 160, 200, 256, 300, 1000])
def test_pairwise_distances_are_bit_identical(oracle, kind, dim):
    n = 200
    a, b = random_vectors(n, dim, seed=dim), random_vectors(n, dim, seed=dim + 1)
    a[0], b[0] = 0.0, b[1]                   # zero vector: normalise leaves zeros / i8 norm 0 -> NaN -> r = 0
    a[1] = b[1]                              # identical
    a[2] = -b[2]                             # opposite
    a[3] = b[3] * 1e-3                       # same direction, different scale
    got = granne_b200.compute_distances(kind, a, b)
    want = _oracle_pairs(oracle, kind, a, b)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert abs(float(got[1])) < 1e-5 and abs(float(got[2]) - 2.0) < 1e-5     # angular.rs:112-126
    assert granne_b200.compute_distance(kind, a[5], b[5]) == float(want[5])


def test_nan_is_an_error_like_the_reference_panic():
    a, b = random_vectors(4, 16, seed=1), random_vectors(4, 16, seed=2)
    a[2, 3] = np.nan
    with pytest.raises(granne_b200.GranneError) as ei:
        granne_b200.compute_distances("angular", a, b)
    assert ei.value.code == -6                                               # NotNan panic, angular.rs:70
    with pytest.raises(ValueError):
        granne_b200.compute_distance("embeddings", a[0], b[0])               # "Unsupported element type"
