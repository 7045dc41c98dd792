"""Pins the CPU oracle against every known-answer test the reference holds for the search path.

Each test names the reference test it re-encodes (paths relative to /root/reference).  The reference's data is
unseeded (src/test_helper.rs:3-18), so statistical tests are re-run here with fixed seeds.
"""
import numpy as np
import pytest

EPS = float(np.finfo(np.float32).eps)
DIST_EPSILON = 10.0 * EPS  # src/elements/angular.rs:97


def random_floats(rng, n):  # src/test_helper.rs:3-6: U(0,1) - 0.5
    return (rng.random(n, dtype=np.float32) - np.float32(0.5)).astype(np.float32)


# ---- src/index/tests.rs:305-335 test_num_elements_in_layer (exact) ----
@pytest.mark.parametrize("total,mult,expected", [
    (1000, 10.0, [10, 100, 1000]),
    (32, 2.0, [1, 2, 4, 8, 16, 32]),
    (10_000, 10.0, [1, 10, 100, 1000, 10_000, 10_000]),
    (20, 1.9, [2, 3, 6, 11, 20, 20]),
    (1_000_000_000, 20.0, [16, 313, 6250, 125_000, 2_500_000, 50_000_000, 1_000_000_000, 1_000_000_000]),
    (50, 100.0, [50]),
    (133689866, 15.0, [12, 177, 2641, 39612, 594178, 8912658, 133689866]),
])
def test_num_elements_in_layer(oracle, total, mult, expected):
    assert [oracle.num_elements_in_layer(total, mult, l) for l in range(len(expected))] == expected


# ---- src/slice_vector/set_vector.rs:231-248 ----
def test_delta_encode(oracle):
    assert oracle.delta_encode([1, 2, 2, 4]) == [1, 1, 0, 2]


def test_delta_encode_decode(oracle):
    data = [123, 345, 555, 555, 6999, 7000]
    assert oracle.set_decode(oracle.set_encode(data)) == data


# ---- src/slice_vector/set_vector.rs:250-312 push_and_get* ----
def test_push_and_get(oracle):
    s = list(range(10))
    assert oracle.set_decode(oracle.set_encode(s)) == s


def test_push_and_get_empty(oracle):
    enc = oracle.set_encode([])
    assert enc == b"\x00"
    assert oracle.set_decode(enc) == []


def test_push_and_get_4_bytes_per_number(oracle):
    # set_vector.rs:275-283: "the size of the encoded data has the same size as the original (2 * 4 bytes)"
    # -> vbyte (1 control + 2 + 3 + 1 + 1 = 8 bytes) is not smaller than raw (8 bytes): stored raw, little endian.
    exp = [37717, 660380]
    enc = oracle.set_encode(exp)
    assert len(enc) == 1 + 8
    assert enc[0] == 2
    assert int.from_bytes(enc[1:5], "little") == 37717
    assert int.from_bytes(enc[5:9], "little") == 660380 - 37717
    assert oracle.set_decode(enc) == exp


def test_push_and_get_one_and_duplicates(oracle):
    assert oracle.set_decode(oracle.set_encode([5])) == [5]
    assert oracle.set_decode(oracle.set_encode([5, 5])) == [5, 5]


def test_push_multiple(oracle):
    for i in range(20):
        s = list(range(i, 20))
        assert oracle.set_decode(oracle.set_encode(s)) == s


def test_stream_vbyte_layout(oracle):
    # published Stream VByte layout: control bytes first, 2-bit (len-1) codes, first number in the low bits.
    enc = oracle.set_encode([1, 1 + 300, 1 + 300 + 70000, 1 + 300 + 70000 + 20000000, 1 + 300 + 70000 + 20000000])
    assert enc[0] == 5
    # deltas: 1 (1B), 300 (2B), 70000 (3B), 20000000 (4B) | 0 (1B)
    assert enc[1] == (0 | (1 << 2) | (2 << 4) | (3 << 6))
    assert enc[2] == 0
    assert enc[3:4] == bytes([1])
    assert enc[4:6] == (300).to_bytes(2, "little")
    assert enc[6:9] == (70000).to_bytes(3, "little")
    assert enc[9:13] == (20000000).to_bytes(4, "little")
    assert enc[13:14] == b"\x00"
    assert len(enc) == 14


# ---- src/odd_byte_int.rs:43-79 ----
def test_odd_byte_ints(oracle):
    assert oracle.read_uint(oracle.write_uint(123456, 3), 3) == 123456
    five_max = (1 << 40) - 1
    for v in [0, 1, 2, 3, 1234567, five_max - 1, five_max, 7_301_010_345]:
        b = oracle.write_uint(v, 5)
        assert len(b) == 5
        assert oracle.read_uint(b, 5) == v
        assert int.from_bytes(b, "little") == v


# ---- src/math.rs:166-196 ----
def test_math_sum(oracle):
    rng = np.random.default_rng(1)
    for n in range(1, 101):
        x, y = random_floats(rng, n), random_floats(rng, n)
        assert np.array_equal(oracle.sum_into_f32(x, y), x + y)


def test_math_dot_product(oracle):
    rng = np.random.default_rng(2)
    for n in range(1, 101):
        x, y = random_floats(rng, n), random_floats(rng, n)
        expected = np.float32(0)
        for i in range(n):
            expected = np.float32(expected + np.float32(x[i] * y[i]))
        assert abs(float(expected) - oracle.dot_product_f32(x, y)) < 0.000001


def test_dot_product_order_is_the_32_lane_fma_order(oracle):
    # src/math.rs:16-42 restated independently in numpy (fma emulated in f64: the product of two f32 is exact in
    # f64; product + f32 addend fits well inside f64 for these magnitudes, so a single rounding to f32 follows).
    rng = np.random.default_rng(3)
    for n in [1, 5, 31, 32, 33, 64, 100, 128, 131]:
        x, y = random_floats(rng, n), random_floats(rng, n)
        chunk = np.zeros(32, dtype=np.float32)
        full = n // 32
        for c in range(full):
            for i in range(32):
                chunk[i] = np.float32(np.float64(x[c * 32 + i]) * np.float64(y[c * 32 + i]) + np.float64(chunk[i]))
        r = np.float32(0)
        for i in range(32):
            r = np.float32(r + chunk[i])
        for i in range(full * 32, n):
            r = np.float32(np.float64(x[i]) * np.float64(y[i]) + np.float64(r))
        assert oracle.dot_product_f32(x, y) == float(r), n


# ---- src/elements/angular.rs:97-126 ----
def _reference_dist(x, y):  # angular.rs:78-90
    r = np.float32(0)
    dx = np.float32(0)
    dy = np.float32(0)
    for a, b in zip(x, y):
        r = np.float32(r + np.float32(a * b))
        dx = np.float32(dx + np.float32(a * a))
        dy = np.float32(dy + np.float32(b * b))
    d = np.float32(1) - np.float32(r / np.float32(np.sqrt(dx) * np.sqrt(dy)))
    return max(np.float32(0), d)


def test_angular_reference_dist(oracle):
    rng = np.random.default_rng(4)
    for _ in range(100):
        x = oracle.normalize_f32(random_floats(rng, 100))
        y = oracle.normalize_f32(random_floats(rng, 100))
        assert abs(float(oracle.dist_f32(x, y)) - float(_reference_dist(x, y))) < DIST_EPSILON


def test_angular_dist_same_and_opposite(oracle):
    rng = np.random.default_rng(5)
    for _ in range(100):
        x = oracle.normalize_f32(random_floats(rng, 100))
        assert float(oracle.dist_f32(x, x)) < DIST_EPSILON
        y = oracle.normalize_f32(-x)
        assert float(oracle.dist_f32(x, y)) > 2.0 - DIST_EPSILON


def test_angular_small_and_large_arrays(oracle):
    a = oracle.normalize_f32(np.array([0, 1, 2], dtype=np.float32))
    assert float(oracle.dist_f32(a, a)) >= 0.0
    b = oracle.normalize_f32(np.ones(100, dtype=np.float32))
    assert float(oracle.dist_f32(b, b)) < DIST_EPSILON


# ---- src/elements/angular_int.rs:28-59 ----
def test_angular_int_quantize_and_dist(oracle):
    q = oracle.quantize_i8(np.array([0.5, -0.25, 0.1, -0.5], dtype=np.float32))
    assert q.tolist() == [127, -63, 25, -127]  # x*127/max|x| truncated toward zero
    s = 127 * 127 + 63 * 63 + 25 * 25 + 127 * 127
    assert oracle.dot_i8(q, q) == (s, s, s)
    assert float(oracle.dist_i8(q, q)) < DIST_EPSILON
    z = np.zeros(4, dtype=np.int8)
    assert float(oracle.dist_i8(z, q)) == 1.0  # 0/0 = NaN -> r = 0 -> d = 1 (angular_int.rs:55)
    rng = np.random.default_rng(6)
    for _ in range(50):
        x, y = random_floats(rng, 100), random_floats(rng, 100)
        qx, qy = oracle.quantize_i8(x), oracle.quantize_i8(y)
        ref = _reference_dist(qx.astype(np.float32), qy.astype(np.float32))
        assert abs(float(oracle.dist_i8(qx, qy)) - float(ref)) < 1e-5
