"""Integer parity of the counter-based RNG (include/rt3_rng.h) with its definition,
reference src/lib/shaders/random_v1.glsl:22-53, restated here in Python integers."""
import ctypes as C

import numpy as np

import oraclelib as ol
from rt3_b200 import scenes

M = 0xFFFFFFFF


def py_hash1(x):
    x = (x + (x << 10)) & M
    x ^= x >> 6
    x = (x + (x << 3)) & M
    x ^= x >> 11
    x = (x + (x << 15)) & M
    return x


def py_hash4(x, y, z, w):
    return py_hash1(x ^ py_hash1(y) ^ py_hash1(z) ^ py_hash1(w))


def test_hash_known_answers(built):
    lib = ol.port()
    rng = np.random.default_rng(1)
    values = [0, 1, 2, 0x7FFFFFFF, 0x80000000, M] + rng.integers(0, 2**32, 200).tolist()
    for v in values:
        assert lib.orc_hash1(v) == py_hash1(v)
    assert py_hash1(0) == 0 and lib.orc_hash1(0) == 0
    for a, b, c, d in rng.integers(0, 2**32, (100, 4)).tolist():
        assert lib.orc_hash4(a, b, c, d) == py_hash4(a, b, c, d)
    assert np.array_equal(scenes.hash1(np.array(values, np.uint64)), np.array([py_hash1(v) for v in values], np.uint32))


def test_float_construct(built):
    lib = ol.port()
    assert lib.orc_float_construct(0) == 0.0
    assert lib.orc_float_construct(0xFF800000) == 0.0  # only the 23 mantissa bits are used
    top = lib.orc_float_construct(M)
    assert top == np.float32(1.0) - np.float32(2.0 ** -23)
    for m in (1, 0x400000, 0x7FFFFF, 0x12345678):
        assert lib.orc_float_construct(m) == np.float32((m & 0x7FFFFF) / 2.0 ** 23)


def test_draw_is_hash4_with_tagged_dimension(built):
    lib = ol.port()
    for pixel, sample, seed, dim in [(0, 0, 0, 0), (89999, 99, 7, 5), (959999, 499, 0x5EED, 4 + 8 * 49 + 4)]:
        expect = lib.orc_float_construct(py_hash4(pixel, sample, 0x80000000 | dim, seed))
        assert lib.orc_draw(pixel, sample, seed, dim) == expect
    # the dimension tag removes the (sample, dim) <-> (dim, sample) symmetry of the xor-combined hash
    assert lib.orc_draw(5, 3, 1, 9) != lib.orc_draw(5, 9, 1, 3)


def test_draws_are_uniform(built):
    lib = ol.port()
    xs = np.array([lib.orc_draw(p, s, 11, 4) for p in range(200) for s in range(50)], np.float64)
    assert 0.0 <= xs.min() and xs.max() < 1.0
    assert abs(xs.mean() - 0.5) < 0.01 and abs(xs.var() - 1 / 12) < 0.005
    hist, _ = np.histogram(xs, bins=16, range=(0, 1))
    assert hist.min() > 0.85 * len(xs) / 16


def test_sincos_polynomial(built):
    lib = ol.port()
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for x in np.linspace(0, 1, 4001, endpoint=False, dtype=np.float32).tolist() + [0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 0.99999994]:
        lib.orc_sincos_2pi(x, C.byref(s), C.byref(c))
        worst = max(worst, abs(s.value - np.sin(2 * np.pi * x)), abs(c.value - np.cos(2 * np.pi * x)))
    assert worst < 4e-7
