"""Pins the oracle's threefry PRNG against the Random123 known-answer tests and values published in JAX's docs."""
import numpy as np

from oracle import prng


def test_threefry2x32_random123_kats():
    kats = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
            ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
            ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, want in kats:
        got = prng.threefry2x32(key[0], key[1], ctr[0], ctr[1])
        assert (int(got[0]), int(got[1])) == want


def test_split_and_uniform_match_published_jax_values():
    k = prng.prng_key(0)
    s = prng.split(k)
    assert s.tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert abs(float(prng.uniform_f32(k, ())) - 0.41845703) < 1e-8


def test_random_bits_match_jax_own_test_vectors():
    """Known answers of JAX's own test-suite (jax/tests/random_test.py::testRngRandomBits, non-partitionable threefry --
    the scheme of the JAX <= 0.3.15 the reference pins): key = PRNGKey(1701), shape (3,).  The odd size exercises the
    padding / split-in-halves convention of random_bits; the 64-bit words are what uniform() consumes under x64."""
    key = prng.prng_key(1701)
    assert [int(x) for x in prng.random_bits(key, 32, 3)] == [56197195, 4200222568, 961309823]
    assert [int(x) for x in prng.random_bits(key, 64, 3)] == [3982329540505020460, 16822122385914693683,
                                                              7882654074788531506]


def test_uniform_f64_range_and_shape():
    u = prng.uniform_f64(prng.prng_key(5), (100, 2))
    assert u.shape == (100, 2) and u.dtype == np.float64
    assert (u >= 0).all() and (u < 1).all() and 0.3 < u.mean() < 0.7


def test_permutation_is_a_permutation_and_odd_sizes():
    for n in (1, 2, 7, 100, 101, 1700):
        p = prng.permutation(prng.prng_key(n), n)
        assert sorted(p.tolist()) == list(range(n))
    assert prng.shuffle_rounds(1625) == 1 and prng.shuffle_rounds(1626) == 2
