"""JAX threefry PRNG restated in NumPy (oracle; test infrastructure only).

The reference draws its per-iteration update order and its truncated-normal Monte-Carlo
samples from jax.random (caviar.py:76 PRNGKey, :196 choice, :209 split, :210 uniform,
:304 split).  JAX is a third-party dependency that is absent from /root/reference and
from this image (README.md:22 pins "JAX versions up to 0.3.15"), so its published
algorithm (jax/_src/prng.py, jax/_src/random.py of that era) is restated here:

  threefry2x32      Random123 Threefry-2x32, 20 rounds (KATs in tests/test_oracle_prng.py)
  threefry_2x32     JAX's array wrapper: flat uint32 counts split in two halves
  split             threefry over iota(2*num) reshaped (num, 2)
  random_bits       32- and 64-bit draws (64-bit = first half << 32 | second half)
  uniform_f64       mantissa-fill construction in [0, 1)
  permutation       sort-based shuffle, ceil(3 ln N / ln(2^32-1)) rounds

PARITY UNPINNED against a live JAX for the 64-bit uniform packing and the shuffle round
structure (restated from the published source; no JAX available to confirm).
"""
import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 block function on uint32 arrays (broadcasting)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def prng_key(seed):
    """jax.random.PRNGKey(seed): uint32 pair (seed >> 32, seed & 0xffffffff)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def threefry_2x32(key, counts):
    """JAX wrapper: counts (flat uint32) -> same-length uint32 random words."""
    counts = np.asarray(counts, dtype=U32).ravel()
    n = counts.size
    odd = n % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, U32)])
    h = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:h], counts[h:])
    out = np.concatenate([y0, y1])
    return out[:-1] if odd else out


def split(key, num=2):
    return threefry_2x32(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def random_bits(key, bit_width, size):
    """Flat array of `size` random unsigned ints of width 32 or 64."""
    if bit_width == 32:
        return threefry_2x32(key, np.arange(size, dtype=U32))
    assert bit_width == 64
    bits = threefry_2x32(key, np.arange(2 * size, dtype=U32))
    hi, lo = bits[:size].astype(np.uint64), bits[size:].astype(np.uint64)
    return (hi << np.uint64(32)) | lo


def uniform_f64(key, shape):
    """jax.random.uniform(key, shape) with x64 enabled -> float64 in [0, 1)."""
    size = int(np.prod(shape))
    bits = random_bits(key, 64, size)
    fb = (bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)
    return (fb.view(np.float64) - 1.0).reshape(shape)


def uniform_f32(key, shape):
    size = int(np.prod(shape)) if shape != () else 1
    bits = random_bits(key, 32, size)
    fb = (bits >> U32(9)) | U32(0x3F800000)
    out = fb.view(np.float32) - np.float32(1.0)
    return out.reshape(shape)


def shuffle_rounds(n):
    return int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))


def permutation(key, n):
    """jax.random.permutation(key, n) == choice(key, n, [n], replace=False) (caviar.py:196)."""
    x = np.arange(n)
    for _ in range(shuffle_rounds(n)):
        key, sub = split(key)
        sort_keys = random_bits(sub, 32, n)
        x = x[np.argsort(sort_keys, kind="stable")]
    return x
