"""
numpy restatement of the part of `jax.random` the reference uses to initialise parameters
(fbpinns/trainers.py:583, 603-605; fbpinns/networks.py:44-58): Threefry-2x32 keys, `PRNGKey`, `split`, `uniform`.

JAX is not installed here, so this is a restatement of its published algorithm (jax/_src/prng.py, the default
`threefry2x32` implementation with `jax_threefry_partitionable=False`, the default of every JAX release before 0.5.0 —
the reference pins `jax>=0.4.8`).  Pinned by the known-answer vectors of the Threefry reference implementation
(Random123) that JAX's own test-suite uses, and by two widely published JAX outputs
(`split(PRNGKey(0))`, `uniform(PRNGKey(0))`) — see tests/test_host_logic.py::test_jax_prng_known_answers.  Anything
beyond those vectors (e.g. a JAX release with the partitionable default) is unpinned; DESIGN.md §7 says so.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    "Threefry-2x32, 20 rounds, on uint32 arrays x0, x1 (elementwise) with the key (k0, k1)."
    with np.errstate(over="ignore"):
        k0, k1 = _U32(k0), _U32(k1)
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 = (np.asarray(x0, dtype=_U32) + ks[0]).astype(_U32)
        x1 = (np.asarray(x1, dtype=_U32) + ks[1]).astype(_U32)
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r).astype(_U32)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(i + 1) % 3]).astype(_U32)
            x1 = (x1 + ks[(i + 2) % 3] + _U32(i + 1)).astype(_U32)
    return x0, x1


def threefry_2x32(key, count):
    "jax._src.prng.threefry_2x32: hash a uint32 array of any shape; odd sizes are padded with one zero"
    count = np.asarray(count, dtype=_U32)
    flat = count.ravel()
    odd = flat.size % 2
    if odd:
        flat = np.concatenate([flat, np.zeros(1, dtype=_U32)])
    half = flat.size // 2
    y0, y1 = threefry2x32(key[0], key[1], flat[:half], flat[half:])
    out = np.concatenate([y0, y1])
    return (out[:-1] if odd else out).reshape(count.shape)


def PRNGKey(seed):
    "key of an integer seed: (high word, low word) — (0, seed) for the 32-bit seeds of a default (x64 off) JAX"
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def split(key, num=2):
    "jax.random.split -> (num, 2) uint32"
    return threefry_2x32(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits(key, shape):
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    return threefry_2x32(key, np.arange(size, dtype=_U32)).reshape(shape)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    "jax.random.uniform for float32: 23 random mantissa bits -> [1, 2) - 1, then the affine map in float32"
    bits = random_bits(key, tuple(shape))
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    minval, maxval = np.float32(minval), np.float32(maxval)
    return np.maximum(minval, (f * (maxval - minval) + minval).astype(np.float32))


# ---- the same functions over a batch of keys (what `vmap(network.init_params)(subkeys, ...)` computes row by row) ------
def _hash_batched(keys, count):
    "threefry_2x32 of the same `count` vector under every key of `keys` (B, 2) -> (B, count.size)"
    keys = np.asarray(keys, dtype=_U32)
    flat = np.asarray(count, dtype=_U32).ravel()
    odd = flat.size % 2
    if odd:
        flat = np.concatenate([flat, np.zeros(1, dtype=_U32)])
    half = flat.size // 2
    y0, y1 = threefry2x32(keys[:, 0:1], keys[:, 1:2], flat[None, :half], flat[None, half:])
    out = np.concatenate([y0, y1], axis=1)
    return out[:, :-1] if odd else out


def split_batched(keys, num=2):
    "(B, 2) keys -> (B, num, 2)"
    return _hash_batched(keys, np.arange(2 * num, dtype=_U32)).reshape(len(keys), num, 2)


def uniform_batched(keys, shape, minval, maxval):
    "(B, 2) keys -> (B, *shape) float32, row b == uniform(keys[b], shape, minval, maxval)"
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    bits = _hash_batched(keys, np.arange(size, dtype=_U32))
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    minval, maxval = np.float32(minval), np.float32(maxval)
    return np.maximum(minval, (f * (maxval - minval) + minval).astype(np.float32)).reshape((len(keys),) + tuple(shape))


def normal(key, shape=()):
    """jax.random.normal for float32: sqrt(2) * erfinv(u), u uniform on (nextafter(-1, 0), 1).  XLA's float32 erf_inv
    polynomial is replaced by scipy's float64 erfinv rounded to float32 (agrees to float32 rounding, not bit for bit)."""
    from scipy.special import erfinv
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, lo, np.float32(1.0))
    return (np.sqrt(2.0) * erfinv(u.astype(np.float64))).astype(np.float32)


def is_key(key):
    return isinstance(key, np.ndarray) and key.dtype == np.uint32 and key.shape == (2,)
