"""TEST INFRASTRUCTURE (oracle): NumPy restatement of jax's default PRNG as the reference uses it for the
turbulent driving -- jax.random.key(42) (time_integration.py:787-789), jax.random.split(key, 3) and
jax.random.normal(sk, shape=(3, nx, ny, nz)) (_turbulent_forcing.py:48-51).

jax itself is absent from /root/reference and from this image (pyproject.toml:9-16 asks for jax>=0.4.0,
unpinned), so the published algorithm is restated:
  * Threefry-2x32, 20 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
    Random123 known-answer vectors pin the block function bit for bit: tests/test_threefry.py);
  * jax/_src/prng.py: threefry_seed (key = (seed >> 32, seed & 0xffffffff)), threefry_split and
    threefry_random_bits in both layouts -- `partitionable=True` (default since jax 0.5.0: 64-bit iota
    counters, one block per sample, bits = out0 ^ out1) and `partitionable=False` (the older default: 32-bit
    iota counters, first half of the counters in word 0, second half in word 1);
  * jax/_src/random.py: uniform (mantissa bits | 1.0, minus 1, scaled to [nextafter(-1, 0), 1)) and
    _normal_real = sqrt(2) * erf_inv(u); XLA's fp32 erf_inv is Giles' single-precision polynomial
    (xla/client/lib/math.cc ErfInv32), restated here in float32; fp64 uses scipy.special.erfinv.
Parity status: the block function is pinned exactly; the layer above it is pinned by construction against the
published algorithm and by one remembered documentation value (see the test); not by a run of jax.
"""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = np.uint32(0x1BD11BDA)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32-20 of the counter words (c0, c1) under the key (k0, k1); uint32 arrays in and out."""
    k0, k1 = np.uint32(k0), np.uint32(k1)
    x0 = np.asarray(c0, dtype=np.uint32).copy()
    x1 = np.asarray(c1, dtype=np.uint32).copy()
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    with np.errstate(over="ignore"):
        x0 += ks[0]
        x1 += ks[1]
        for blk in range(5):
            for r in _ROT[blk & 1]:
                x0 += x1
                x1 = _rotl(x1, r)
                x1 ^= x0
            x0 += ks[(blk + 1) % 3]
            x1 += ks[(blk + 2) % 3] + np.uint32(blk + 1)
    return x0, x1


def seed_key(seed):
    """jax.random.key(seed) / PRNGKey(seed): (high word, low word)."""
    seed = int(seed)
    return np.uint32((seed >> 32) & 0xFFFFFFFF), np.uint32(seed & 0xFFFFFFFF)


def _threefry_2x32_flat(key, counts):
    """prng.threefry_2x32: the flat counter array is cut in two halves = the two counter words (odd sizes are
    padded with one zero); the two output words are concatenated again."""
    counts = np.asarray(counts, dtype=np.uint32).ravel()
    odd = counts.size % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, np.uint32)])
    h = counts.size // 2
    o0, o1 = threefry2x32(key[0], key[1], counts[:h], counts[h:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def split(key, num, partitionable=True):
    """jax.random.split(key, num) -> list of keys."""
    if partitionable:
        o0, o1 = threefry2x32(key[0], key[1], np.zeros(num, np.uint32), np.arange(num, dtype=np.uint32))
        return [(o0[i], o1[i]) for i in range(num)]
    out = _threefry_2x32_flat(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)
    return [(out[i, 0], out[i, 1]) for i in range(num)]


def random_bits(key, shape, bit_width=32, partitionable=True):
    size = int(np.prod(shape)) if len(shape) else 1
    if partitionable:
        idx = np.arange(size, dtype=np.uint64)
        o0, o1 = threefry2x32(key[0], key[1], (idx >> np.uint64(32)).astype(np.uint32), idx.astype(np.uint32))
        if bit_width == 32:
            return (o0 ^ o1).reshape(shape)
        return ((o0.astype(np.uint64) << np.uint64(32)) | o1.astype(np.uint64)).reshape(shape)
    nwords = size * bit_width // 32
    bits = _threefry_2x32_flat(key, np.arange(nwords, dtype=np.uint32))
    if bit_width == 32:
        return bits.reshape(shape)
    hi, lo = bits[:size].astype(np.uint64), bits[size:].astype(np.uint64)
    return ((hi << np.uint64(32)) | lo).reshape(shape)


def erfinv32(x):
    """XLA ErfInv32 (Giles 2010, single precision), evaluated in float32."""
    x = np.asarray(x, dtype=np.float32)
    f = np.float32
    w = -np.log1p(-(x * x)).astype(np.float32)
    lt = w < f(5.0)
    wa = np.where(lt, w - f(2.5), np.sqrt(np.maximum(w, f(0))) - f(3.0)).astype(np.float32)
    ca = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
          -0.00417768164, 0.246640727, 1.50140941]
    cb = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
          0.00943887047, 1.00167406, 2.83297682]
    p = np.where(lt, f(ca[0]), f(cb[0])).astype(np.float32)
    for a, b in zip(ca[1:], cb[1:]):
        p = (np.where(lt, f(a), f(b)) + p * wa).astype(np.float32)
    out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == f(1.0), np.copysign(np.float32(np.inf), x), out)


def uniform(key, shape, dtype, minval, maxval, partitionable=True):
    dtype = np.dtype(dtype)
    nbits, nmant, ut = (32, 23, np.uint32) if dtype == np.float32 else (64, 52, np.uint64)
    bits = random_bits(key, shape, nbits, partitionable)
    one = np.array(1.0, dtype).view(ut)
    fb = (bits >> ut(nbits - nmant)) | one
    floats = fb.view(dtype) - dtype.type(1.0)
    lo, hi = dtype.type(minval), dtype.type(maxval)
    return np.maximum(lo, (floats * (hi - lo) + lo).astype(dtype))


def normal(key, shape, dtype=np.float32, partitionable=True):
    """jax.random.normal(key, shape, dtype)."""
    dtype = np.dtype(dtype)
    lo = np.nextafter(dtype.type(-1.0), dtype.type(0.0), dtype=dtype)
    u = uniform(key, shape, dtype, lo, 1.0, partitionable)
    if dtype == np.float32:
        return (np.float32(np.sqrt(2)) * erfinv32(u)).astype(np.float32)
    from scipy.special import erfinv
    return np.sqrt(2.0) * erfinv(u)
