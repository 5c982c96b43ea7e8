"""CPU-only checks of the two small headers the product shares with its checker: the counter RNG (include/crt_rng.h)
and the shading sine/cosine (include/crt_math.h).  Both are exercised through the oracle library, which includes the
very same headers the CUDA kernels include."""
import numpy as np

from oracle import orc

PHI = np.uint32(0x9E3779B9)


def mix32(x):
    """numpy restatement of crt_mix32 (uint32 arithmetic wraps)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16); x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15); x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def rng_key(seed, pixel, sample):
    s = mix32(np.array([seed ^ 0x85EBCA6B], np.uint32))[0]
    rot = np.uint32(((int(s) << 16) | (int(s) >> 16)) & 0xFFFFFFFF)
    return mix32(pixel.astype(np.uint32) ^ s), mix32((sample.astype(np.uint32) + np.uint32(0x7F4A7C15)) ^ rot)


def rng_bits(k0, k1, dim):
    return mix32(k1 + mix32(k0 ^ (np.uint32(dim) * PHI)))


def test_numpy_restatement_equals_the_header():
    rs = np.random.RandomState(3)
    for _ in range(300):
        seed, pixel, sample, dim = (int(v) for v in rs.randint(0, 2 ** 31, 4))
        k0, k1 = rng_key(seed, np.array([pixel]), np.array([sample]))
        assert int(rng_bits(k0, k1, dim)[0]) == orc.counter_bits(seed, pixel, sample, dim)
    assert orc.counter_uniform(1, 2, 3, 4) == (orc.counter_bits(1, 2, 3, 4) >> 8) / 16777216.0


def test_keys_are_unique_per_pixel_and_sample():
    """pixel -> k0 and sample -> k1 are bijections of 32-bit words: a 8.5e9-path render (BASELINE config 5) has no two
    paths with one key.  Checked on 2^22 consecutive and 2^22 random words each."""
    rs = np.random.RandomState(5)
    for words in (np.arange(1 << 22, dtype=np.uint32), rs.randint(0, 2 ** 32, 1 << 22, dtype=np.uint64).astype(np.uint32)):
        words = np.sort(words)
        words = words[np.concatenate(([True], words[1:] != words[:-1]))]
        k0, k1 = rng_key(1234, words, words)
        for k in (k0, k1):
            k = np.sort(k)
            assert (k[1:] != k[:-1]).all()


def shifted_stream_hits(bits_fn, n_pixels, n_samples, max_shift):
    """How many (path A, shift s >= 1) pairs exist such that draws s, s+1 of A equal draws 0, 1 of some path B (the
    signature of B's stream being A's stream shifted by s), plus how many paths share draws 0, 1 outright."""
    pixel = np.repeat(np.arange(n_pixels, dtype=np.uint32), n_samples)
    sample = np.tile(np.arange(n_samples, dtype=np.uint32), n_pixels)
    draw = lambda d: bits_fn(pixel, sample, d).astype(np.uint64)
    prev = draw(0)
    cur = draw(1)
    anchor = np.sort((prev << np.uint64(32)) | cur)
    same = int((anchor[1:] == anchor[:-1]).sum())
    hits = 0
    for s in range(1, max_shift + 1):
        prev, cur = cur, draw(s + 1)
        code = np.sort((prev << np.uint64(32)) | cur)          # sorted queries: a cache-friendly merge
        pos = np.searchsorted(anchor, code)
        pos[pos == len(anchor)] = 0
        hits += int((anchor[pos] == code).sum())
    return same, hits


def test_no_two_of_ten_million_paths_have_shifted_equal_streams():
    """1e7 (pixel, sample) keys, shifts up to 12 draws (a path consumes at most ~45): none.  The round-1 generator
    (one 32-bit key, draw = mix(key + dim * phi)) is run through the same check on 1e6 keys to show the check bites."""
    new = lambda px, sm, d: rng_bits(*rng_key(1234, px, sm), d)
    assert shifted_stream_hits(new, 100_000, 100, 12) == (0, 0)

    def old(px, sm, d):
        k = mix32(np.array([1234 ^ 0x85EBCA6B], np.uint32))[0]
        k = mix32(k ^ px)
        k = mix32(k + sm * PHI + np.uint32(0x7F4A7C15))
        return mix32(k + np.uint32(d) * PHI)
    same, hits = shifted_stream_hits(old, 10_000, 100, 12)
    assert hits > 100


def test_uniforms_are_uniform_and_uncorrelated():
    pixel = np.repeat(np.arange(2000, dtype=np.uint32), 50)
    sample = np.tile(np.arange(50, dtype=np.uint32), 2000)
    k0, k1 = rng_key(1234, pixel, sample)
    u = np.stack([(rng_bits(k0, k1, d) >> np.uint32(8)).astype(np.float64) / 16777216.0 for d in range(16)])
    assert 0.0 <= u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 2e-3
    hist = np.histogram(u, bins=64, range=(0, 1))[0]
    assert abs(hist / hist.mean() - 1).max() < 0.02
    c = np.corrcoef(u)                                     # between dimensions, over paths
    assert np.abs(c - np.eye(16)).max() < 0.02
    by_px = u[0].reshape(2000, 50)                         # between neighbouring samples / pixels of one dimension
    assert abs(np.corrcoef(by_px[:, 0], by_px[:, 1])[0, 1]) < 0.08
    assert abs(np.corrcoef(by_px[:-1, 0], by_px[1:, 0])[0, 1]) < 0.08


def test_shading_sincos_is_accurate():
    """include/crt_math.h against double precision over the angles the path tracer passes: [-pi, 2 pi)."""
    worst = 0.0
    for a in np.linspace(-3.1416, 6.2832, 40001, dtype=np.float32):
        s, c = orc.sincos(a)
        for got, want in ((s, np.sin(np.float64(a))), (c, np.cos(np.float64(a)))):
            worst = max(worst, abs(got - want) / float(np.spacing(np.float32(max(abs(want), 1e-30)))))
            assert abs(got - want) < 2e-7
    assert worst < 2.0
    assert orc.sincos(0.0) == (0.0, 1.0)
    s, c = orc.sincos(np.float32(np.pi / 2))
    assert s == 1.0 and abs(c) < 1e-7
