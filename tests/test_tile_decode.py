"""Host-side check of the division-free tile decoding of the CUDA kernels (csrc/mono_params.cuh: small_divmod):
q = trunc((float(t) + 0.5f) * fl(1 / d)) must equal t // d for every tile index the host accepts (t < 2^22).
The arithmetic is emulated in numpy float32 (same IEEE single-precision rounding as the device code)."""
import numpy as np


def small_div(t, d):
    rd = np.float32(1.0) / np.float32(d)
    return ((t.astype(np.float32) + np.float32(0.5)) * rd).astype(np.int64)


def test_small_divmod_is_exact_below_2_pow_22():
    rng = np.random.default_rng(0)
    limit = 1 << 22
    divisors = np.unique(np.concatenate([np.arange(1, 130), rng.integers(130, 5000, 200), [4096, 8191, 65535]]))
    for d in divisors:
        ks = np.unique(np.concatenate([np.arange(0, 64), rng.integers(0, limit // d + 1, 4000), [limit // d]]))
        # multiples of d and their neighbours are where a rounding error would flip the quotient
        t = np.unique(np.clip(np.concatenate([ks * d - 1, ks * d, ks * d + 1, rng.integers(0, limit, 4000)]), 0, limit - 1))
        assert np.array_equal(small_div(t, d), t // d), f"divisor {d}"


def test_small_divmod_exhaustive_for_the_bench_shapes():
    # tile counts of cfg2 / cfg3 / cfg4 (forward and backward tiles), every index of a scale
    for d, n in ((10, 1440), (12, 1440), (11, 1848), (14, 1848), (16, 2560), (20, 2560), (30, 19200), (80, 19200), (32, 23552), (92, 23552)):
        t = np.arange(n)
        assert np.array_equal(small_div(t, d), t // d)
