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


def test_halving_butterfly_delivers_every_total_to_its_slot():
    """Emulation of warp_sum16 / warp_slot (csrc/sde_common.cuh): 16 values per lane, offsets 16, 8, 4, 2 with the kept
    half selected by the lane bit, final exchange at offset 1 -- lane l must end with the warp total of value warp_slot(l)."""
    rng = np.random.default_rng(1)
    v = rng.integers(-1000, 1000, size=(32, 16)).astype(np.int64)   # integers: order-independent reference
    total = v.sum(axis=0)
    lanes = np.arange(32)
    cur = v.copy()
    half, bit = 8, 16
    while half >= 1:
        up = (lanes & bit) != 0
        keep = np.where(up[:, None], cur[:, half:2 * half], cur[:, :half])
        send = np.where(up[:, None], cur[:, :half], cur[:, half:2 * half])
        cur = keep + send[lanes ^ bit]
        half >>= 1
        bit >>= 1
    res = cur[:, 0] + cur[lanes ^ 1, 0]
    slot = ((lanes >> 4) & 1) * 8 + ((lanes >> 3) & 1) * 4 + ((lanes >> 2) & 1) * 2 + ((lanes >> 1) & 1)
    assert np.array_equal(res, total[slot])
    assert sorted(set(slot[(lanes & 1) == 0])) == list(range(16))   # the even lanes cover every slot once
