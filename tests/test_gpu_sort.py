"""The hand-written onesweep radix sort against numpy's stable sort."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 31, 2048, 2049, 100_000, 3_000_001])
@pytest.mark.parametrize("bits", [(0, 64), (24, 64), (32, 64), (0, 8), (5, 18)])
def test_sort_pairs_matches_numpy_stable_sort(rjb, n, bits):
    rng = np.random.default_rng(n + bits[0])
    ctx = rjb.Context(device=0)
    keys = rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if n > 1000:  # heavy duplicates in some digits
        keys[: n // 3] &= np.uint64(0xFFFF0000FFFF0000)
    vals = np.arange(n, dtype=np.uint32)
    gk, gv = ctx.debug_sort_pairs(keys, vals, *bits)
    b, e = bits
    mask = np.uint64(((1 << (e - b)) - 1) if e - b < 64 else 0xFFFFFFFFFFFFFFFF)
    digit = (keys >> np.uint64(b)) & mask
    order = np.argsort(digit, kind="stable")
    assert np.array_equal(gv, vals[order])
    assert np.array_equal(gk, keys[order])
    ctx.close()


@pytest.mark.parametrize("n", [1, 33, 4096, 4097, 250_000, 5_000_001])
@pytest.mark.parametrize("bits", [(0, 32), (8, 32), (0, 8), (3, 27), (0, 30)])
def test_packed_sort_matches_numpy_stable_sort(rjb, n, bits):
    """The packed-pair sort of every product path (LBVH build, grid build, query ordering,
    overlay): key in the high half of a 64-bit word, payload in the low half."""
    rng = np.random.default_rng(7 * n + bits[0])
    ctx = rjb.Context(device=0)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64)
    if n > 1000:  # heavy duplicates in some digits
        keys[: n // 3] &= np.uint64(0xFF00FF00)
    words = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    got = ctx.debug_sort_packed(words, *bits)
    b, e = bits
    digit = (keys >> np.uint64(b)) & np.uint64((1 << (e - b)) - 1)
    order = np.argsort(digit, kind="stable")
    assert np.array_equal(got, words[order])
    ctx.close()
