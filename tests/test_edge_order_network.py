"""The compare-exchange schedule of k_edge_order (liodom_b200/csrc/register.cu), restated in NumPy: 64-key blocks whose
stages at distance <= 32 run per block (distance 32 between a lane's two registers, <= 16 by lane exchange) and
shared-memory stages at distance >= 64.  The kernel's output only orders work (any permutation gives the same poses),
so the GPU parity tests cannot tell a schedule that fails to sort from one that does; this does."""
import numpy as np
import pytest


def edge_order_network(keys):
    n = len(keys)
    np2 = 64
    while np2 < n:
        np2 *= 2
    sk = np.full(np2, np.iinfo(np.uint64).max, np.uint64)
    sk[:n] = keys
    ln = np.arange(32)

    def block_stages(k, from32):
        for b in range(np2 >> 6):
            e0 = (b << 6) + ln
            e1 = e0 + 32
            a, c = sk[e0].copy(), sk[e1].copy()
            if from32:
                sw = (a > c) == ((e0 & k) == 0)
                a, c = np.where(sw, c, a), np.where(sw, a, c)
            asc_a, asc_c = (e0 & k) == 0, (e1 & k) == 0
            j = 16 if from32 else min(k >> 1, 16)
            while j > 0:
                oa, oc = a[ln ^ j], c[ln ^ j]
                lower = (ln & j) == 0
                a = np.where((lower == asc_a) == (a < oa), a, oa)
                c = np.where((lower == asc_c) == (c < oc), c, oc)
                j >>= 1
            sk[e0], sk[e1] = a, c

    k = 2
    while k <= 32:
        block_stages(k, False)
        k <<= 1
    k = 64
    while k <= np2:
        j = k >> 1
        while j >= 64:
            t = np.arange(np2 >> 1)
            lo = ((t & ~(j - 1)) << 1) | (t & (j - 1))
            hi = lo | j
            a, b = sk[lo].copy(), sk[hi].copy()
            sw = (a > b) == ((lo & k) == 0)
            sk[lo], sk[hi] = np.where(sw, b, a), np.where(sw, a, b)
            j >>= 1
        block_stages(k, True)
        k <<= 1
    return sk[:n]


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 500, 1024, 1536, 2047, 2048])
def test_schedule_sorts(n):
    rng = np.random.default_rng(n)
    for _ in range(3):
        # 30-bit Morton key << 32 | index, as the kernel packs them (unique, so any correct network gives one result)
        keys = (rng.integers(0, 1 << 30, n).astype(np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
        assert np.array_equal(edge_order_network(keys), np.sort(keys))
