"""Host logic (no GPU): the column-wise canonicalisation of x = N / det (one big gcd per column instead
of one per entry, slip_solve.c canonical_column) gives exactly GMP's canonical form, which is what
the reference's mpq_div (slip_array_div.c) produces."""
import ctypes as C
import random
from fractions import Fraction

import pytest

from slip_lu_b200 import capi


def run(product, nums, det):
    n = len(nums)
    x = product.dll.SLIP_create_mpq_mat(n, 1)
    g = capi.gmp()
    for t, v in enumerate(nums):
        capi.int_to_mpz(x[t][0]._mp_num, v)
    d = (capi.MpzStruct * 1)()
    g.mpz_init(C.byref(d[0]))
    capi.int_to_mpz(d[0], det)
    f = product.dll.slip_b200_canonical_column_selftest
    f.restype = C.c_int
    f.argtypes = [C.POINTER(C.POINTER(capi.MpqStruct)), C.c_int32, C.POINTER(capi.MpzStruct)]
    assert f(x, n, d) == 0
    out = product.mpq_mat_to_py(x, n, 1)
    product.free_mpq_mat(x, n, 1)
    g.mpz_clear(C.byref(d[0]))
    return [r[0] for r in out]


def expect(nums, det):
    out = []
    for v in nums:
        f = Fraction(v, det)
        out.append((f.numerator, f.denominator))
    return out


BIG_PRIMES = [(1 << 61) - 1, (1 << 89) - 1, (1 << 127) - 1, 1000000007, 998244353]


@pytest.mark.parametrize("seed", range(8))
def test_canonical_column_equals_fraction_normal_form(product, seed):
    rng = random.Random(seed)
    # determinant with small prime powers, primes just below and above the trial-division bound, big primes
    det = rng.choice([1, -1])
    for p in (2, 3, 5, 7, 65521, 65537):
        det *= p ** rng.randrange(0, 4)
    for p in BIG_PRIMES:
        if rng.random() < 0.6:
            det *= p ** rng.randrange(1, 3)
    det *= rng.getrandbits(200) | 1
    nums = []
    for t in range(200):
        v = rng.getrandbits(rng.choice([10, 64, 300, 900])) - (1 << 8)
        for p in (2, 3, 5, 65521, 65537) + tuple(BIG_PRIMES):
            if rng.random() < 0.15:
                v *= p
        if rng.random() < 0.05:
            v = 0
        if rng.random() < 0.05:
            v = det * rng.randrange(-3, 4)
        nums.append(v)
    assert run(product, nums, det) == expect(nums, det)


def test_canonical_column_coprime_fast_path(product):
    """No numerator shares anything with the rough part: the single gcd proves it for all of them."""
    rng = random.Random(99)
    det = -(BIG_PRIMES[2] * BIG_PRIMES[1] * 12)
    nums = [rng.getrandbits(400) * 2 + 1 for _ in range(300)]
    nums = [v for v in nums if v % BIG_PRIMES[2] and v % BIG_PRIMES[1]]
    assert run(product, nums, det) == expect(nums, det)


def test_canonical_column_unit_determinant(product):
    assert run(product, [5, -7, 0, 12], -1) == [(-5, 1), (7, 1), (0, 1), (-12, 1)]
