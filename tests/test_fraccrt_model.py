"""Host-side model of the approximate magnitude used by the pivot search (k_fraccrt):
for X = x mod M, M = p_0...p_{s-1}, |x| < M/4,

    F = sum_i c_i * floor(2^(32W) / p_i)  mod 2^(32W),   c_i = x_i * (M/p_i)^-1 mod p_i,

lies within s*2^31 units below 2^(32W) * frac(X/M) (circularly), so g = min(F, 2^(32W) - F) is
2^(32W)*|x|/M to that accuracy.  The CUDA kernel computes exactly this F; the selection kernel's
proof rule relies on the bound.  Pure Python integers, no GPU."""
import random


def _primes_below_2_31(count):
    def is_prime(v):
        if v % 2 == 0:
            return False
        d, r = v - 1, 0
        while d % 2 == 0:
            d //= 2; r += 1
        for a in (2, 3, 5, 7):
            y = pow(a, d, v)
            if y in (1, v - 1):
                continue
            for _ in range(r - 1):
                y = y * y % v
                if y == v - 1:
                    break
            else:
                return False
        return True
    out, v = [], (1 << 31) - 1
    while len(out) < count:
        if is_prime(v):
            out.append(v)
        v -= 2
    return out


def test_fractional_crt_error_bound():
    rng = random.Random(7)
    for s, W in ((20, 8), (40, 12), (64, 30)):
        P = _primes_below_2_31(s)
        M = 1
        for p in P:
            M *= p
        U = [(1 << (32 * W)) // p for p in P]
        Minv = [pow(M // p % p, -1, p) for p in P]
        one = 1 << (32 * W)
        for trial in range(60):
            bits = rng.randrange(1, M.bit_length() - 3)
            x = rng.getrandbits(bits) * rng.choice((-1, 1))
            if trial == 0:
                x = 1
            if trial == 1:
                x = -1
            assert 4 * abs(x) < M
            c = [(x % p) * mi % p for p, mi in zip(P, Minv)]
            F = sum(ci * ui for ci, ui in zip(c, U)) % one
            true = (x % M) * one // M                      # floor of 2^(32W) * frac(X/M)
            delta = (true - F) % one                       # circular distance, true is at or above F
            assert delta <= s * (1 << 31), (s, W, x.bit_length(), delta.bit_length())
            g = min(F, one - F)
            exact = abs(x) * one // M
            assert abs(g - exact) <= s * (1 << 31) + 1
            # sign: the top bit of F, whenever the magnitude is resolved at all
            if exact > s * (1 << 32):
                assert (F >> (32 * W - 1)) == (1 if x < 0 else 0)


def test_proof_rule_orders_correctly():
    """Two magnitudes whose 96-bit windows at the larger one's leading word (at least five words
    above the bottom) differ by >= 2 units are ordered as the windows say."""
    rng = random.Random(11)
    s, W = 32, 16
    P = _primes_below_2_31(s)
    M = 1
    for p in P:
        M *= p
    U = [(1 << (32 * W)) // p for p in P]
    Minv = [pow(M // p % p, -1, p) for p in P]
    one = 1 << (32 * W)

    def key(x):
        c = [(x % p) * mi % p for p, mi in zip(P, Minv)]
        F = sum(ci * ui for ci, ui in zip(c, U)) % one
        g = F if F < one // 2 else one - 1 - F            # one's complement, as the kernel does
        words = [(g >> (32 * (W - 1 - w))) & 0xffffffff for w in range(W)]
        lead = next((w for w in range(W) if words[w]), W)
        return lead, (words + [0, 0, 0, 0])[lead:lead + 4]

    def proven_less(small, large):
        (ls, ks), (ll, kl) = small, large
        if ll > W - 6 or ls < ll:
            return False
        d = ls - ll
        L = (kl[0] << 64) | (kl[1] << 32) | kl[2]
        S = [(ks[0] << 64) | (ks[1] << 32) | ks[2], (ks[0] << 32) | ks[1], ks[0]][d] if d < 3 else 0
        return L > S and L - S >= 2

    checked = 0
    for _ in range(400):
        bits = rng.randrange(M.bit_length() - 8 * 32, M.bit_length() - 3)
        a = rng.getrandbits(bits) or 1
        b = a + rng.choice((0, 1, 2, rng.getrandbits(max(1, bits - rng.randrange(1, 120)))))
        ka, kb = key(a * rng.choice((-1, 1))), key(b * rng.choice((-1, 1)))
        if proven_less(ka, kb):
            assert a < b
            checked += 1
        if proven_less(kb, ka):
            assert b < a
            checked += 1
    assert checked > 50
