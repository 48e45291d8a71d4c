"""The limb arithmetic of bsgs-cuda_b200/csrc/fp.cuh, compiled for the HOST with the PTX carry
primitives emulated (csrc/fp_hostemu.cpp), against Python big ints.  CPU-only check of the exact
code the kernel runs; the same cases run on the device in tests/test_gpu_parity.py."""
import ctypes as C
import os
import random
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**256 - 2**32 - 977
M64 = 2**64 - 1


@pytest.fixture(scope="module")
def he():
    d = os.path.join(ROOT, "bsgs-cuda_b200")
    subprocess.check_call(["make", "-C", d, "libfp_hostemu.so"], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(d, "libfp_hostemu.so"))


def limbs(x):
    return (C.c_uint32 * 8)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def val(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


EDGE = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2, 2**255, 977, 2**32 + 977, 2**224 - 1,
        (2**256 - 1) ^ (2**64 - 1), 2**256 - 2**96, 2**256 - 2**32 - 978, P - 977]


def test_mul_sqr_lazy_inputs(he):
    rnd = random.Random(1)
    vals = EDGE + [rnd.getrandbits(256) for _ in range(400)]
    for a in vals:
        b = vals[rnd.randrange(len(vals))]
        r = limbs(0)
        he.he_mul(r, limbs(a), limbs(b)); assert val(r) % P == a * b % P
        he.he_sqr(r, limbs(a)); assert val(r) % P == a * a % P
    for a in EDGE:                                      # all edge pairs
        for b in EDGE:
            r = limbs(0); he.he_mul(r, limbs(a), limbs(b)); assert val(r) % P == a * b % P


def test_no_carry_leaves_a_chain_declared_carry_free(he):
    """fe_mul skips the carry capture of 9 of its 16 rows and fe_sqr of all its cross-product chains,
    because their top slot is fresh (a*b + carry < 2^64).  The host build counts any carry that such a
    chain drops; worst-case operands (all-ones limbs, single saturated limbs) and random ones."""
    he.he_lost_carries.restype = C.c_uint64
    rnd = random.Random(21)
    F = 0xFFFFFFFF
    sat = [2**256 - 1] + [F << (32 * i) for i in range(8)] + [(2**256 - 1) ^ (F << (32 * i)) for i in range(8)]
    sat += [sum(F << (32 * i) for i in range(0, 8, 2)), sum(F << (32 * i) for i in range(1, 8, 2))]
    vals = sat + EDGE + [rnd.getrandbits(256) for _ in range(200)]
    r = limbs(0)
    for a in vals:
        he.he_sqr(r, limbs(a)); assert val(r) % P == a * a % P
        for b in sat + [rnd.getrandbits(256) for _ in range(4)]:
            he.he_mul(r, limbs(a), limbs(b)); assert val(r) % P == a * b % P
    assert he.he_lost_carries() == 0


def test_sub_add_canon(he):
    rnd = random.Random(2)
    vals = [v % P for v in EDGE] + [rnd.randrange(P) for _ in range(400)]
    for a in vals:
        b = vals[rnd.randrange(len(vals))]
        r = limbs(0)
        he.he_sub_canon(r, limbs(a), limbs(b)); assert val(r) == (a - b) % P
        he.he_add_lazy(r, limbs(a), limbs(b)); assert val(r) < 2**256 and val(r) % P == (a + b) % P
    for a in EDGE:
        r = limbs(a); he.he_canon(r)
        if a < 2 * P:
            assert val(r) == a % P


def test_inverse_chain(he):
    rnd = random.Random(3)
    for a in [1, 2, P - 1, 977] + [rnd.randrange(1, P) for _ in range(20)]:
        r = limbs(0); he.he_inv(r, limbs(a)); assert val(r) % P == pow(a, P - 2, P)
    r = limbs(5); he.he_inv(r, limbs(0)); assert val(r) % P == 0      # inv(0) = 0 like m_Ecc_modInvX64 (Curve64.pb:2475-2477)


def test_low64_of_canonical_sum(he):
    he.he_add_low64_canon.restype = C.c_uint64
    rnd = random.Random(4)
    hi = (2**256 - 1) ^ (2**96 - 1)
    cases = []
    for _ in range(300):
        cases.append((rnd.getrandbits(256), rnd.randrange(P)))
        cases.append((hi | rnd.getrandbits(96), rnd.randrange(P)))                 # exact slow path
        cases.append((rnd.getrandbits(256), P - 1 - rnd.getrandbits(40)))
        cases.append((2**256 - 1 - rnd.getrandbits(34), P - 1 - rnd.getrandbits(34)))
        cases.append(((0xFFFFFFFF << 224) | rnd.getrandbits(224), rnd.getrandbits(200)))     # top limb of the sum all ones, the others not
        cases.append((rnd.getrandbits(256), ((0xFFFFFFFF << 224) | rnd.getrandbits(224)) % P))
    cases += [(a, u % P) for a in EDGE for u in EDGE]
    for a, u in cases:
        assert he.he_add_low64_canon(limbs(a), limbs(u)) == ((a + u) % P) & M64, (hex(a), hex(u))


def test_reduce512_all_branches(he):
    """Arbitrary 512-bit inputs (not only products), including the crafted cases in which the second
    fold wraps 2^256 again and the one in which the first fold overflows by more than 32 bits."""
    import ctypes as C
    C0 = 2**32 + 977
    rnd = random.Random(9)
    cases = [rnd.getrandbits(512) for _ in range(2000)] + [2**512 - 1, 0, 2**256 - 1, (2**256 - 1) << 256, P * P, (P - 1) * (2**256 - 1)]
    for _ in range(300):                                 # second addition wraps: (s mod 2^256) within f of 2^256
        hi = rnd.getrandbits(256) | (1 << 255)
        T = (hi * C0) >> 256
        want_low = 2**256 - 1 - rnd.randrange(0, T * C0 + 1)
        lo = (want_low - hi * C0) % 2**256
        cases.append((hi << 256) | lo)
    for _ in range(100):                                 # top2 = 1
        hi = 2**256 - 1 - rnd.getrandbits(20)
        cases.append((hi << 256) | (2**256 - 1 - rnd.getrandbits(40)))
    he.he_rare_calls.restype = C.c_uint64
    n_plain = len(cases)
    crafted = []
    for j in range(300):                                 # carry out of limb 2 in the second fold (top2 = 0): limb 2 of s all ones,
        hi = rnd.getrandbits(256) >> 1                   # limbs 0..1 of s within top*C of wrapping; with limbs 3..7 all ones too: full wrap
        T = (hi * C0) >> 256
        low = 2**64 - 1 - rnd.randrange(0, T * C0 + 1 if T else 1)
        upper = (2**160 - 1) if j % 3 == 0 else rnd.getrandbits(160)
        want_low = (upper << 96) | (0xFFFFFFFF << 64) | (low & (2**64 - 1))
        lo = (want_low - hi * C0) % 2**256
        crafted.append((hi << 256) | lo)
    before = he.he_rare_calls()
    for i, v in enumerate(cases + crafted):
        if i == n_plain:
            mid = he.he_rare_calls()
        arr = (C.c_uint32 * 16)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(16)])
        r = limbs(0)
        he.he_reduce512(r, arr)
        assert val(r) % P == v % P, hex(v)
    assert mid - before >= 300 and he.he_rare_calls() - mid >= 100      # both groups of crafted inputs reach the out-of-line tail


def test_key_low64_by_top_limb_classification(he):
    """fe_key_low64: low 64 bits of canonical (L + A - Q) from the low 64 bits and the top limbs only.
    Whenever it does not ask for the exact path its answer must be exact -- checked on random inputs
    and on inputs crafted around every boundary of the classification (S = L7 + A7 - Q7 near 0, 2^32,
    2^33) with the low 224 bits at their extremes."""
    he.he_key_low64.restype = C.c_uint64
    rnd = random.Random(12)
    B = 1 << 32
    lows = [0, 1, 2**224 - 1, 2**224 - 2, 2**223, 2**64 - 1, 2**64, 977]
    cases = []
    for _ in range(3000):
        cases.append((rnd.getrandbits(256), rnd.randrange(P), rnd.randrange(P)))
    for S in [-3, -2, -1, 0, 1, 2, B - 4, B - 3, B - 2, B - 1, B, B + 1, B + 2, 2 * B - 5, 2 * B - 4, 2 * B - 3, 2 * B - 2]:
        for _ in range(40):
            # pick top limbs with L7 + A7 - Q7 == S
            for _try in range(100):
                l7, q7 = rnd.randrange(B), rnd.randrange(B)
                a7 = S - l7 + q7
                if 0 <= a7 < B:
                    break
            else:
                continue
            lo = [rnd.choice(lows + [rnd.getrandbits(224)]) for _ in range(3)]
            L = (l7 << 224) | lo[0]; A = (a7 << 224) | lo[1]; Q = (q7 << 224) | lo[2]
            if A < P and Q < P:
                cases.append((L, A, Q))
    assert len(cases) > 3300
    asked = 0
    for L, A, Q in cases:
        ex = C.c_int()
        k = he.he_key_low64(limbs(L), limbs(A), limbs(Q), C.byref(ex))
        S = (L >> 224) + (A >> 224) - (Q >> 224)
        want_exact = S in (-1, 0, B - 2, B - 1, B, 2 * B - 3, 2 * B - 2)
        assert bool(ex.value) == want_exact, hex(S)
        if ex.value:
            asked += 1
        else:
            assert k == ((L + A - Q) % P) & M64, (hex(L), hex(A), hex(Q))
    assert 0 < asked < len(cases) // 3


def limbs16(x):
    return (C.c_uint32 * 16)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(16)])


def _wide_case_with_s(rnd, s_target, extreme):
    """A triple (v, A, Q) whose classification word s = S7 mod 2^32 equals s_target; `extreme` pushes the
    part below 2^224 (the unknown carry c) to its minimum (-1) or maximum (4)."""
    C0 = 977
    while True:
        hi = rnd.getrandbits(256)
        if extreme == "max":                     # everything below 2^224 as large as possible
            hi |= (1 << 192) - 1; hi |= 0xFFFFFFFF << 192
        elif extreme == "min":
            hi &= ~((1 << 192) - 1); hi &= ~(0xFFFFFFFF << 192)
        hi6, hi7 = (hi >> 192) & 0xFFFFFFFF, hi >> 224
        a7, q7 = rnd.randrange(1 << 32), rnd.randrange(1 << 32)
        a_low = (2**224 - 1) if extreme == "max" else 0 if extreme == "min" else rnd.getrandbits(224)
        q_low = 0 if extreme == "max" else (2**224 - 1) if extreme == "min" else rnd.getrandbits(224)
        A, Q = (a7 << 224) | a_low, (q7 << 224) | q_low
        if not (A < P and Q < P):
            continue
        lo_low = (2**224 - 1) if extreme == "max" else 0 if extreme == "min" else rnd.getrandbits(224)
        base = (a7 - q7) + hi6 + ((C0 * hi7) & 0xFFFFFFFF) + ((C0 * hi6) >> 32)
        lo7 = (s_target - base) % 2**32
        return (hi << 256) | (lo7 << 224) | lo_low, A, Q


def test_key_from_unreduced_square(he):
    """fe_key_from_wide: low 64 bits of canonical (v + A - Q) mod p straight from a 512-bit v.  Exact whenever
    it does not ask for the exact path; asks exactly for s outside [2, 2^32 - 6]; the bounds of the carry
    from below 2^224 (c in [-1, 4]) are attained by the crafted extremes, on both sides of every boundary."""
    he.he_key_from_wide.restype = C.c_uint64
    rnd = random.Random(33)
    THR = 0xFFFFFFF8
    cases = []
    for _ in range(3000):
        cases.append((rnd.getrandbits(512), rnd.randrange(P), rnd.randrange(P)))
    for a in EDGE + [rnd.getrandbits(256) for _ in range(300)]:      # real squares, incl. the largest
        cases.append((a * a, rnd.randrange(P), rnd.randrange(P)))
    cases.append(((2**256 - 1) ** 2, P - 1, 0)); cases.append(((2**256 - 1) ** 2, 0, P - 1)); cases.append((0, 0, P - 1)); cases.append((0, 0, 0))
    B = 1 << 32
    for s in [0, 1, 2, 3, 4, 5, 6, 7, B - 9, B - 8, B - 7, B - 6, B - 5, B - 4, B - 3, B - 2, B - 1]:
        for extreme in ("max", "min", "rnd"):
            for _ in range(25):
                cases.append(_wide_case_with_s(rnd, s, extreme))
    asked = 0
    for v, A, Q in cases:
        ex = C.c_int()
        k = he.he_key_from_wide(limbs16(v), limbs(A), limbs(Q), THR, C.byref(ex))
        hi, lo = v >> 256, v & (2**256 - 1)
        S7 = (lo >> 224) + ((A >> 224) - (Q >> 224)) + ((hi >> 192) & 0xFFFFFFFF) + ((977 * (hi >> 224)) & 0xFFFFFFFF) + ((977 * ((hi >> 192) & 0xFFFFFFFF)) >> 32)
        s = S7 % B
        assert bool(ex.value) == (not 2 <= s <= B - 6), hex(s)
        if ex.value:
            asked += 1
        else:
            assert k == ((v + A - Q) % P) & M64, (hex(v), hex(A), hex(Q))
    assert 7 * 75 <= asked < 7 * 75 + 20                       # s in {0, 1, 2^32-5 .. 2^32-1}: 7 of the 17 crafted values
    # the carry bounds are tight: c = -1 and c = 4 both occur
    def carry(v, A, Q):
        hi, lo = v >> 256, v & (2**256 - 1)
        U = hi * (2**32 + 977) + lo + A - Q
        S7 = (lo >> 224) + ((A >> 224) - (Q >> 224)) + ((hi >> 192) & 0xFFFFFFFF) + ((977 * (hi >> 224)) & 0xFFFFFFFF) + ((977 * ((hi >> 192) & 0xFFFFFFFF)) >> 32)
        M = (hi >> 224) + ((977 * (hi >> 224)) >> 32)
        return (U >> 224) - (M * B + S7)
    cs = {carry(*_wide_case_with_s(rnd, 12345, e)) for e in ("max", "min") for _ in range(20)}
    assert min(cs) == -1 and max(cs) >= 3 and all(-1 <= c <= 4 for c in cs)
    # thr = 0 (the forced-exact test mode of the library): everything but s == 2 asks for the exact path
    ex = C.c_int()
    he.he_key_from_wide(limbs16(cases[0][0]), limbs(cases[0][1]), limbs(cases[0][2]), 0, C.byref(ex))
    assert ex.value == 1


def test_sqr_wide_is_the_full_square(he):
    rnd = random.Random(34)
    for a in EDGE + [rnd.getrandbits(256) for _ in range(300)]:
        v = (C.c_uint32 * 16)()
        he.he_sqr_wide(v, limbs(a))
        assert sum(int(v[i]) << (32 * i) for i in range(16)) == a * a
