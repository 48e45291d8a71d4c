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
    for v in cases:
        arr = (C.c_uint32 * 16)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(16)])
        r = limbs(0)
        he.he_reduce512(r, arr)
        assert val(r) % P == v % P, hex(v)


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
