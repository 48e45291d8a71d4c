"""GPU parity tests proper: the CUDA path (through the C ABI of libbsgs_b200.so) against the CPU
oracle, and against the reference's own kernel (oracle/_ref, when present) on identical inputs.
Bit-exact: integer work, sorted multisets of (code, idx)."""
import os
import random
import numpy as np
import pytest
from tests import pyec
from tests.helpers import M64, random_point, giant, candidate_x, synthetic_keys, norm_hits, build_case

pytestmark = pytest.mark.gpu
P = pyec.P


@pytest.fixture(scope="module")
def pkg():
    import bsgs_cuda_b200 as b
    b.native.load()
    return b


# ------------------------------------------------------------------ limb arithmetic on the device
def test_field_ops_on_device(pkg):
    rnd = random.Random(7)
    edge = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2, 2**255, 977, 2**32 + 977, 2**224 - 1,
            (2**256 - 1) ^ (2**64 - 1), 2**256 - 2**96, 2**256 - 2**32 - 978, P - 977]
    a = edge + [rnd.getrandbits(256) for _ in range(3000)]
    b = [a[rnd.randrange(len(a))] for _ in a]
    dbg = pkg.native.debug_field
    assert [x % P for x in dbg(0, a, b)] == [x * y % P for x, y in zip(a, b)]
    assert [x % P for x in dbg(1, a, b)] == [x * x % P for x in a]
    ac, bc = [x % P for x in a], [x % P for x in b]
    assert dbg(3, ac, bc) == [(x - y) % P for x, y in zip(ac, bc)]
    r = dbg(4, ac, bc)
    assert all(v < 2**256 for v in r) and [v % P for v in r] == [(x + y) % P for x, y in zip(ac, bc)]
    assert dbg(5, a, b) == [x % P if x < 2 * P else None for x in a]
    assert [v & M64 for v in dbg(6, a, bc)] == [((x + y) % P) & M64 for x, y in zip(a, bc)]
    sub = a[:256]
    assert [x % P for x in dbg(2, sub, sub)] == [pow(x % P, P - 2, P) for x in sub]
    # X(2P) low 64 bits from (Px, 1/(2Py))
    pts = [random_point(rnd) for _ in range(32)]
    got = dbg(7, [p[0] for p in pts], [pow(2 * p[1], -1, P) for p in pts])
    assert [g & M64 for g in got] == [pyec.add(p, p)[0] & M64 for p in pts]


def test_reduction_rare_tail_on_device(pkg):
    """reduce512 with its out-of-line tail (about 2^-22 of the calls) ON THE DEVICE: 512-bit inputs crafted so
    that the first fold overflows by more than 32 bits (top2 = 1), so that the second fold carries out of limb
    2 (limb 2 of the folded value all ones), and so that 2^256 wraps once more (limbs 3..7 all ones too)."""
    rnd = random.Random(9)
    C0 = 2**32 + 977
    cases = [rnd.getrandbits(512) for _ in range(2000)] + [2**512 - 1, 0, 2**256 - 1, (2**256 - 1) << 256, P * P, (P - 1) * (2**256 - 1)]
    for _ in range(300):
        hi = rnd.getrandbits(256) | (1 << 255)
        T = (hi * C0) >> 256
        want_low = 2**256 - 1 - rnd.randrange(0, T * C0 + 1)
        cases.append((hi << 256) | ((want_low - hi * C0) % 2**256))
    for _ in range(100):                                             # top2 = 1
        hi = 2**256 - 1 - rnd.getrandbits(20)
        cases.append((hi << 256) | (2**256 - 1 - rnd.getrandbits(40)))
    for j in range(300):                                             # carry out of limb 2 / full wrap
        hi = rnd.getrandbits(256) >> 1
        T = (hi * C0) >> 256
        low = 2**64 - 1 - rnd.randrange(0, T * C0 + 1 if T else 1)
        upper = (2**160 - 1) if j % 3 == 0 else rnd.getrandbits(160)
        want_low = (upper << 96) | (0xFFFFFFFF << 64) | (low & (2**64 - 1))
        cases.append((hi << 256) | ((want_low - hi * C0) % 2**256))
    lo = [v & (2**256 - 1) for v in cases]; hi = [v >> 256 for v in cases]
    got = pkg.native.debug_field(8, lo, hi)
    assert all(g < 2**256 for g in got) and [g % P for g in got] == [v % P for v in cases]


def test_low64_slow_path_on_device(pkg):
    """Values whose upper limbs are all ones exercise the exact branch of fe_add_low64_canon."""
    rnd = random.Random(11)
    a, u = [], []
    for _ in range(500):
        hi = (2**256 - 1) ^ (2**96 - 1)
        a.append(hi | rnd.getrandbits(96)); u.append(rnd.randrange(P))
        a.append(rnd.getrandbits(256)); u.append(P - 1 - rnd.getrandbits(40))
        x = rnd.getrandbits(256); a.append(x); u.append((hi | rnd.getrandbits(96)) % P if x else 0)
        a.append(2**256 - 1 - rnd.getrandbits(34)); u.append(P - 1 - rnd.getrandbits(34))
    got = pkg.native.debug_field(6, a, u)
    assert [v & M64 for v in got] == [((x + y) % P) & M64 for x, y in zip(a, u)]


def test_hot_loop_key_extraction_on_device_incl_the_exact_path(pkg):
    """candidate_key (giant_step.cuh) ON THE DEVICE, as PASS 2 calls it: lambda is squared without reduction and
    the key read off a few limbs of the 512-bit square (fp::fe_key_from_wide); ~2^-29 of the candidates (about
    100 per second per GPU in production, essentially never in the other tests) take the out-of-line exact path.
    (a) 512-bit values crafted around every boundary of the classification word s (with the part below 2^224
    at both extremes of its carry) in place of the square; (b) 10^6 random lambdas through the real thing;
    (c) the same lambdas with the exact path forced."""
    from tests.test_fp_hostemu import _wide_case_with_s
    rnd = random.Random(12)
    Bq = 1 << 32
    vs, As, Qs, want_exact = [], [], [], []
    for s in [0, 1, 2, 3, 4, 5, 6, 7, Bq - 9, Bq - 8, Bq - 7, Bq - 6, Bq - 5, Bq - 4, Bq - 3, Bq - 2, Bq - 1]:
        for extreme in ("max", "min", "rnd"):
            for _ in range(120):
                v, A, Q = _wide_case_with_s(rnd, s, extreme)
                vs.append(v); As.append(A); Qs.append(Q); want_exact.append(not 2 <= s <= Bq - 6)
    for _ in range(2000):
        vs.append(rnd.getrandbits(512)); As.append(rnd.randrange(P)); Qs.append(rnd.randrange(P)); want_exact.append(None)
    keys, ex = pkg.native.debug_key(vs, As, Qs, wide=True)
    assert keys.tolist() == [((v + a - q) % P) & M64 for v, a, q in zip(vs, As, Qs)]
    assert all(w is None or bool(e) == w for e, w in zip(ex, want_exact)) and sum(1 for w in want_exact if w) >= 2500
    rs = np.random.RandomState(5)
    nrand = 1_000_000
    raw = rs.randint(0, 2**32, size=(3, nrand, 8), dtype=np.uint64).astype(np.uint32)
    raw[1:, :, 7] &= 0x7FFFFFFF                                       # A, Q < 2^255 < p
    def ints(a):
        return [int.from_bytes(row.tobytes(), "little") for row in a]
    Lr, Ar, Qr = ints(raw[0]), ints(raw[1]), ints(raw[2])
    want = [((l * l + a - q) % P) & M64 for l, a, q in zip(Lr, Ar, Qr)]
    keys, ex = pkg.native.debug_key(Lr, Ar, Qr)
    assert keys.tolist() == want
    assert int(ex.sum()) <= 8                                         # ~2^-29 each: random inputs stay on the fast path
    n2 = 100_000
    edge = [0, 1, P - 1, P, 2**256 - 1, 2**255, 2**128]
    keys, ex = pkg.native.debug_key(edge + Lr[:n2], [P - 1] * len(edge) + Ar[:n2], [0] * len(edge) + Qr[:n2], force=True)
    assert keys.tolist() == [((l * l + P - 1) % P) & M64 for l in edge] + want[:n2]
    assert int(ex.sum()) >= len(keys) - 4                             # forced: all but s == 2


def test_whole_launch_with_the_exact_key_path_forced(pkg, orc):
    """Option force_exact: every candidate of a launch goes through the out-of-line exact path of the hot loop
    (key_exact_sqr); the records must not change."""
    T, B, p, w, htsz = 64, 8, 16, 20000, 11
    Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=777)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        fast = norm_hits(e.step(Ppt))
        e._ck(e.L.bsgs_set_option(e.h, b"force_exact", 1), "force_exact")
        slow = norm_hits(e.step(Ppt))
        multi = sorted((wi, c, i) for c, i, wi in e.step_multi([Ppt, pyec.add(Ppt, pyec.G), Ppt]))
        e._ck(e.L.bsgs_set_option(e.h, b"force_exact", 0), "force_exact")
        assert fast == slow == norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt)) and set(plant) <= set(slow)
        assert [(c, i) for wi, c, i in multi if wi == 0] == [(c, i) for wi, c, i in multi if wi == 2] == fast


# ------------------------------------------------------------------ table layouts
@pytest.mark.parametrize("w,htsz", [(50000, 14), (50000, 9), (3000, 12), (1, 1), (200000, 13)])
def test_probe_layout_vs_oracle(pkg, orc, w, htsz):
    keys = synthetic_keys(w, [], seed=w + htsz)
    _, gpu = orc.build_ht(keys, htsz)
    with pkg.Engine(4, 4, 4, w, htsz) as e:
        e.upload_ht(gpu)
        rs = np.random.RandomState(3)
        q = np.concatenate([keys[rs.randint(0, w, size=4000)],                        # members
                            rs.randint(0, 2**62, size=4000).astype(np.uint64),             # random
                            keys[rs.randint(0, w, size=4000)] ^ np.uint64(1 << 40),        # same bucket, other tag
                            keys[rs.randint(0, w, size=2000)] ^ np.uint64(1 << 20)])       # bits the table ignores
        got = e.probe_keys(q)
        exp = np.array([orc.probe_htgpu(gpu, htsz, int(k)) for k in q])
        assert np.array_equal(got, exp)
        assert got[:4000].all()


# ------------------------------------------------------------------ one launch
CASES = [
    # T, B, p, w, htsz, blocks_per_sm
    (32, 4, 16, 4096, 10, 2),          # N = 2048: one point per thread
    (7, 3, 5, 1000, 8, 2),             # ragged N = 105
    (64, 20, 100, 60000, 14, 1),       # N = 128000 > one wave of threads: several points per thread
    (128, 2, 8, 5000, 4, 2),           # load 312 per bucket: every probe takes the CSR fallback
]


@pytest.mark.parametrize("case", CASES)
def test_step_matches_oracle(pkg, orc, case):
    T, B, p, w, htsz, bps = case
    Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=T * 1000 + p)
    exp = orc.launch(g2, T, B, p, gpu, htsz, Ppt)
    assert set(plant) <= set(exp)                      # planted hits are real hits of the oracle
    with pkg.Engine(T, B, p, w, htsz, blocks_per_sm=bps) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        assert norm_hits(e.step(Ppt)) == norm_hits(exp)
        # same engine, other base points (tables stay resident, like the reference's launch loop)
        rnd = random.Random(99)
        for _ in range(2):
            Q = random_point(rnd)
            assert norm_hits(e.step(Q)) == norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Q))


@pytest.mark.parametrize("nwin", [2, 3, 4, 7, 16])
def test_multi_window_equals_separate_launches(pkg, orc, nwin):
    T, B, p, w, htsz = 32, 8, 12, 20000, 12
    rnd = random.Random(nwin)
    pts = [random_point(rnd) for _ in range(nwin)]
    N = T * B * p
    plant = [(i % nwin, 1 + (i & 1), rnd.randrange(N)) for i in range(16)]
    xs = [candidate_x(pts[wi], w, k, c) for wi, c, k in plant]
    keys = synthetic_keys(w, xs, 5)
    cpu, gpu = orc.build_ht(keys, htsz)
    g2 = orc.build_g2(T, B, p, w)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        got = sorted((win, c, 0 if c == 5 else i) for c, i, win in e.step_multi(pts))
        exp = sorted((wi, c, 0 if c == 5 else i) for wi in range(nwin) for c, i in orc.launch(g2, T, B, p, gpu, htsz, pts[wi]))
        assert got == exp
        assert {(wi, c, k) for wi, c, k in plant} <= set(got)


def test_windows_8_small_chain(pkg, orc):
    T, B, p, w, htsz = 16, 4, 6, 3000, 9
    rnd = random.Random(8)
    pts = [random_point(rnd) for _ in range(8)]
    xs = [candidate_x(pts[i], w, (i * 37) % (T * B * p), 1 + (i & 1)) for i in range(8)]
    cpu, gpu = orc.build_ht(synthetic_keys(w, xs, 8), htsz)
    g2 = orc.build_g2(T, B, p, w)
    with pkg.Engine(T, B, p, w, htsz, max_windows=8) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        got = sorted((win, c, 0 if c == 5 else i) for c, i, win in e.step_multi(pts))
        exp = sorted((wi, c, 0 if c == 5 else i) for wi in range(8) for c, i in orc.launch(g2, T, B, p, gpu, htsz, pts[wi]))
        assert got == exp and len(got) >= 8


def test_planted_codes_4_and_5(pkg, orc):
    T, B, p, w, htsz = 16, 4, 8, 2000, 8
    N = T * B * p
    g2 = orc.build_g2(T, B, p, w)
    rnd = random.Random(45)
    for k, sign in ((0, 1), (N - 1, -1), (137, 1)):
        Q = giant(w, k)
        Ppt = Q if sign > 0 else pyec.neg(Q)                       # Px == Qx
        dbl = pyec.add(Ppt, Ppt)
        for planted in (True, False):
            keys = synthetic_keys(w, [dbl[0]] + ([Ppt[0]] if planted else []) if planted else [], k)
            cpu, gpu = orc.build_ht(keys, htsz)
            exp = orc.launch(g2, T, B, p, gpu, htsz, Ppt)
            if planted:
                assert (4, k) in exp and (5, 0) in exp
            with pkg.Engine(T, B, p, w, htsz) as e:
                e.upload_giants(g2); e.upload_ht(gpu)
                assert norm_hits(e.step(Ppt)) == norm_hits(exp)
                other = random_point(rnd)
                got = sorted((win, c, 0 if c == 5 else i) for c, i, win in e.step_multi([other, Ppt]))
                ex2 = sorted((wi, c, 0 if c == 5 else i) for wi, pt in enumerate([other, Ppt]) for c, i in orc.launch(g2, T, B, p, gpu, htsz, pt))
                assert got == ex2


# ------------------------------------------------------------------ the reference's own kernel on the same inputs
def _ref_available(orc):
    return os.path.exists(orc.REF_LIB) and os.path.exists(orc.REF_CUBIN)


@pytest.mark.parametrize("case", [(32, 4, 16, 4096, 10), (64, 20, 100, 60000, 14), (128, 2, 8, 5000, 4), (512, 8, 6, 30000, 13)])
def test_reference_kernel_oracle_and_product_agree(pkg, orc, case):
    if not _ref_available(orc):
        pytest.skip("oracle/_ref not built (reference tree absent at build time)")
    T, B, p, w, htsz = case
    Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=4242 + T)
    ref = orc.RefKernel(T, B, p, w, htsz)
    ref.upload(g2, gpu)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        rnd = random.Random(5)
        for pt in [Ppt] + [random_point(rnd) for _ in range(2)]:
            r_hits, r_count, _ = ref.launch(pt)
            assert r_count <= 240
            o_hits = norm_hits(orc.launch(g2, T, B, p, gpu, htsz, pt))
            assert r_hits == o_hits                                   # oracle pinned by the reference's kernel
            assert norm_hits(e.step(pt)) == r_hits                     # product == reference
    ref.close()


def test_reference_kernel_codes_4_5(pkg, orc):
    if not _ref_available(orc):
        pytest.skip("oracle/_ref not built")
    T, B, p, w, htsz = 32, 4, 8, 2000, 8
    N = T * B * p
    g2 = orc.build_g2(T, B, p, w)
    for k, sign in ((5, 1), (N - 1, -1)):
        Q = giant(w, k)
        Ppt = Q if sign > 0 else pyec.neg(Q)
        dbl = pyec.add(Ppt, Ppt)
        cpu, gpu = orc.build_ht(synthetic_keys(w, [dbl[0], Ppt[0]], k), htsz)
        ref = orc.RefKernel(T, B, p, w, htsz); ref.upload(g2, gpu)
        r_hits, _, _ = ref.launch(Ppt)
        ref.close()
        assert (4, k) in r_hits and (5, 0) in r_hits
        assert r_hits == norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt))
        with pkg.Engine(T, B, p, w, htsz) as e:
            e.upload_giants(g2); e.upload_ht(gpu)
            assert norm_hits(e.step(Ppt)) == r_hits


# ------------------------------------------------------------------ the drop-in module (boundary option A)
DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bsgs-cuda_b200", "dropin_test1.cubin")


def _harness_available(orc):
    return os.path.exists(orc.REF_LIB) and os.path.exists(DROPIN)


@pytest.mark.parametrize("case", [(32, 4, 16, 4096, 10), (64, 20, 100, 60000, 14), (128, 2, 8, 5000, 4), (512, 8, 6, 30000, 13), (256, 3, 1, 3000, 9)])
def test_dropin_module_matches_oracle_and_reference_kernel(orc, case):
    """bsgs-cuda_b200/dropin_test1.cubin exports the reference kernel's own interface (`_test1`, `_A`,
    the memory map of cuda(), 1_9_7File.pb:2181-2509): loaded by the same driver-API harness that
    runs the reference's PTX it must produce the same (code, idx) records."""
    if not _harness_available(orc):
        pytest.skip("harness or drop-in module not built")
    T, B, p, w, htsz = case
    Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=977 + T)
    mine = orc.RefKernel(T, B, p, w, htsz, cubin=DROPIN)
    mine.upload(g2, gpu)
    ref = None
    if _ref_available(orc):
        ref = orc.RefKernel(T, B, p, w, htsz); ref.upload(g2, gpu)
    rnd = random.Random(6)
    for pt in [Ppt] + [random_point(rnd) for _ in range(2)]:
        m_hits, m_count, _ = mine.launch(pt)
        assert m_hits == norm_hits(orc.launch(g2, T, B, p, gpu, htsz, pt))
        if ref is not None:
            r_hits, r_count, _ = ref.launch(pt)
            assert m_hits == r_hits and m_count == r_count
    assert set(plant) <= set(mine.launch(Ppt)[0])
    mine.close()
    if ref is not None:
        ref.close()


def test_dropin_module_codes_4_5_and_repeat_launches(orc):
    if not _harness_available(orc):
        pytest.skip("harness or drop-in module not built")
    T, B, p, w, htsz = 32, 4, 8, 2000, 8
    N = T * B * p
    g2 = orc.build_g2(T, B, p, w)
    for k, sign in ((5, 1), (N - 1, -1)):
        Q = giant(w, k)
        Ppt = Q if sign > 0 else pyec.neg(Q)
        dbl = pyec.add(Ppt, Ppt)
        cpu, gpu = orc.build_ht(synthetic_keys(w, [dbl[0], Ppt[0]], k), htsz)
        mine = orc.RefKernel(T, B, p, w, htsz, cubin=DROPIN); mine.upload(g2, gpu)
        first, _, _ = mine.launch(Ppt)
        again, _, _ = mine.launch(Ppt)                                 # the giant and chain blocks survive a launch
        mine.close()
        assert (4, k) in first and (5, 0) in first
        assert first == again == norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt))


def test_dropin_module_as_ptx_text_like_the_reference_embeds_it(orc, tmp_path):
    """The reference hands PTX TEXT to cuModuleLoadData (1_9_7File.pb:2194), de-obfuscated from its
    DataSection: take dropin_test1.ptx through scripts/make_datasection.py and back, let the driver
    JIT it, and compare records with the oracle."""
    import importlib.util
    ptx = DROPIN[:-len("cubin")] + "ptx"
    if not (os.path.exists(orc.REF_LIB) and os.path.exists(ptx)):
        pytest.skip("harness or drop-in PTX not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_datasection", os.path.join(root, "scripts", "make_datasection.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    text = m.decode(m.encode(open(ptx, "rb").read())).rstrip(b"\0")          # what the host holds after its XOR loop
    path = tmp_path / "embedded.ptx"; path.write_bytes(text)
    T, B, p, w, htsz = 64, 6, 10, 20000, 11
    Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=4321)
    k = orc.RefKernel(T, B, p, w, htsz, cubin=str(path)); k.upload(g2, gpu)
    got, count, _ = k.launch(Ppt)
    k.close()
    assert got == norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt)) and set(plant) <= set(got)


def test_dropin_module_record_area_is_bounded(orc):
    """More than 240 hits: the counter keeps counting, the giants just above the record area stay
    intact (the reference overwrites them, 1_9_7File.pb:2337) -- a second launch still agrees."""
    if not _harness_available(orc):
        pytest.skip("harness or drop-in module not built")
    T, B, p, w, htsz = 64, 4, 8, 4000, 8
    N = T * B * p
    g2 = orc.build_g2(T, B, p, w)
    rnd = random.Random(11)
    Ppt = random_point(rnd)
    planted = [candidate_x(Ppt, w, k, 1 + (k & 1)) for k in range(0, 600, 2)]
    cpu, gpu = orc.build_ht(synthetic_keys(w, planted, 3), htsz)
    expect = norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt, cap=4096))
    assert len(expect) >= 300
    mine = orc.RefKernel(T, B, p, w, htsz, cubin=DROPIN); mine.upload(g2, gpu)
    got, count, _ = mine.launch(Ppt)
    assert count == len(expect) and len(got) == 240 and set(got) <= set(expect)
    got2, count2, _ = mine.launch(Ppt)
    assert count2 == count and set(got2) <= set(expect)
    mine.close()


# ------------------------------------------------------------------ whole problems
def test_solve_small_range_true_tables(pkg, orc):
    """Real baby table + giants (built by the oracle on the CPU), key recovered through bsgs_run."""
    T, B, p, w, htsz = 32, 4, 8, 1 << 14, 12
    keys = orc.baby_keys(w)
    cpu, gpu = orc.build_ht(keys, htsz)
    g2 = orc.build_g2(T, B, p, w)
    rnd = random.Random(2026)
    pk = 1 << 35
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        for _ in range(6):
            secret = pk + rnd.randrange(1, 1 << 34)
            if (secret - pk - 1 - p * w) % (2 * w) == 0:
                secret += 1
            target = pyec.mul(pyec.G, secret)
            key, done, seen = pkg.solve(e, cpu, T, B, p, w, htsz, pk, pk + (1 << 34), target, batch=4)
            assert key == secret


def test_run_equals_stepwise(pkg, orc):
    T, B, p, w, htsz = 16, 4, 8, 5000, 9
    rnd = random.Random(77)
    s = pkg.Search(T, B, p, w, htsz, 12345, random_point(rnd))
    pts = [s.window_point(j) for j in range(7)]
    xs = [candidate_x(pts[j], w, (j * 53) % (T * B * p), 1 + (j & 1)) for j in range(7)]
    cpu, gpu = orc.build_ht(synthetic_keys(w, xs, 3), htsz)
    g2 = orc.build_g2(T, B, p, w)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        hits, done = e.run(s.P0, s.D, 7, batch=3)                 # two launches in flight (alternating lanes), ragged last batch
        assert done == 7
        exp = sorted((j, c, 0 if c == 5 else i) for j in range(7) for c, i in orc.launch(g2, T, B, p, gpu, htsz, pts[j]))
        assert sorted((win, c, 0 if c == 5 else i) for c, i, win in hits) == exp and len(exp) >= 7
        e._ck(e.L.bsgs_set_option(e.h, b"overlap", 0), "overlap")  # one launch at a time: same records
        hits1, done1 = e.run(s.P0, s.D, 7, batch=3)
        assert done1 == 7 and sorted(hits1) == sorted(hits)
        e._ck(e.L.bsgs_set_option(e.h, b"overlap", 1), "overlap")
        for batch in (1, 2, 7):
            h2, d2 = e.run(s.P0, s.D, 7, batch=batch)
            assert d2 == 7 and sorted(h2) == sorted(hits)
        # early stop: the windows of the launch already in flight are searched and counted too
        first = min(win for c, i, win in hits)
        h3, d3 = e.run(s.P0, s.D, 7, batch=1, stop_after_hits=1)
        assert first + 1 <= d3 <= first + 2 and sorted(h3) == sorted(x for x in hits if x[2] < d3)
        e.set_verify(False)
        ms = e.time_launches(pts[:6], 2, 3)                        # timing entry point alternates lanes as well
        assert ms > 0 and sorted((c, i) for c, i in e.step(pts[0])) == sorted((c, i) for c, i, win in hits if win == 0)


@pytest.mark.parametrize("opts", [dict(block=128), dict(block=64), dict(block=320), dict(blocks_per_sm=2), dict(blocks_per_sm=1)])
def test_every_launch_geometry_emits_the_same_records(pkg, orc, opts):
    """The kernel is instantiated for 256 x 3 (default), 256 x 2, 320 x 2, 128 x 6 and 64 x 12 threads per SM: same records."""
    T, B, p, w, htsz = 64, 20, 50, 30000, 11
    Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=99)
    want = norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt))
    with pkg.Engine(T, B, p, w, htsz, **opts) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        assert norm_hits(e.step(Ppt)) == want and set(plant) <= set(want)
        pts = [Ppt, pyec.add(Ppt, pyec.G), pyec.add(Ppt, pyec.mul(pyec.G, 5))]
        multi = sorted((wi, c, i) for c, i, wi in e.step_multi(pts))
        assert multi == sorted((wi, c, 0 if c == 5 else i) for wi, pt in enumerate(pts) for c, i in orc.launch(g2, T, B, p, gpu, htsz, pt))


def test_trace_diagnostics(pkg, orc):
    """bsgs_debug_trace: per-warp {SM id, start, end of pass 1, end} of the last launch, in order."""
    T, B, p, w, htsz = 64, 20, 50, 30000, 11
    Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=98)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        with pytest.raises(pkg.BsgsError):
            e.trace()                                              # not enabled
        e._ck(e.L.bsgs_set_option(e.h, b"trace", 1), "trace")
        hits = norm_hits(e.step(Ppt))
        tr = e.trace()
        info = e.info()
        assert tr.shape == (info["grid"] * info["block"] // 32, 4)
        assert (tr[:, 0] < info["sm_count"]).all() and (tr[:, 1] <= tr[:, 2]).all() and (tr[:, 2] <= tr[:, 3]).all() and (tr[:, 1] > 0).all()
        e._ck(e.L.bsgs_set_option(e.h, b"trace", 0), "trace")
        assert norm_hits(e.step(Ppt)) == hits


def test_error_behaviour(pkg):
    with pytest.raises(pkg.BsgsError):
        pkg.Engine(1 << 12, 1 << 12, 1 << 8, 1000, 10)              # N >= 2^31
    with pkg.Engine(8, 2, 2, 100, 4) as e:
        with pytest.raises(pkg.BsgsError):
            e.step(pyec.G)                                            # tables not uploaded
        with pytest.raises(pkg.BsgsError):
            e.upload_ht(np.zeros(10, dtype=np.uint8))                 # wrong size


# ------------------------------------------------------------------ device-side table builders
@pytest.mark.parametrize("T,B,p,w", [(32, 4, 16, 4096), (7, 3, 5, 1000), (64, 20, 100, 60000), (1, 1, 1, 5), (2, 1, 1, 9), (512, 8, 6, 1 << 30)])
def test_giants_generate_equals_oracle_g2(pkg, orc, T, B, p, w):
    with pkg.Engine(T, B, p, w, 10) as e:
        e.giants_generate()
        assert bytes(e.giants_export()) == bytes(orc.build_g2(T, B, p, w))


def test_giants_upload_export_roundtrip(pkg, orc):
    T, B, p, w = 24, 5, 7, 777
    g2 = orc.build_g2(T, B, p, w)
    with pkg.Engine(T, B, p, w, 8) as e:
        e.upload_giants(g2)
        assert bytes(e.giants_export()) == bytes(g2)


@pytest.mark.parametrize("w,htsz", [(50000, 14), (50000, 9), (3000, 12), (1, 1), (100000, 4)])
def test_ht_device_build_equals_oracle_images(pkg, orc, w, htsz):
    keys = synthetic_keys(w, [], seed=w * 31 + htsz)
    if w > 100:
        keys[17] = keys[5]; keys[99] = keys[5]                      # duplicate (bucket, tag): later insert first
    cpu, gpu = orc.build_ht(keys, htsz)
    with pkg.Engine(4, 4, 4, w, htsz) as e:
        e.ht_build_from_keys(keys)
        assert bytes(e.ht_export("gpu")) == bytes(gpu)
        assert bytes(e.ht_export("cpu")) == bytes(cpu)
        assert e.probe_keys(keys[:2000]).all()


def test_ht_true_baby_keys_device_build(pkg, orc):
    w, htsz = 1 << 16, 13
    keys = orc.baby_keys(w)
    cpu, gpu = orc.build_ht(keys, htsz)
    with pkg.Engine(4, 4, 4, w, htsz) as e:
        e.ht_build_from_keys(keys)
        assert bytes(e.ht_export("cpu")) == bytes(cpu)


@pytest.mark.parametrize("T,B,p,w,htsz", [(32, 4, 8, 1 << 16, 13), (16, 2, 3, 5000, 9), (64, 20, 100, 100000, 12), (512, 8, 6, 3000, 8),
                                          (8, 1, 1, 1, 1), (4, 1, 1, 9, 2), (128, 16, 8, 1 << 20, 18)])
def test_ht_generate_equals_oracle_baby_table(pkg, orc, T, B, p, w, htsz):
    """Baby-step table generated by the giant-step kernel in emit mode == the reference's table
    ((j+1)G, low 8 bytes of X, 1_9_7File.pb:1127/1221) as restated by the oracle, byte for byte."""
    keys = orc.baby_keys(w)
    cpu, gpu = orc.build_ht(keys, htsz)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.ht_generate()
        assert bytes(e.ht_export("gpu")) == bytes(gpu)
        assert bytes(e.ht_export("cpu")) == bytes(cpu)


def test_solve_with_device_generated_tables(pkg):
    """Whole problem with nothing from the oracle: device-generated giants + baby table, key recovered."""
    T, B, p, w, htsz = 64, 8, 16, 1 << 18, 16
    pk = 0x8000000000
    secret = 0xE9AE4933D6                                           # puzzle #40 (BASELINE.json configs[0])
    target = pkg.ec.parse_pubkey("03a2efa402fd5268400c77c20e574ba86409ededee7c4020e4b9f0edbee53de0d4")
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.giants_generate(); e.ht_generate()
        htcpu = e.ht_export("cpu")
        key, done, seen = pkg.solve(e, htcpu, T, B, p, w, htsz, pk, 0xffffffffff, target)
        assert key == secret


# ------------------------------------------------------------------ BASELINE.json full sizes
@pytest.mark.slow
def test_full_size_config2_tables_and_puzzle65(pkg, orc):
    """configs[1] at full size (-t 512 -b 148 -p 256 -w 30 -htsz 28), through size-independent
    properties: the reference's own run-time validators restated (checkGiantArr 1_9_7File.pb:1524,
    checkHTpackFile :3101, sortedness :2959) on the device-built tables, a launch compared with the
    oracle on a slice of the giants, and the published puzzle #65 key recovered."""
    T, B, p, w, htsz = 512, 148, 256, 1 << 30, 28
    N, H = T * B * p, 1 << 28
    rnd = random.Random(65)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.giants_generate(); e.ht_generate()
        g2 = e.giants_export()
        assert g2.nbytes == 1241513984
        D = pyec.neg(pyec.mul(pyec.G, 2 * w))
        for k in [0, 1, N - 1, N // 2] + [rnd.randrange(N) for _ in range(60)]:      # checkGiantArr: arr[k] == (k+1)*ADDPUBG
            assert orc.g2_get(g2, T, B, p, k) == pyec.mul(D, k + 1)
        gpu = e.ht_export("gpu")
        assert gpu.nbytes == 5368709124
        off = gpu[:4 * (H + 1)].view(np.uint32)
        assert off[0] == 0 and off[H] == w and bool(np.all(off[1:] >= off[:-1]))
        tags = gpu[4 * (H + 1):].view(np.uint32)
        for h in [rnd.randrange(H) for _ in range(20000)]:                           # ascending tags inside buckets
            seg = tags[int(off[h]):int(off[h + 1])]
            assert np.all(seg[1:] >= seg[:-1])
        js = [0, 1, w - 1] + [rnd.randrange(w) for _ in range(200)]                  # checkHTpackFile: (j+1)G is found
        keys = np.array([pyec.mul(pyec.G, j + 1)[0] & M64 for j in js], dtype=np.uint64)
        assert e.probe_keys(keys).all()
        assert all(orc.probe_htgpu(gpu, htsz, int(k)) for k in keys[:50])
        assert not e.probe_keys(keys ^ np.uint64(1 << 45)).any()
        # one launch == the oracle on two slices of the giants (first rows and the ragged tail)
        target = pkg.ec.parse_pubkey("0230210c23b1a047bc9bdbb13448e67deddc108946de6de639bcc75d47c0216b1b")
        s = pkg.Search(T, B, p, w, htsz, 1 << 64, target)
        P0 = s.window_point(3)
        got = norm_hits(e.step(P0))
        for lo, hi in ((0, 300000), (N - 200000, N)):
            exp = [h for h in orc.launch(g2, T, B, p, gpu, htsz, P0, nthreads=16, k_begin=lo, k_end=hi) if h[0] != 5]
            assert [h for h in got if h[0] != 5 and lo <= h[1] < hi] == exp
        del gpu, tags, off
        htcpu = e.ht_export("cpu")
        assert htcpu.nbytes == 9663676420
        key, done, seen = pkg.solve(e, htcpu, T, B, p, w, htsz, 1 << 64, (1 << 65) - 1, target)
        assert key == 0x1a838b13505b26867


# ------------------------------------------------------------------ command line (reference flag surface)
def test_cli_config0_puzzle40(pkg, tmp_path, monkeypatch):
    """BASELINE.json configs[0]: puzzle #40, -pk 0x8000000000 -pke 0xffffffffff -w 22 -htsz 25."""
    import io
    monkeypatch.chdir(tmp_path)
    out = io.StringIO()
    argv = "-t 32 -b 4 -p 16 -w 22 -htsz 25 -pk 0x8000000000 -pke 0xffffffffff -pb 03a2efa402fd5268400c77c20e574ba86409ededee7c4020e4b9f0edbee53de0d4".split()
    assert pkg.cli.main(argv, out=out) == 0
    text = out.getvalue()
    assert "KEY[1]: 0x" + format(0xE9AE4933D6, "064x") in text
    assert open("win.txt").read().startswith("KEY[1]: 0x" + format(0xE9AE4933D6, "064x"))
    names = sorted(os.listdir("."))
    assert pkg.formats.htgpu_name(4194304, 25) in names and pkg.formats.htcpu_name(4194304, 25) in names and "32_4_16_4194304_g2.BIN" in names
    assert os.path.getsize(pkg.formats.htcpu_name(4194304, 25)) == 167772164 and os.path.getsize(pkg.formats.htgpu_name(4194304, 25)) == 150994948
    # second run: tables are loaded from the files the first run wrote (reference behaviour, Save_HTpacked 1_9_7File.pb:3645-3759)
    out2 = io.StringIO()
    assert pkg.cli.main(argv, out=out2) == 0 and "Load hashtable" in out2.getvalue() and "KEY[1]" in out2.getvalue()


def test_cli_infile_sequential_targets(pkg, tmp_path, monkeypatch):
    """configs[3] in miniature: -infile with several pubkeys over one range, tables loaded once."""
    import io
    monkeypatch.chdir(tmp_path)
    rnd = random.Random(20261017)
    pk = 0x49dccfd96dc5df56487436f5a1b18c4f5d34f65ddb48cb5e0000000000000000 >> 200 << 36      # a 2^36-wide range
    secrets = [pk + rnd.randrange(1, 1 << 36) for _ in range(4)]
    with open("pubs.txt", "w") as f:
        for i, s in enumerate(secrets):
            pt = pyec.mul(pyec.G, s)
            f.write((pkg.ec.compress(pt) if i % 2 else format(pt[0], "064x") + format(pt[1], "064x")) + "\n")
    out = io.StringIO()
    argv = f"-t 32 -b 4 -p 16 -w 18 -htsz 16 -pk {pk:x} -pke {pk + (1 << 36):x} -infile pubs.txt".split()
    assert pkg.cli.main(argv, out=out) == 0
    text = out.getvalue()
    for i, s in enumerate(secrets, start=1):
        assert f"KEY[{i}]: 0x{s:064x}" in text
    assert text.count("Generate") == 2 and "found 4 of 4" in text                  # tables built once (baby + giants)


@pytest.mark.slow
def test_full_size_config5_max_table(pkg):
    """configs[4] (-w 31.5 -htsz 29): 3,037,000,499 entries, CSR offsets beyond 2^31.  Members are found
    (incl. the last entries, whose offsets are the largest), non-members are not, image size as SURVEY 8a."""
    w, htsz = pkg.ec.parse_w("31.5"), 29
    assert w == 3037000499
    rnd = random.Random(5)
    with pkg.Engine(512, 148, 256, w, htsz) as e:
        e.ht_generate()
        ptr, nbytes = e.ht_device_image()
        assert nbytes == 14295485648
        js = [0, 1, w - 1, w - 2, (1 << 31) - 1, 1 << 31, (1 << 31) + 1] + [rnd.randrange(w) for _ in range(120)]
        keys = np.array([pyec.mul(pyec.G, j + 1)[0] & M64 for j in js], dtype=np.uint64)
        assert e.probe_keys(keys).all()
        assert not e.probe_keys(keys ^ np.uint64(1 << 50)).any()
        ms, found = e.probe_bench(1 << 26, seed=7, iters=1)
        assert found <= 64                                          # expected 2^26 * 5.66 * 2^-61 * ... ~ 0.2


# ------------------------------------------------------------------ boundary behaviour
def test_record_cap_overflow_is_reported_not_overrun(pkg, orc):
    """More hits than `cap`: status BSGS_E_OVERFLOW, true count returned, first `cap` records stored
    (the reference overruns into the giant array at record 240, 1_9_7File.pb:2337)."""
    import ctypes as C
    T, B, p, w, htsz = 32, 4, 8, 4000, 4
    N = T * B * p
    rnd = random.Random(3)
    Ppt = random_point(rnd)
    xs = [candidate_x(Ppt, w, k, 1 + (k & 1)) for k in range(0, 40)]
    cpu, gpu = orc.build_ht(synthetic_keys(w, xs, 9), htsz)
    g2 = orc.build_g2(T, B, p, w)
    exp = orc.launch(g2, T, B, p, gpu, htsz, Ppt)
    assert len(exp) >= 40
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_ht(gpu)
        hits = (pkg.native.Hit * 8)(); n = C.c_uint32()
        rc = e.L.bsgs_step(e.h, Ppt[0].to_bytes(32, "little"), Ppt[1].to_bytes(32, "little"), hits, 8, C.byref(n))
        assert rc == -6 and n.value == len(exp)
        assert all((hits[i].code, hits[i].idx) in set(exp) for i in range(8))
        assert norm_hits(e.step(Ppt)) == norm_hits(exp)            # the context stays usable


def test_two_contexts_driven_from_two_threads(pkg, orc):
    """One context per host thread, like CreateThread(@cuda(), i) (1_9_7File.pb:4776)."""
    import threading
    T, B, p, w, htsz = 32, 4, 8, 3000, 9
    g2 = orc.build_g2(T, B, p, w)
    rnd = random.Random(12)
    pts = [random_point(rnd) for _ in range(6)]
    xs = [candidate_x(pt, w, (i * 91) % (T * B * p), 1 + (i & 1)) for i, pt in enumerate(pts)]
    cpu, gpu = orc.build_ht(synthetic_keys(w, xs, 4), htsz)
    exp = [norm_hits(orc.launch(g2, T, B, p, gpu, htsz, pt)) for pt in pts]
    res = {}

    def work(tid):
        with pkg.Engine(T, B, p, w, htsz) as e:
            e.upload_giants(g2); e.upload_ht(gpu)
            res[tid] = [norm_hits(e.step(pt)) for pt in pts for _ in range(3)][::3]

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    assert res[0] == exp and res[1] == exp


def test_reupload_and_argument_errors(pkg, orc):
    T, B, p, w, htsz = 16, 2, 4, 1000, 7
    g2 = orc.build_g2(T, B, p, w)
    rnd = random.Random(1)
    Ppt = random_point(rnd)
    k1, k2 = synthetic_keys(w, [candidate_x(Ppt, w, 3, 1)], 1), synthetic_keys(w, [candidate_x(Ppt, w, 5, 2)], 2)
    _, gpu1 = orc.build_ht(k1, htsz); _, gpu2 = orc.build_ht(k2, htsz)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2)
        e.upload_ht(gpu1); a = norm_hits(e.step(Ppt))
        e.upload_ht(gpu2); b2 = norm_hits(e.step(Ppt))             # table replaced in place
        assert (1, 3) in a and (2, 5) in b2 and a == norm_hits(orc.launch(g2, T, B, p, gpu1, htsz, Ppt)) and b2 == norm_hits(orc.launch(g2, T, B, p, gpu2, htsz, Ppt))
        with pytest.raises(pkg.BsgsError):
            e.step_multi([Ppt] * 17)                               # more than BSGS_MAX_WINDOWS
        with pytest.raises(pkg.BsgsError):
            e.ht_export("cpu")                                     # payloads unknown after a plain htGPU upload
        with pytest.raises(pkg.BsgsError):
            e.upload_giants(g2[:-64])


# ------------------------------------------------------------------ device-side verification
def test_verify_drops_false_positives_and_keeps_true_hits(pkg, orc):
    """Real baby table + real giants: with verify=1 the records are exactly those of the raw launch whose
    candidate X equals a baby point on all 256 bits (what checkerThread establishes on the CPU), each with
    its payload j; planted 60-bit collisions (same bucket + tag, different point) are dropped."""
    T, B, p, w, htsz = 32, 4, 8, 1 << 12, 6                          # load 64 per bucket: many tag candidates per probe
    N = T * B * p
    keys = orc.baby_keys(w)
    rnd = random.Random(77)
    # true hits: P = (m + (k+1)*2w)G  ->  P + Q_k = m*G with m in [1, w]  (code 1), and the mirrored code 2
    k1, m1 = 17, 1234
    k2, m2 = N - 1, 1
    P1 = pyec.mul(pyec.G, m1 + (k1 + 1) * 2 * w)
    P2 = pyec.mul(pyec.G, (pyec.N - m2 - (k2 + 1) * 2 * w) % pyec.N)       # P - Q_k = -(m2)G: X equals X(m2 G)
    # false positives: overwrite some table keys with the 64-bit key of random candidates (same bucket+tag, other point)
    Pf = random_point(rnd)
    fake = [(1, 5), (2, 99), (1, 700)]
    keys2 = keys.copy()
    for i, (c, k) in enumerate(fake):
        keys2[3000 + i] = np.uint64(candidate_x(Pf, w, k, c) & M64)
    cpu, gpu = orc.build_ht(keys2, htsz)
    g2 = orc.build_g2(T, B, p, w)
    with pkg.Engine(T, B, p, w, htsz) as e:
        e.upload_giants(g2); e.upload_htcpu(cpu)
        assert bytes(e.ht_export("cpu")) == bytes(cpu) and bytes(e.ht_export("gpu")) == bytes(gpu)
        for Ppt, want in ((P1, (1, k1, m1 - 1)), (P2, (2, k2, m2 - 1)), (Pf, None)):
            e.set_verify(False)
            raw = norm_hits(e.step(Ppt))
            assert raw == norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt))
            e.set_verify(True)
            ver = e.step(Ppt); pay = list(e.payloads)
            assert set(ver) <= set(raw)
            if want:
                assert (want[0], want[1]) in ver and pay[ver.index((want[0], want[1]))] == want[2]
            else:
                assert set(fake) <= set(raw) and not (set(fake) & set(ver))
            for (c, i), pl in zip(ver, pay):                          # every survivor is a genuine 256-bit match
                if c in (1, 2):
                    assert candidate_x(Ppt, w, i, c) == pyec.mul(pyec.G, pl + 1)[0]


def test_verify_codes_4_5_and_solve(pkg, orc):
    T, B, p, w, htsz = 16, 4, 8, 1 << 14, 10
    with pkg.Engine(T, B, p, w, htsz, verify=1) as e:
        e.giants_generate(); e.ht_generate()
        # code 5: P itself is a baby point; code 4: P == Q_k (doubling) -- handled without a table entry
        Pb = pyec.mul(pyec.G, 4321)
        hits = e.step(Pb)
        assert (5, 0) in hits and e.payloads[hits.index((5, 0))] == 4320
        k = 9
        Q = giant(w, k)
        hits = e.step(Q)
        assert all(c != 4 or pl == 0xFFFFFFFE for (c, i), pl in zip(hits, e.payloads))
        # whole problems through solve(): only verified records reach the host-side checker
        rnd = random.Random(31)
        pk = 1 << 36
        for _ in range(4):
            secret = pk + rnd.randrange(1, 1 << 35)
            target = pyec.mul(pyec.G, secret)
            key, done, seen = pkg.solve(e, None, T, B, p, w, htsz, pk, pk + (1 << 35), target)
            assert key == secret and seen <= 2


def test_device_listing_and_suggestion(pkg):
    """bsgs_device_query + the Tune mirror: what the reference prints at start-up (1_9_7File.pb:778-860)."""
    from bsgs_cuda_b200 import native, tune, cli
    devs = native.devices()
    assert devs and devs[0]["sm_count"] > 0 and 0 < devs[0]["free_bytes"] <= devs[0]["total_bytes"]
    lines = []
    cli.list_devices(lines.append)
    assert lines[0].startswith("Cuda device:") and lines[1].startswith("Device have: MP:")
    assert any(l.startswith("Try -t ") and "(this build)" in l for l in lines)
    s = tune.tune_b200(devs[0]["free_bytes"], devs[0]["sm_count"])
    assert s["device_mb"] * 1048576 < devs[0]["free_bytes"]


def test_bcast_ht_between_contexts(pkg, orc):
    """bsgs_bcast_ht: the table of context 0 (a device-side build, with payloads) copied device to
    device into two more contexts; all three then emit the same records, verified ones included."""
    T, B, p, w, htsz = 32, 4, 8, 1 << 14, 12
    rnd = random.Random(77)
    with pkg.Engine(T, B, p, w, htsz) as e0, pkg.Engine(T, B, p, w, htsz) as e1, pkg.Engine(T, B, p, w, htsz) as e2:
        for e in (e0, e1, e2):
            e.giants_generate()
        e0.ht_generate()
        pkg.native.bcast_ht([e0, e1, e2])
        assert bytes(e1.ht_export("gpu")) == bytes(e0.ht_export("gpu")) == bytes(e2.ht_export("gpu"))
        assert bytes(e2.ht_export("cpu")) == bytes(e0.ht_export("cpu"))
        secret = rnd.randrange(1, 1 << 24)                               # inside the span of one launch: a real hit exists
        base = pyec.mul(pyec.G, secret + p * w + 1)                   # some base point whose neighbourhood holds baby points
        pts = [base, random_point(rnd)]
        for e in (e0, e1, e2):
            e.set_verify(True)
        ref = [norm_hits(e0.step(pt)) for pt in pts]
        assert ref[0]
        for e in (e1, e2):
            assert [norm_hits(e.step(pt)) for pt in pts] == ref
        with pytest.raises(pkg.BsgsError):                             # different table geometry is refused
            with pkg.Engine(T, B, p, w, htsz + 1) as bad:
                pkg.native.bcast_ht([e0, bad])
