"""Oracle tables / launch / checker vs. an independent big-int model and the reference's own
run-time validators restated (checkGiantArr 1_9_7File.pb:1524, checkHTpackFile :3101, sortedness :2959)."""
import random
import numpy as np
import pytest
from tests import pyec

P = pyec.P


def test_baby_keys_and_ht_format(orc):
    w, htsz = 5000, 10
    keys = orc.baby_keys(w, nthreads=3)
    # entry j is (j+1)*G, key = X mod 2^64
    pt = None
    for j in range(0, 40):
        pt = pyec.add(pt, pyec.G)
        assert int(keys[j]) == pt[0] & (2**64 - 1)
    for j in (4095, 4096, 4097, 4999):                    # across the #helpsize batch boundary
        assert int(keys[j]) == pyec.mul(pyec.G, j + 1)[0] & (2**64 - 1)
    cpu, gpu = orc.build_ht(keys, htsz)
    H = 1 << htsz
    assert cpu.nbytes == 4 * H + 4 + 8 * w and gpu.nbytes == 4 * H + 4 + 4 * w
    off = cpu[:4 * (H + 1)].view(np.uint32)
    assert off[0] == 0 and off[H] == w and np.all(np.diff(off.astype(np.int64)) >= 0)
    assert np.array_equal(off, gpu[:4 * (H + 1)].view(np.uint32))
    items = cpu[4 * (H + 1):].view(np.uint32).reshape(-1, 2)
    tags = gpu[4 * (H + 1):].view(np.uint32)
    assert np.array_equal(items[:, 0], tags)
    for h in range(H):                                    # ascending tags per bucket, right bucket, right payload
        lo, hi = int(off[h]), int(off[h + 1])
        t = items[lo:hi, 0].astype(np.int64)
        assert np.all(np.diff(t) >= 0)
        for tag, j in items[lo:hi]:
            k = int(keys[j])
            assert (k & (H - 1)) == h and (k >> 32) == int(tag)
    # checkHTpackFile restated: random (k+1)*G found through the packed format
    rnd = random.Random(5)
    for _ in range(64):
        j = rnd.randrange(w)
        assert orc.probe_htgpu(gpu, htsz, int(keys[j]))
        assert j in orc.lookup_htcpu(cpu, htsz, int(keys[j]))
    for _ in range(64):
        k = rnd.getrandbits(64)
        present = any(int(x) == k for x in ())            # random 64-bit keys are absent w.h.p.
        assert orc.probe_htgpu(gpu, htsz, k) == bool(orc.lookup_htcpu(cpu, htsz, k))


def test_g2_layout_and_values(orc):
    T, B, p, w = 4, 3, 6, 100
    N = T * B * p
    g2 = orc.build_g2(T, B, p, w, nthreads=2)
    assert g2.nbytes == N * 64
    A = pyec.neg(pyec.mul(pyec.G, 2 * w))
    X = g2[:N * 32].view(np.uint32); Y = g2[N * 32:].view(np.uint32)
    TB = T * B
    q = None
    for k in range(N):                                    # checkGiantArr restated, exhaustively
        q = pyec.add(q, A)
        assert orc.g2_get(g2, T, B, p, k) == q
        blk, thr, r = k // (T * p), (k % (T * p)) // p, k % p
        g = blk * T + thr
        for i in range(8):                                # word 0 most significant
            assert int(X[r * 8 * TB + g + i * TB]) == (q[0] >> (32 * (7 - i))) & 0xFFFFFFFF
            assert int(Y[r * 8 * TB + g + i * TB]) == (q[1] >> (32 * (7 - i))) & 0xFFFFFFFF


def model_launch(Pxy, T, B, p, w, table):
    """Independent model of one launch on python ints. table: set of (bucket, tag)."""
    N = T * B * p
    A = pyec.neg(pyec.mul(pyec.G, 2 * w))
    out = []

    def probe(x):
        return ((x & 0xFFFFFFFF) & table["mask"], (x >> 32) & 0xFFFFFFFF) in table["set"]
    if probe(Pxy[0]):
        out.append((5, 0))
    q = None
    for k in range(N):
        q = pyec.add(q, A)
        if q[0] == Pxy[0]:
            d = pyec.add(Pxy, Pxy)
            if probe(d[0]):
                out.append((4, k))
            continue
        m = pyec.add(Pxy, pyec.neg(q)); s = pyec.add(Pxy, q)
        if probe(m[0]):
            out.append((2, k))
        if probe(s[0]):
            out.append((1, k))
    return sorted(out)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_launch_vs_model_and_key_recovery(orc, seed):
    """Toy search (w=64, N=8 like SURVEY A.7's check): every code path, oracle == model, key recovered."""
    T, B, p, w, htsz = 2, 2, 2, 64, 4
    N = T * B * p
    keys = orc.baby_keys(w, 1)
    cpu, gpu = orc.build_ht(keys, htsz)
    g2 = orc.build_g2(T, B, p, w, 1)
    H = 1 << htsz
    table = {"mask": H - 1, "set": {((int(k) & 0xFFFFFFFF) & (H - 1), int(k) >> 32) for k in keys}}
    rnd = random.Random(seed)
    pk = rnd.randrange(1, 2**40)
    codes_seen = set()
    for trial in range(60):
        kprime = rnd.randrange(1, 12 * N * w)             # spans a few windows
        if (kprime - 1 - p * w) % (2 * w) == 0:
            # P_j == +-Q_k exactly: the reference only probes X(2P) there (code 4, PTX197:L28996-),
            # so such keys are found only through a table false positive -- not asserted here
            kprime += 1
        secret = pk + kprime
        target = pyec.mul(pyec.G, secret)
        st = orc.SearchState(T, B, p, w, htsz, pk, target)
        found = None
        for j in range(8):
            wkey, Pj = st.get_job()
            assert wkey == 1 + j * 4 * N * w
            expP = pyec.add(pyec.add(target, pyec.neg(pyec.mul(pyec.G, pk))), pyec.neg(pyec.mul(pyec.G, wkey + p * w)))
            if expP is None:                              # P_j at infinity: unrepresentable
                break
            assert Pj == expP
            recs = orc.launch(g2, T, B, p, gpu, htsz, Pj, nthreads=2)
            assert recs == model_launch(Pj, T, B, p, w, table)
            assert recs == orc.launch(g2, T, B, p, gpu, htsz, Pj, batch=3, nthreads=1)   # batch size is not observable
            for code, idx in recs:
                got = st.check_hit(cpu, code, idx, wkey, Pj)
                if got is not None:
                    assert got == secret
                    found = got; codes_seen.add(code)
            if found:
                break
        if expP is not None:
            assert found == secret, (pk, kprime)
    assert {1, 2} <= codes_seen


def test_planted_codes_4_and_5(orc):
    T, B, p, w, htsz = 2, 2, 2, 64, 4
    N = T * B * p
    keys = orc.baby_keys(w, 1); cpu, gpu = orc.build_ht(keys, htsz); g2 = orc.build_g2(T, B, p, w, 1)
    pk = 12345
    # code 5: P_0 itself is a baby point: k' = 1 + p*w + (b+1)  
    for b in (0, 7, 63):
        secret = pk + 1 + p * w + (b + 1)
        st = orc.SearchState(T, B, p, w, htsz, pk, pyec.mul(pyec.G, secret))
        wkey, P0 = st.get_job()
        recs = orc.launch(g2, T, B, p, gpu, htsz, P0)
        assert (5, 0) in recs
        assert st.check_hit(cpu, 5, 0, wkey, P0) == secret
    # code 4: P_0 == -+Q_k  =>  k' = 1 + p*w -+ (k+1)*2w.  The kernel reports it only if X(2P) probes
    # true, so plant X(2P) as an extra table entry (w+1 items) to exercise the record and its decoding.
    hits4 = 0
    for k in range(N):
        for sign in (1, -1):
            kp = 1 + p * w + sign * (k + 1) * 2 * w
            secret = pk + kp
            if secret <= 0 or kp == 1:                    # kp == 1: T' - key_0*G is the point at infinity
                continue
            st = orc.SearchState(T, B, p, w, htsz, pk, pyec.mul(pyec.G, secret))
            wkey, P0 = st.get_job()
            qk = orc.g2_get(g2, T, B, p, k)
            assert P0[0] == qk[0]
            recs = orc.launch(g2, T, B, p, gpu, htsz, P0)
            dbl = pyec.add(P0, P0)
            assert ((4, k) in recs) == orc.probe_htgpu(gpu, htsz, dbl[0] & (2**64 - 1))
            assert not any(c in (1, 2) and i == k for c, i in recs)      # no +- probes on the equal path
            keys2 = np.concatenate([keys, np.array([dbl[0] & (2**64 - 1)], dtype=np.uint64)])
            cpu2, gpu2 = orc.build_ht(keys2, htsz)
            recs = orc.launch(g2, T, B, p, gpu2, htsz, P0)
            assert (4, k) in recs
            hits4 += 1
            assert st.check_hit(cpu2, 4, k, wkey, P0) == secret
    assert hits4 > 0


@pytest.mark.slow
def test_config1_plumbing_puzzle40_on_cpu(orc):
    """BASELINE.json configs[0] at its full table size, CPU only: -pk 0x8000000000 -pke 0xffffffffff -w 22 -htsz 25,
    tables built with the Curve64 restatement, windows enumerated by GetJob, records decoded by the
    checkerThread restatement -> key 0xE9AE4933D6 (-t 32 -b 4 -p 16, 16 windows)."""
    T, B, p, w, htsz = 32, 4, 16, 1 << 22, 25
    pk, pke, secret = 0x8000000000, 0xffffffffff, 0xE9AE4933D6
    keys = orc.baby_keys(w)
    cpu, gpu = orc.build_ht(keys, htsz)
    assert cpu.nbytes == 167772164 and gpu.nbytes == 150994948
    g2 = orc.build_g2(T, B, p, w)
    target = orc.parse_pubkey("03a2efa402fd5268400c77c20e574ba86409ededee7c4020e4b9f0edbee53de0d4")
    assert target == pyec.mul(pyec.G, secret)
    s = orc.SearchState(T, B, p, w, htsz, pk, target)
    found, windows = None, 0
    while found is None and windows < 20:
        wkey, P = s.get_job()
        windows += 1
        for code, idx in orc.launch(g2, T, B, p, gpu, htsz, P):
            k = s.check_hit(cpu, code, idx, wkey, P)
            if k is not None:
                found = k
        if wkey > pke - pk:
            break
    assert found == secret and windows <= 17
