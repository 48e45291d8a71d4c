"""Host-side mirror of the reference interface (bsgs-cuda_b200/{ec,formats,search}.py) against the
oracle's restatement of the same reference code (GetJob 1_9_7File.pb:2077-2092, set-up 4689-4765 /
5042-5064, checkerThread 3933-4296, -w parsing 1009-1022, file names 3652-3655 / 1916)."""
import random
import numpy as np
import pytest
import bsgs_cuda_b200 as b
from bsgs_cuda_b200 import ec, formats
from tests import pyec


def test_parse_w_matches_reference_rule(orc):
    for arg, want in (("22", 4194304), ("30", 1073741824), ("30.5", 1518500249), ("31.5", 3037000499), ("29.87", None), ("1000000", 1000000), ("33", 33)):
        assert ec.parse_w(arg) == orc.parse_w(arg)
        if want:
            assert ec.parse_w(arg) == want


def test_pubkey_forms(orc):
    k = 0x1E9AD
    pt = pyec.mul(pyec.G, k)
    forms = [format(pt[0], "064x") + format(pt[1], "064x"), "04" + format(pt[0], "064x") + format(pt[1], "064x"), ec.compress(pt)]
    for f in forms:
        assert ec.parse_pubkey(f) == pt == orc.parse_pubkey(f)
    assert ec.parse_pubkey("03a2efa402fd5268400c77c20e574ba86409ededee7c4020e4b9f0edbee53de0d4") == pyec.mul(pyec.G, 0xE9AE4933D6)
    with pytest.raises(ValueError):
        ec.parse_pubkey("05" + "00" * 32)


def test_file_names_and_sizes():
    assert formats.htgpu_name(4194304, 25) == "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798_4194304_33554432_htGPUv0.BIN"
    assert formats.htcpu_name(1073741824, 28).endswith("_1073741824_268435456_htCPUv0.BIN")
    assert formats.g2_name(512, 148, 256, 1073741824) == "512_148_256_1073741824_g2.BIN"
    assert formats.htgpu_bytes(1 << 30, 28) == 5368709124 and formats.htcpu_bytes(1 << 30, 28) == 9663676420
    assert formats.htgpu_bytes(3037000499, 29) == 14295485648 and formats.g2_bytes(512, 148, 256) == 1241513984


def test_window_enumeration_equals_getjob(orc):
    T, B, p, w, htsz = 8, 3, 6, 777, 9
    pk = 0x8000000000
    target = pyec.mul(pyec.G, pk + 123456789)
    s = b.Search(T, B, p, w, htsz, pk, target)
    o = orc.SearchState(T, B, p, w, htsz, pk, target)
    for j in range(6):
        key, P = o.get_job()
        assert key == s.window_key(j) and P == s.window_point(j)
    # multi-GPU striding: rank r of R sees windows r, r+R, ... of the same sequence
    R = 3
    ranks = [b.Search(T, B, p, w, htsz, pk, target, rank=r, world=R) for r in range(R)]
    for j in range(6):
        assert ranks[j % R].window_key(j // R) == s.window_key(j)
        assert ranks[j % R].window_point(j // R) == s.window_point(j)
    assert sum(r.windows_needed(pk * 2 - 1) for r in ranks) == s.windows_needed(pk * 2 - 1)


@pytest.mark.parametrize("seed", [1, 2])
def test_check_hit_equals_oracle_checker(orc, seed):
    """Plant a key so that the hit lands on code 1 / 2 and compare both checkers on real tables."""
    T, B, p, w, htsz = 4, 2, 4, 64, 5
    N = T * B * p
    keys = orc.baby_keys(w)
    cpu, gpu = orc.build_ht(keys, htsz)
    g2 = orc.build_g2(T, B, p, w)
    rnd = random.Random(seed)
    pk = rnd.randrange(1 << 40, 1 << 41)
    found = 0
    for _ in range(12):
        secret = pk + rnd.randrange(1, 4 * N * w * 3)
        target = pyec.mul(pyec.G, secret)
        s = b.Search(T, B, p, w, htsz, pk, target)
        o = orc.SearchState(T, B, p, w, htsz, pk, target)
        for j in range(4):
            wkey, P = o.get_job()
            if P is None or P[0] == 0:
                continue
            for code, idx in orc.launch(g2, T, B, p, gpu, htsz, P):
                mine = s.check_hit(cpu, code, idx, j)
                theirs = o.check_hit(cpu, code, idx, wkey, P)
                assert mine == theirs
                if mine is not None:
                    assert mine == secret
                    found += 1
    assert found >= 8


def _ref_loop_launch_count(orc, T, B, p, w, htsz, pk, pke, target):
    """cuda()'s loop as written (1_9_7File.pb:2402-2523): launch window j, THEN break if key_j > pke - pk,
    else GetJob."""
    o = orc.SearchState(T, B, p, w, htsz, pk, target)
    width = pke - pk
    launches = 0
    while True:
        key, _ = o.get_job()
        launches += 1
        if key > width:
            return launches


@pytest.mark.parametrize("T,B,p,w,pk,width", [(2, 2, 4, 8, 1, 500), (2, 2, 4, 8, 1, 511), (2, 2, 4, 8, 1, 512), (2, 2, 4, 8, 1, 513), (2, 2, 4, 8, 77, 0),
                                              (4, 3, 2, 100, 1 << 40, 123457), (8, 3, 6, 777, 0x8000000000, 0x7fffffffff)])
def test_windows_needed_equals_the_reference_loop_and_covers_the_range(orc, T, B, p, w, pk, width):
    """The end-of-range test runs after the launch, so the window whose key first exceeds the width is
    still launched -- it holds the tail of [pk, pke] (with T=2,B=2,p=4,w=8,pk=1,pke=501 the offset
    400 lies only in window 1)."""
    N = T * B * p
    target = pyec.mul(pyec.G, pk + 5)
    s = b.Search(T, B, p, w, 5, pk, target)
    n = s.windows_needed(pk + width)
    if width < (1 << 24):
        assert n == _ref_loop_launch_count(orc, T, B, p, w, 5, pk, pk + width, target)
    else:
        assert n == (width - 1) // s.stride + 2
    # coverage: window j reaches key_j + p*w +- (2Nw + w); consecutive windows overlap; the union holds [1, width]
    lo = s.window_key(0) + s.centerbig - (2 * N * w + w)
    hi = s.window_key(n - 1) + s.centerbig + (2 * N * w + w)
    assert lo <= 1 and hi >= width
    for j in range(min(n - 1, 50)):
        assert s.window_key(j + 1) + s.centerbig - (2 * N * w + w) <= s.window_key(j) + s.centerbig + (2 * N * w + w) + 1
    # ranks partition exactly that count
    for R in (2, 3):
        assert sum(b.Search(T, B, p, w, 5, pk, target, rank=r, world=R).windows_needed(pk + width) for r in range(R)) == n


def test_key_in_the_last_half_stride_is_found_by_the_cpu_model(orc):
    """Keys planted near pke are recovered when exactly windows_needed() windows are searched."""
    T, B, p, w, htsz = 2, 2, 4, 8, 4
    N = T * B * p
    keys = orc.baby_keys(w)
    cpu, gpu = orc.build_ht(keys, htsz)
    g2 = orc.build_g2(T, B, p, w)
    pk = 1
    for pke, secret_off in ((501, 400), (501, 499), (1025, 1024), (700, 650)):
        target = pyec.mul(pyec.G, pk + secret_off)
        s = b.Search(T, B, p, w, htsz, pk, target)
        found = None
        for j in range(s.windows_needed(pke)):
            P = s.window_point(j)
            if P is None:
                continue
            for code, idx in orc.launch(g2, T, B, p, gpu, htsz, P):
                found = found or s.check_hit(cpu, code, idx, j)
        assert found == pk + secret_off, (pke, secret_off, found)
