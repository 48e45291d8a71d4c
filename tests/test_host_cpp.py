"""The C++ host program (bsgs-cuda_b200/bsgs_b200, csrc/host_main.cpp): the reference's main program
restated above the C ABI.  CPU part: command line, validation and the settings fingerprint agree with
the Python mirror (bsgs-cuda_b200/cli.py) and with the reference's rules; the end-to-end runs are GPU
tests at the bottom."""
import hashlib
import os
import random
import subprocess
import pytest
import bsgs_cuda_b200 as b
from tests import pyec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bsgs-cuda_b200", "bsgs_b200")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "bsgs-cuda_b200"), "bsgs_b200"])
    return EXE


def run(exe, args, cwd=None):
    r = subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd, timeout=600)
    return r.returncode, r.stdout


@pytest.mark.parametrize("line", [
    "-t 32 -b 4 -p 16 -w 22 -htsz 25 -pk 0x8000000000 -pke 0xFFFFFFFFFF",
    "-t 512 -b 148 -p 256 -w 30 -htsz 28 -pk 0x10000000000000000 -pke 0x1ffffffffffffffff",
    "-w 30.5 -htsz 29 -pk 80000000000000000000 -pke FFFFFFFFFFFFFFFFFFFF",
    "-w 31.5 -htsz 29", "-w 4194304", "",
])
def test_fingerprint_equals_python_mirror_and_reference_rule(exe, line):
    rc, out = run(exe, line.split() + ["-fingerprint"])
    o = b.cli.parse_args(line.split())
    assert rc == 0 and out.strip().splitlines()[-1] == b.cli.fingerprint(o)
    # the rule itself (1_9_7File.pb:4634-4636): SHA-1 of Str(T)+Str(B)+Str(p)+Str(w)+pk+pke+Str(htsz)
    s = f"{o['t']}{o['b']}{o['p']}{o['w']}{o['pk_str']}{o['pke_str']}{o['htsz']}"
    assert out.strip().splitlines()[-1] == hashlib.sha1(s.encode()).hexdigest()


@pytest.mark.parametrize("bad,msg", [("-zz", "Unknown parameter"), ("-p 7", "multiple of 2"), ("-htsz 32", "less or equal to 31"),
                                     ("-w 3069485951", "less than 3069485951"), ("-pk 0", "more than 0"), ("-pk 0x10 -pke 0x10", "more than range begin"),
                                     ("-infile /nonexistent/pubs.txt", "not found")])
def test_usage_errors_match_the_python_mirror(exe, bad, msg):
    rc, out = run(exe, bad.split() + ["-fingerprint"])
    assert rc == 2 and msg in out
    with pytest.raises(b.cli.UsageError):
        b.cli.validate(b.cli.parse_args(bad.split())); b.cli.read_pubs(b.cli.parse_args(bad.split()))


def test_unsafe_warning(exe):
    rc, out = run(exe, "-htsz 28 -w 1777178604 -fingerprint".split())
    assert rc == 0 and "UNSAFE" in out                                                  # README.md:19-24 of the reference


@pytest.mark.parametrize("cnt,running,w,job,total,rates", [(0x1a2b3c, 2, 1 << 30, 3725, 90061, [27000000000, 26500000000]),
                                                          (1, 1, 4194304, 0, 2, [0]), (0xfff00000000000001, 8, 3037000499, 59, 3600, [25 * 10**9] * 8)])
def test_status_line_is_the_reference_format_in_both_hosts(exe, cnt, running, w, job, total, rates):
    """1_9_7File.pb:5135: "Cnt:"+cnt$+" ["+Str(isruning)+"]"+"[ "+perf$+"] = "+Str(totalhash/#MB)+" MKeys/s x2^"+StrD(wald,1)+"=2^"
    +StrD(hashd+wald,2)+" Jt:"+hh:ii:ss+" Tt:"+hh:ii:ss, perf$ = every GPU's rate / 2^20 followed by a blank."""
    import math
    rc, out = run(exe, ["-statusline", ",".join([format(cnt, "x"), str(running), str(w), str(job), str(total)] + [str(r) for r in rates])])
    mine = b.cli.format_status(cnt, running, rates, w, job, total)
    assert rc == 0 and out.strip("\n") == mine
    tot = sum(rates)
    hms = lambda t: "%02d:%02d:%02d" % (t // 3600 % 24, t // 60 % 60, t % 60)       # noqa: E731
    want = ("Cnt:" + format(cnt, "x") + " [" + str(running) + "]" + "[ " + "".join(str(r >> 20) + " " for r in rates) + "] = " + str(tot >> 20)
            + " MKeys/s x2^" + "%.1f" % math.log2(2 * w) + "=2^" + "%.2f" % ((math.log2(tot) if tot else 0) + math.log2(2 * w)) + " Jt:" + hms(job) + " Tt:" + hms(total))
    assert mine == want


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cpp_host_config0_puzzle40_and_table_files(exe, tmp_path):
    """BASELINE.json configs[0] through the C++ host: key, win.txt, the reference's file names and
    sizes; the files it writes are byte-identical to the ones the Python mirror writes; a second run
    loads them instead of generating."""
    args = "-t 32 -b 4 -p 16 -w 22 -htsz 25 -pk 0x8000000000 -pke 0xffffffffff -pb 03a2efa402fd5268400c77c20e574ba86409ededee7c4020e4b9f0edbee53de0d4".split()
    d1 = tmp_path / "cpp"; d1.mkdir()
    rc, out = run(exe, args, cwd=d1)
    key = "KEY[1]: 0x" + format(0xE9AE4933D6, "064x")
    assert rc == 0 and key in out and "Generate 4194304 baby points" in out and "finished, found 1 of 1" in out
    assert open(d1 / "win.txt").read().startswith(key)
    names = {b.formats.htcpu_name(4194304, 25): 167772164, b.formats.htgpu_name(4194304, 25): 150994948, "32_4_16_4194304_g2.BIN": 32 * 4 * 16 * 64}
    for n, size in names.items():
        assert os.path.getsize(d1 / n) == size
    rc, out = run(exe, args, cwd=d1)
    assert rc == 0 and "Load hashtable" in out and "Load giants" in out and key in out
    # the Python mirror writes the same bytes
    import io
    d2 = tmp_path / "py"; d2.mkdir()
    cwd = os.getcwd(); os.chdir(d2)
    try:
        assert b.cli.main(args, out=io.StringIO()) == 0
    finally:
        os.chdir(cwd)
    for n in names:
        assert hashlib.sha256(open(d1 / n, "rb").read()).digest() == hashlib.sha256(open(d2 / n, "rb").read()).digest()


@pytest.mark.gpu
def test_cpp_host_infile_and_resume(exe, tmp_path):
    """-infile with several targets over a 256-bit range start (tables built once), then -wl resume
    from a checkpoint written in the reference's format (1_9_7File.pb:3913-3921)."""
    rnd = random.Random(20261017)
    pk = 0x49dccfd96dc5df56487436f5a1b18c4f5d34f65ddb48cb5e0000000000000000
    secrets = [pk + rnd.randrange(1, 1 << 36) for _ in range(4)]
    pts = [pyec.mul(pyec.G, s) for s in secrets]
    with open(tmp_path / "pubs.txt", "w") as f:
        for i, pt in enumerate(pts):
            f.write((b.ec.compress(pt) if i % 2 else format(pt[0], "064x") + format(pt[1], "064x")) + "\n")
    common = f"-t 32 -b 4 -p 16 -w 18 -htsz 16 -pk {pk:x} -pke {pk + (1 << 36):x} -infile pubs.txt -nosave"
    rc, out = run(exe, common.split(), cwd=tmp_path)
    assert rc == 0 and "found 4 of 4" in out and out.count("Generate") == 2
    for i, s in enumerate(secrets, start=1):
        assert f"KEY[{i}]: 0x{s:064x}" in out
    # resume: position 3, counter of a window before the one that holds key 3
    o = b.cli.parse_args(common.split())
    srch = b.Search(32, 4, 16, 1 << 18, 16, pk, pts[2])
    jwin = next(j for j in range(srch.windows_needed(pk + (1 << 36))) if abs((secrets[2] - pk) - (srch.window_key(j) + srch.centerbig)) <= srch.stride // 2 + (1 << 18))
    cnt = srch.window_key(max(jwin - 1, 0))
    b.cli.save_checkpoint(3, format(pts[2][0], "064x") + format(pts[2][1], "064x"), cnt, b.cli.fingerprint(o), path=str(tmp_path / "currentwork.txt"))
    rc, out = run(exe, (common + " -wl currentwork.txt").split(), cwd=tmp_path)
    assert rc == 0 and f"KEY[3]: 0x{secrets[2]:064x}" in out and "KEY[1]" not in out and "KEY[2]" not in out and f"KEY[4]: 0x{secrets[3]:064x}" in out
    # a checkpoint of another configuration is refused (1_9_7File.pb:4668-4672)
    rc, out = run(exe, (common.replace("-p 16", "-p 18") + " -wl currentwork.txt").split(), cwd=tmp_path)
    assert rc == 2 and "different" in out


@pytest.mark.gpu
def test_onlygen_writes_the_reference_files(exe, tmp_path):
    """-onlygen = the role of onlygen1_9_6File.pb: the three table files, then "finished ok"; both hosts
    write the same bytes, and the files pass the reference's own validators restated in the oracle."""
    import io
    args = "-t 32 -b 4 -p 16 -w 16 -htsz 12 -onlygen".split()
    d1 = tmp_path / "cpp"; d1.mkdir()
    rc, out = run(exe, args, cwd=d1)
    assert rc == 0 and out.strip().endswith("finished ok") and "KEY" not in out
    d2 = tmp_path / "py"; d2.mkdir()
    cwd = os.getcwd(); os.chdir(d2)
    try:
        buf = io.StringIO()
        assert b.cli.main(args, out=buf) == 0 and buf.getvalue().strip().endswith("finished ok")
    finally:
        os.chdir(cwd)
    w, htsz = 1 << 16, 12
    for n, size in ((b.formats.htcpu_name(w, htsz), b.formats.htcpu_bytes(w, htsz)), (b.formats.htgpu_name(w, htsz), b.formats.htgpu_bytes(w, htsz)), ("32_4_16_65536_g2.BIN", 32 * 4 * 16 * 64)):
        a1, a2 = open(d1 / n, "rb").read(), open(d2 / n, "rb").read()
        assert len(a1) == size and a1 == a2
    from oracle import oracle as orc
    orc.build(); orc.lib()
    import numpy as np
    assert bytes(orc.build_g2(32, 4, 16, w)) == open(d1 / "32_4_16_65536_g2.BIN", "rb").read()
    cpu, gpu = orc.build_ht(orc.baby_keys(w), htsz)
    assert bytes(cpu) == open(d1 / b.formats.htcpu_name(w, htsz), "rb").read() and bytes(gpu) == open(d1 / b.formats.htgpu_name(w, htsz), "rb").read()
    rc, out = run(exe, "-t 31 -b 4 -p 16 -onlygen -fingerprint".split())
    assert rc == 2 and "multiple of 2" in out


@pytest.mark.parametrize("mem,sm,major,minor", [(178_000_000_000, 148, 10, 0), (10_500_000_000, 68, 7, 5), (80_000_000_000, 132, 9, 0),
                                                (24_000_000_000, 82, 8, 6), (1_000_000_000, 6, 8, 6), (6_000_000_000, 20, 6, 1)])
def test_tune_lines_equal_the_python_mirror(exe, mem, sm, major, minor):
    """Both hosts print the same `Try ...` suggestions (Tune, 1_9_7File.pb:324-431 restated, and this build's own)."""
    from bsgs_cuda_b200 import tune
    rc, out = run(exe, ["-tune", f"{mem},{sm},{major},{minor}"])
    lines = out.strip().splitlines()
    cores = tune.cores_for(major, minor, sm)
    assert rc == 0 and lines[0] == f"Device have: MP:{sm} Cores+{cores}"
    want = []
    r = tune.tune_reference(mem, sm, cores)
    if r:
        want.append(tune.format_try(r).split(" Gen RAM")[0] + "   (reference Tune)")
    o = tune.tune_b200(mem, sm)
    if o:
        want.append(tune.format_try(o) + "   (this build)")
    assert lines[1:] == want


@pytest.mark.gpu
def test_both_hosts_find_a_key_in_the_last_half_stride(exe, tmp_path):
    """The reference tests the end of the range AFTER a launch (1_9_7File.pb:2512-2518): the window whose
    key first exceeds pke - pk is still searched and holds the tail of the range."""
    import io
    T, B, p, wl, htsz = 32, 4, 16, 12, 10
    w = 1 << wl
    stride = 4 * T * B * p * w
    pk = 0x4000000000
    for tail in (stride // 2 - 3, 5 * w):                 # distance of the key below pke
        width = 3 * stride + stride // 4                   # est = 3 -> 5 launches; the key lies beyond window 3's reach
        secret = pk + width - tail
        pt = pyec.mul(pyec.G, secret)
        args = f"-t {T} -b {B} -p {p} -w {wl} -htsz {htsz} -pk {pk:x} -pke {pk + width:x} -nosave -pb {b.ec.compress(pt)}".split()
        rc, out = run(exe, args, cwd=tmp_path)
        assert rc == 0 and f"KEY[1]: 0x{secret:064x}" in out, out
        buf = io.StringIO()
        cwd = os.getcwd(); os.chdir(tmp_path)
        try:
            assert b.cli.main(args, out=buf) == 0
        finally:
            os.chdir(cwd)
        assert f"0x{secret:064x}" in buf.getvalue(), buf.getvalue()
