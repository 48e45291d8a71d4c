"""Command-line surface (getprogparam 1_9_7File.pb:875-1042) -- parsing and validation on CPU; the
end-to-end runs (configs[0] puzzle #40 and the -infile path) are GPU tests."""
import os
import pytest
import bsgs_cuda_b200 as b
from bsgs_cuda_b200 import cli
from tests import pyec


def test_defaults_and_flags():
    o = cli.parse_args([])
    assert (o["t"], o["b"], o["p"], o["w"], o["htsz"], o["wt"]) == (256, 132, 400, 1 << 25, 25, 180)      # 1_9_7File.pb:181-186, 168, 152
    o = cli.parse_args("-t 512 -b 148 -p 256 -w 30 -htsz 28 -pk 0x10000000000000000 -pke 0x1FFFFFFFFFFFFFFFF -d 0,1 -wt 5 -pb 0x02AB".split())
    assert (o["t"], o["b"], o["p"], o["w"], o["htsz"]) == (512, 148, 256, 1 << 30, 28)
    assert o["pk"] == 1 << 64 and o["pke"] == (1 << 65) - 1 and o["d"] == "0,1" and o["wt"] == 30 and o["pb"] == "02ab"
    assert cli.parse_args(["-w", "30.5"])["w"] == 1518500249 and cli.parse_args(["-w", "4194304"])["w"] == 4194304
    with pytest.raises(cli.UsageError):
        cli.parse_args(["-zz"])


def test_validation_limits():
    ok = cli.parse_args("-w 22 -htsz 25 -p 16".split())
    assert cli.validate(ok) == []
    for bad in ("-p 7", "-htsz 32", "-w 3069485951", "-pk 0", "-pk 0x10 -pke 0x10"):
        with pytest.raises(cli.UsageError):
            cli.validate(cli.parse_args(bad.split()))
    assert any("UNSAFE" in m for m in cli.validate(cli.parse_args("-htsz 28 -w 1777178604".split())))     # README:19-24


def test_fingerprint_and_checkpoint_format(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    o = cli.parse_args("-t 32 -b 4 -p 16 -w 22 -htsz 25 -pk 0x8000000000 -pke 0xFFFFFFFFFF".split())
    fp = cli.fingerprint(o)
    import hashlib
    assert fp == hashlib.sha1(b"32416" + b"4194304" + b"8000000000" + b"ffffffffff" + b"25").hexdigest() and len(fp) == 40
    cli.save_checkpoint(3, "ab" * 64, 0x1234, fp)
    lines = open("currentwork.txt").read().split("\n")
    assert lines[0] == "3" and lines[1] == "ab" * 64 and lines[2] == format(0x1234, "064x") and lines[3] == fp   # 1_9_7File.pb:3913-3921
    assert not os.path.exists("currentwork.temp")


def test_resume_window_count():
    target = pyec.mul(pyec.G, (1 << 40) + 99)
    s = b.Search(8, 2, 4, 500, 7, 1 << 40, target)
    n = s.windows_needed((1 << 41) - 1)
    assert n == ((1 << 40) - 1 - 1) // s.stride + 2                  # every key_j <= width, plus the one launched before the test fires
    r = b.Search(8, 2, 4, 500, 7, 1 << 40, target, start_key=s.window_key(10))
    assert r.windows_needed((1 << 41) - 1) == n - 10 and r.window_point(0) == s.window_point(10)
