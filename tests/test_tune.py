"""Parameter suggestion: the reference's Tune() restated (1_9_7File.pb:324-431) and this build's own."""
import bsgs_cuda_b200 as b
from bsgs_cuda_b200 import tune


def test_core_table_matches_reference_and_continues():
    # rows of 1_9_7File.pb:816-846
    assert tune.cores_for(2, 1, 10) == 480 and tune.cores_for(2, 0, 10) == 320
    assert tune.cores_for(3, 5, 13) == 13 * 192 and tune.cores_for(5, 2, 8) == 8 * 128
    assert tune.cores_for(6, 1, 20) == 20 * 64 and tune.cores_for(7, 5, 68) == 68 * 64
    assert tune.cores_for(8, 6, 82) == 82 * 128
    # continued: Hopper, Blackwell
    assert tune.cores_for(9, 0, 132) == 132 * 128 and tune.cores_for(10, 0, 148) == 148 * 128


def test_reference_tune_hand_evaluated_cases():
    # 2080 Ti-like: 68 SMs x 64 cores, 10.5e9 bytes free.  By hand: t starts at 128, b at 136;
    # k = 31 does not fit (8.59e9 + 2.15e9 + giants), k = 30 / htsz 27 does (4.83e9); t climbs to 512,
    # b to 272 (frmem 848 -> 565 -> 424); p = 424 (even).
    r = tune.tune_reference(10_500_000_000, 68, 68 * 64)
    assert (r["t"], r["b"], r["w"], r["htsz"]) == (512, 272, 30, 27)
    assert r["p"] in (422, 424) and r["p"] % 2 == 0
    gsize = r["t"] * r["b"] * r["p"] * 96 + 160 + 128
    assert gsize + (1 << 30) * 4 + (1 << 27) * 4 <= 10_500_000_000
    # tiny card: nothing fits above k = 27 / htsz 25
    r = tune.tune_reference(1_000_000_000, 6, 6 * 128)
    assert r["w"] <= 27 and r["htsz"] in (25, 27)
    # no room at all -> the reference prints nothing
    assert tune.tune_reference(1_000_000, 80, 80 * 64) is None


def test_reference_tune_on_b200_exceeds_its_own_kernel_limit():
    """148 SMs, 178 GB free: the procedure ends at -t 512 -b 592 with p in the thousands, i.e.
    8*N >= 2^31, which the reference kernel cannot index (signed 32-bit offsets, PTX197:L153-156)."""
    r = tune.tune_reference(178 * 10**9, 148, 148 * 128)
    assert (r["t"], r["b"], r["w"], r["htsz"]) == (512, 592, 31, 29)
    assert r["t"] * r["b"] * r["p"] * 8 >= 1 << 31


def test_own_suggestion_respects_limits_and_memory():
    s = tune.tune_b200(178 * 10**9, 148)
    N = s["t"] * s["b"] * s["p"]
    assert s["b"] == 148 and s["p"] % 2 == 0 and 8 * N < 1 << 31 and s["w_items"] < 1 << 32 and s["htsz"] <= 31
    assert 4 <= s["w_items"] / (1 << s["htsz"]) <= 8
    assert s["device_mb"] * 1048576 < 178 * 10**9
    small = tune.tune_b200(8 * 10**9, 40)
    assert small and small["device_mb"] * 1048576 < 8 * 10**9
    line = tune.format_try(s)
    assert line.startswith("Try -t 512 -b 148 -p ") and "-htsz" in line
    assert b.cli.parse_args(line.split("[")[0].split()[1:] + ["-pk", "1", "-pke", "100", "-pb", "02" + "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"])["p"] == s["p"]
