"""bench.py's host-side helpers (no GPU): the ncu figures it reports are tied to the kernel sources they were
captured on, and the reference arm's table cross-check really checks."""
import json
import os
import numpy as np
import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_profile_stamps_follow_the_kernel_sources(tmp_path, monkeypatch):
    h = bench.kernel_src_hash()
    assert len(h) == 16 and h == bench.kernel_src_hash()
    for name in ("r2_traffic.json", "r2_inst_mix.json"):             # the committed captures carry a stamp and the launch shape
        j = json.load(open(os.path.join(ROOT, "profiles", name)))
        assert len(j["kernel_src_sha16"]) == 16 and j["windows_per_launch"] == 16
    # a capture of other sources, or of another launch shape, is not reported
    prof = tmp_path / "profiles"; prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "kernel_src_hash", lambda: h)
    (prof / "x.json").write_text(json.dumps({"kernel_src_sha16": h, "windows_per_launch": 16, "v": 1}))
    (prof / "y.json").write_text(json.dumps({"kernel_src_sha16": "0" * 16, "windows_per_launch": 16}))
    assert bench.stamped_profile("x.json", 16)["v"] == 1
    assert bench.stamped_profile("x.json", 8) is None and bench.stamped_profile("y.json", 16) is None and bench.stamped_profile("nope.json", 16) is None


def test_table_cross_check_accepts_oracle_tables_and_rejects_damaged_ones(orc):
    T, B, p, w, htsz = 8, 4, 8, 20000, 10
    cpu, gpu = orc.build_ht(orc.baby_keys(w), htsz)
    g2 = orc.build_g2(T, B, p, w)
    ok = bench.check_tables_against_oracle(orc, g2, gpu, T, B, p, w, htsz, nbuckets=1000, ngiants=100)
    assert ok["ok"] and ok["giants_checked"] == 100 and ok["perturbed_keys_found"] == 0
    H = 1 << htsz
    bad = gpu.copy(); bad[4 * (H + 1):4 * (H + 1) + 4 * 2000] = 0          # tags of the first buckets wiped: the first baby keys are gone
    assert not bench.check_tables_against_oracle(orc, g2, bad, T, B, p, w, htsz, nbuckets=1000, ngiants=100)["ok"]
    bad_g = g2.copy(); bad_g[:] = np.roll(bad_g, 4)
    assert not bench.check_tables_against_oracle(orc, bad_g, gpu, T, B, p, w, htsz, nbuckets=1000, ngiants=100)["ok"]
