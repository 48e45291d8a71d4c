"""The oracle against committed outputs of the reference's OWN kernel.

tests/golden/ref_kernel_records.json holds the (code, idx) records the reference kernel `_test1`
(1_9_7File.pb:5181-23979, assembled for sm_100a) emitted on a B200 for seeded inputs
(tests/golden/make_ref_records.py).  The inputs are rebuilt here bit for bit from the seeds, so this
pins oracle/bsgs_oracle.c::orc_launch to the reference without a GPU and without the reference tree."""
import json
import os
from tests import pyec
from tests.helpers import build_case, giant, synthetic_keys, norm_hits

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_kernel_records.json")))


def _recs(r):
    return sorted((int(c), int(i)) for c, i in r)


def test_oracle_launch_equals_reference_kernel_records(orc):
    total = 0
    for c in GOLD["cases"]:
        T, B, p, w, htsz = c["T"], c["B"], c["p"], c["w"], c["htsz"]
        Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=c["seed"])
        assert (format(Ppt[0], "064x"), format(Ppt[1], "064x")) == (c["launches"][0]["px"], c["launches"][0]["py"])
        for l in c["launches"]:
            pt = (int(l["px"], 16), int(l["py"], 16))
            got = norm_hits(orc.launch(g2, T, B, p, gpu, htsz, pt))
            assert got == _recs(l["records"]) and len(got) == l["count"]
            total += len(got)
        assert set(plant) <= set(_recs(c["launches"][0]["records"]))
    assert total >= 100


def test_oracle_codes_4_and_5_equal_reference_kernel_records(orc):
    for c in GOLD["codes45"]:
        T, B, p, w, htsz, k = c["T"], c["B"], c["p"], c["w"], c["htsz"], c["k"]
        g2 = orc.build_g2(T, B, p, w)
        Q = giant(w, k)
        Ppt = Q if c["sign"] > 0 else pyec.neg(Q)
        dbl = pyec.add(Ppt, Ppt)
        cpu, gpu = orc.build_ht(synthetic_keys(w, [dbl[0], Ppt[0]], k), htsz)
        got = norm_hits(orc.launch(g2, T, B, p, gpu, htsz, Ppt))
        assert got == _recs(c["records"]) and (4, k) in got and (5, 0) in got
