#!/usr/bin/env python3
"""Generate tests/golden/ref_kernel_records.json ON THE GPU BOX: outputs of the reference's OWN
kernel (`_test1`, PTX embedded in 1_9_7File.pb:5181-23979, decoded + assembled into oracle/_ref/ by
oracle/Makefile) run through oracle/ref_harness.c on seeded inputs.  The inputs are not stored:
tests/helpers.build_case() rebuilds them bit for bit from (T, B, p, w, htsz, seed) -- giants
Q_k = (k+1)(-2wG), a synthetic 64-bit-key table with planted hits -- so the CPU suite can check the
oracle against what the reference kernel really emitted, with no GPU and no reference tree.

    gpurun -- 'python tests/golden/make_ref_records.py'      (then copy gpurun_out/ref_kernel_records.json here)
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc                                    # noqa: E402
from tests import pyec                                              # noqa: E402
from tests.helpers import build_case, random_point, giant, synthetic_keys   # noqa: E402

CASES = [(32, 4, 16, 4096, 10, 101), (64, 20, 100, 60000, 14, 102), (128, 2, 8, 5000, 4, 103), (512, 8, 6, 30000, 13, 104), (7, 3, 5, 1000, 8, 105)]


def main():
    orc.build(); orc.lib()
    out = {"source": "reference kernel _test1 (1_9_7File.pb:5181-23979) assembled with ptxas -arch=sm_100a, run on a B200 through oracle/ref_harness.c",
           "cases": [], "codes45": []}
    for T, B, p, w, htsz, seed in CASES:
        Ppt, plant, cpu, gpu, g2 = build_case(orc, T, B, p, w, htsz, seed=seed)
        ref = orc.RefKernel(T, B, p, w, htsz); ref.upload(g2, gpu)
        rnd = random.Random(seed)
        launches = []
        for pt in [Ppt] + [random_point(rnd) for _ in range(2)]:
            hits, count, _ = ref.launch(pt)
            assert count <= 240
            launches.append({"px": format(pt[0], "064x"), "py": format(pt[1], "064x"), "count": count, "records": hits})
        ref.close()
        out["cases"].append({"T": T, "B": B, "p": p, "w": w, "htsz": htsz, "seed": seed, "launches": launches})
    # planted codes 4 and 5 (P = +-Q_k: tangent and self probe)
    T, B, p, w, htsz = 32, 4, 8, 2000, 8
    N = T * B * p
    g2 = orc.build_g2(T, B, p, w)
    for k, sign in ((5, 1), (N - 1, -1)):
        Q = giant(w, k)
        Ppt = Q if sign > 0 else pyec.neg(Q)
        dbl = pyec.add(Ppt, Ppt)
        cpu, gpu = orc.build_ht(synthetic_keys(w, [dbl[0], Ppt[0]], k), htsz)
        ref = orc.RefKernel(T, B, p, w, htsz); ref.upload(g2, gpu)
        hits, count, _ = ref.launch(Ppt)
        ref.close()
        out["codes45"].append({"T": T, "B": B, "p": p, "w": w, "htsz": htsz, "k": k, "sign": sign, "count": count, "records": hits})
    dst = os.path.join(ROOT, "gpurun_out", "ref_kernel_records.json")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, sum(len(l["records"]) for c in out["cases"] for l in c["launches"]), "records")


if __name__ == "__main__":
    main()
