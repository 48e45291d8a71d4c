#!/usr/bin/env python3
"""Large-N addressing check (run on a B200; ~150 GB of HBM): N = 512*148*3400 = 257,638,400 giants,
16 windows per launch -> 8.2e9 chain elements of 16 bytes, i.e. element offsets beyond 2^32 and byte
offsets beyond 2^37.  Planted hits at the extreme giant indices of the first and last window must be
reported, and the 16-window launch must equal 16 one-window launches."""
import os, random, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsgs_cuda_b200 as b
from oracle import oracle as orc
from tests.helpers import random_point, candidate_x, synthetic_keys, norm_hits

orc.build(); orc.lib()
T, B, p, w, htsz = 512, 148, int(os.environ.get("BIGN_P", "3400")), 1 << 20, 16
N = T * B * p
rnd = random.Random(99)
pts = [random_point(rnd) for _ in range(16)]
plant = [(0, 1, 0), (0, 2, N - 1), (15, 1, N - 1), (15, 2, 0), (7, 1, N // 2 + 12345), (3, 2, N - 113665), (9, 1, 113663), (12, 2, 2**27 + 1)]
keys = synthetic_keys(w, [candidate_x(pts[wi], w, k, c) for wi, c, k in plant], 5)
cpu, gpu = orc.build_ht(keys, htsz)
t0 = time.time()
with b.Engine(T, B, p, w, htsz, device=0) as e:
    e.giants_generate(); e.upload_ht(gpu)
    info = e.info()
    t1 = time.time()
    multi = sorted((c, i, wd) for c, i, wd in e.step_multi(pts))
    t2 = time.time()
    single = sorted((c, i, wi) for wi, pt in enumerate(pts) for c, i in norm_hits(e.step(pt)))
    ok_plant = all((c, k, wi) in multi for wi, c, k in plant)
    # bsgs_run over two launches: the second chain (another 132 GB) cannot be allocated on a 180 GB part, so the
    # library must fall back to one launch at a time, quietly, with the same records
    from tests import pyec
    D = pyec.mul(pyec.G, 987654321)
    seq = [pts[0]]
    for _ in range(31):
        seq.append(pyec.add(seq[-1], D))
    run_hits, done = e.run(pts[0], D, 32, batch=16)
    two = sorted(run_hits)
    ref2 = sorted([(c, i, wd) for c, i, wd in e.step_multi(seq[:16])] + [(c, i, wd + 16) for c, i, wd in e.step_multi(seq[16:])])
    run_ok = done == 32 and two == ref2
    print(json.dumps({"N": N, "points_per_thread": info["points_per_thread"], "bytes_chain": info["bytes_chain"], "setup_s": round(t1 - t0, 2),
                      "launch_s": round(t2 - t1, 3), "giant_steps_per_s": 2 * N * 16 / (t2 - t1), "records": len(multi),
                      "planted_found": ok_plant, "multi_equals_single": multi == single, "run_two_launches_equals_step_multi": run_ok, "run_records": len(two)}))
    assert ok_plant and multi == single and run_ok
