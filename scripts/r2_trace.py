#!/usr/bin/env python3
"""How evenly do the warps of an SM progress?  One 16-window launch of config 2 with per-warp timestamps
(option trace): per SM the spread of warp finishing times, and the time during which an SM runs with fewer
than its 24 warps.  usage (GPU box): python scripts/r2_trace.py [blocks_per_sm] > gpurun_out/trace.json"""
import json, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsgs_cuda_b200 as b
from bsgs_cuda_b200 import ec
T, B, p, w, htsz = 512, 148, 256, 1 << 30, 28
opts = {}
if len(sys.argv) > 1:
    opts["blocks_per_sm"] = int(sys.argv[1])
e = b.Engine(T, B, p, w, htsz, **opts)
e.giants_generate(); e.ht_build_synthetic(seed=1)
s = b.Search(T, B, p, w, htsz, 1 << 64, ec.parse_pubkey("0230210c23b1a047bc9bdbb13448e67deddc108946de6de639bcc75d47c0216b1b"))
pts = [s.window_point(j) for j in range(32)]
e.time_launches(pts[:16], 16, 1)
e._ck(e.L.bsgs_set_option(e.h, b"trace", 1), "trace")
ms = e.time_launches(pts[16:], 16, 1)
tr = e.trace().astype(np.int64)
t0 = tr[:, 1].min()
sm, start, p1, end = tr[:, 0], tr[:, 1] - t0, tr[:, 2] - t0, tr[:, 3] - t0
total = end.max()
per_sm = []
for k in np.unique(sm):
    m = sm == k
    per_sm.append((int(k), int(m.sum()), int(start[m].max()), int(end[m].min()), int(end[m].max())))
ends_min = np.array([x[3] for x in per_sm]); ends_max = np.array([x[4] for x in per_sm])
# warp-seconds lost: time between a warp's end and the end of the whole kernel
idle_warp_frac = float((total - end).sum() / (total * len(end)))
out = {"ms_launch": ms, "kernel_ns": int(total), "warps": int(len(end)), "sms": len(per_sm),
       "start_skew_ns_max": int(start.max()), "pass1_end_ns": [int(p1.min()), int(np.median(p1)), int(p1.max())],
       "warp_end_ns": {"min": int(end.min()), "p10": int(np.percentile(end, 10)), "median": int(np.median(end)), "p90": int(np.percentile(end, 90)), "max": int(end.max())},
       "sm_first_warp_done_ns": {"min": int(ends_min.min()), "median": int(np.median(ends_min)), "max": int(ends_min.max())},
       "sm_last_warp_done_ns": {"min": int(ends_max.min()), "median": int(np.median(ends_max)), "max": int(ends_max.max())},
       "warp_time_idle_after_own_end_frac": idle_warp_frac,
       "within_sm_spread_ns_median": int(np.median(ends_max - ends_min))}
# the slowest SMs: which CTAs, when did their warps finish pass 1 and the kernel
order = sorted(per_sm, key=lambda x: -x[4])[:4] + sorted(per_sm, key=lambda x: x[4])[:2]
detail = []
for k, n, st, e0, e1 in order:
    m = sm == k
    ctas = sorted(set((np.nonzero(m)[0] // (e.info()["block"] // 32)).tolist()))
    detail.append({"sm": k, "warps": n, "ctas": ctas, "pass1_end_ns": [int(p1[m].min()), int(p1[m].max())], "end_ns": [int(end[m].min()), int(end[m].max())]})
out["slowest_and_fastest_sms"] = detail
out["sm_end_ns_by_smid"] = {str(k): int(e1) for k, n, st, e0, e1 in sorted(per_sm)}
out["sm_first_end_ns_by_smid"] = {str(k): int(e0) for k, n, st, e0, e1 in sorted(per_sm)}
print(json.dumps(out))
e.close()
