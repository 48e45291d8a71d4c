#!/usr/bin/env python3
"""BASELINE.md R5 / BASELINE.json configs[4]: max-table config (-w 31.5 -htsz 29).
  (i)  build the real 3,037,000,499-entry table on the GPU, report build time and image size;
  (ii) probe-only micro-kernel: 2^28 pseudo-random keys (seed 1) against it -> probes/s, sector GB/s;
  (iii) under torchrun: NCCL broadcast GB/s of the htGPU image (14.3 GB).
Also usable for configs[2] (-w 30.5 -htsz 29) with --w 30.5."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsgs_cuda_b200 as b
from bsgs_cuda_b200 import ec

ap = argparse.ArgumentParser()
ap.add_argument("--w", default="31.5"); ap.add_argument("--htsz", type=int, default=29)
ap.add_argument("--nkeys", type=int, default=1 << 28); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = ec.parse_w(a.w); T, B, p = 512, 148, 256
e = b.Engine(T, B, p, w, a.htsz, device=local)
out = {"w": w, "htsz": a.htsz, "n_gpus": world}
ptr, nbytes = e.ht_staging()
if rank == 0:
    t0 = time.time(); e.ht_generate(); out["table_build_s"] = time.time() - t0
out["htgpu_bytes"] = nbytes
if world > 1:
    class DevBuf:
        def __init__(s, p_, n): s.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p_, False), "version": 2}
    buf = torch.as_tensor(DevBuf(ptr, nbytes), device=f"cuda:{local}")
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); dist.broadcast(buf, src=0); ev1.record(); torch.cuda.synchronize()
    out["broadcast_ms"] = ev0.elapsed_time(ev1); out["broadcast_GBps"] = nbytes / ev0.elapsed_time(ev1) / 1e6
    if rank != 0:
        e.ht_commit()
e.probe_bench(a.nkeys, seed=1, iters=1)
ms, found = e.probe_bench(a.nkeys, seed=1, iters=a.steps)
out.update(probe_ms=ms / a.steps, probes_per_s=a.nkeys * a.steps / (ms * 1e-3), sector_GBps=a.nkeys * a.steps * 32 / (ms * 1e-3) / 1e9,
           found=int(found), expected_false_positive_rate=w / 2.0 ** (32 + a.htsz))
# giant-step throughput with this table
e.giants_generate()
s = b.Search(T, B, p, w, a.htsz, 1 << 79, ec.parse_pubkey("037e1238f7b1ce757df94faa9a2eb261bf0aeb9f84dbf81212104e78931c2a19dc"), rank=rank, world=world)
pts = [s.window_point(j) for j in range(16 * (a.steps + 2))]
e.time_launches(pts[:32], 16, 2)
ms = e.time_launches(pts[32:], 16, a.steps)
out["giant_steps_per_s_per_gpu"] = 2 * T * B * p * 16 * a.steps / (ms * 1e-3)
if world > 1:
    tt = torch.tensor([out["giant_steps_per_s_per_gpu"]], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MIN); out["giant_steps_per_s_total"] = float(tt[0]) * world
if rank == 0:
    print(json.dumps(out), flush=True)
e.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
