#!/usr/bin/env python3
"""bench.py -- giant-steps/s of the BSGS giant-step hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's own kernel, same box

A "step" is ONE launch of the hot path: `nwin` consecutive GetJob windows (1_9_7File.pb:2077-2092)
against the resident giant array and baby-step table = nwin * 2*T*B*p giant steps (the reference
counts 2*T*B*p per launch, 1_9_7File.pb:2371).  Workload at N=1 = BASELINE.json configs[1]:
puzzle #65 public key, -t 512 -b 148 -p 256 -w 30 -htsz 28.  At N>1 each rank owns a disjoint
contiguous sub-range of windows with the same per-GPU tables (weak scaling, no per-step collective;
one NCCL broadcast of the htGPU image at start-up).

Timing: device time with CUDA events on the launching stream (inside libbsgs_b200.so), bracketed by
barrier + synchronize, max over ranks.  The library keeps two launches in flight on two streams (the warps
of a launch finish unevenly; the CTAs of the next launch take the freed slots), so K steps are timed as a
whole: first event before launch 1 on either stream, last event after launch K on both.  Inputs (5.4 GB table + 1.2 GB giants) are far larger than
the 126 MB L2.  `e2e` goes through the reference-facing C-ABI call bsgs_run() with HOST buffers:
base point and stride in host memory in, hit records in host memory out, every launch's parameter
upload and record read-back inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PUZZLE65_PUB = "0230210c23b1a047bc9bdbb13448e67deddc108946de6de639bcc75d47c0216b1b"
PUZZLE65_PK, PUZZLE65_PKE = 0x10000000000000000, 0x1ffffffffffffffff
M32_PER_STEP = 256            # algorithmic 32x32->64 multiply-adds per giant step
IMAD_FALLBACK_LANES = 64.0    # lanes/clk/SM if the micro-benchmark cannot run


def workload(args):
    if args.small:            # plumbing-sized variant for quick checks (not a bench line)
        return dict(T=128, B=16, p=32, w=1 << 20, htsz=18, name="small (debug)")
    return dict(T=512, B=148, p=256, w=1 << 30, htsz=28, name="configs[1]: puzzle#65, -t 512 -b 148 -p 256 -w 30 -htsz 28")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows[-3:]]
        sm = sorted(int(float(r[0])) for r in rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(rows[0][1])) if rows else None,
                "reasons": reasons, "samples": len(rows), "power_w_max": max(pw) if pw else None}


KERNEL_SOURCES = ("giant_step.cuh", "fp.cuh")


def kernel_src_hash():
    """sha256 (16 hex digits) over the sources of the dominant kernel: profiles measured on another
    build of the kernel are not reported as this run's."""
    import hashlib
    h = hashlib.sha256()
    for n in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "bsgs-cuda_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def stamped_profile(name, nwin):
    """profiles/<name> if it was captured on exactly this kernel source (and launch shape), else None."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        return None
    j = json.load(open(path))
    if j.get("kernel_src_sha16") != kernel_src_hash() or j.get("windows_per_launch") != nwin:
        return None
    return j


def imad_peak(sm_mhz_max):
    """Measured integer-MAD issue rate of this box (bsgs-cuda_b200/imad_peak); lanes/clk/SM."""
    exe = os.path.join(ROOT, "bsgs-cuda_b200", "imad_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        j = json.loads(out)
        return j, "measured (imad_peak on this box)"
    except Exception as e:                                           # noqa: BLE001
        return {"sms": 148, "imad_wide_u32": {"lanes_per_clk_per_sm": IMAD_FALLBACK_LANES}}, f"fallback ({type(e).__name__})"


def dist_setup(args):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
    return rank, world, local, dist


class DevBuf:
    """A device allocation owned by libbsgs_b200.so, viewed as a torch uint8 tensor (for NCCL)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def run_b200(args):
    import torch
    import bsgs_cuda_b200 as b
    from bsgs_cuda_b200 import ec
    rank, world, local, dist = dist_setup(args)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = workload(args)
    T, B, p, w, htsz = wl["T"], wl["B"], wl["p"], wl["w"], wl["htsz"]
    N = T * B * p
    nwin = args.nwin
    opts = dict(max_windows=max(nwin, 1), blocks_per_sm=args.blocks_per_sm)
    if args.l2_fetch:
        opts["l2_fetch"] = args.l2_fetch
    if args.block != 256:
        opts["block"] = args.block
    e = b.Engine(T, B, p, w, htsz, device=local, **opts)
    t_setup = time.time()
    e.giants_generate()                                              # every rank: 0.1 s of GPU work, no point broadcasting
    bcast = None
    build_ht = (lambda: e.ht_build_synthetic(seed=args.seed)) if args.synthetic_table else e.ht_generate
    if world == 1:
        build_ht()
    else:                                                            # rank 0 builds, one NCCL broadcast of the htGPU image
        ptr, nbytes = e.ht_staging()
        if rank == 0:
            build_ht()
        buf = torch.as_tensor(DevBuf(ptr, nbytes), device=f"cuda:{local}")
        torch.cuda.synchronize(); dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); dist.broadcast(buf, src=0); ev1.record(); torch.cuda.synchronize()
        bcast = {"bytes": nbytes, "ms": ev0.elapsed_time(ev1), "GBps": nbytes / ev0.elapsed_time(ev1) / 1e6}
        if rank != 0:
            e.ht_commit()
    t_setup = time.time() - t_setup
    info = e.info()

    # windows: rank r owns the contiguous block of windows [r*Wr, (r+1)*Wr) of the 2^64 range
    target = ec.parse_pubkey(PUZZLE65_PUB)
    s = b.Search(T, B, p, w, htsz, PUZZLE65_PK, target)
    total_windows = s.windows_needed(PUZZLE65_PKE)
    per_rank = (total_windows + world - 1) // world
    first = rank * per_rank
    K, W = args.steps, max(args.warmup, 3)
    pts = [s.window_point((first + j) % total_windows) for j in range((K + W) * nwin)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput
    e.time_launches(pts[:W * nwin], nwin, W)                         # warm-up
    smp = ClockSampler(local); smp.start(); time.sleep(0.3)
    barrier(); t0 = time.time()
    ms = e.time_launches(pts[W * nwin:(W + K) * nwin], nwin, K)
    barrier(); t1 = time.time()
    clocks = smp.stop(t0, t1)
    # ---- end to end through bsgs_run (host buffers in, host records out)
    e.run(s.window_point(first), s.pubaddbig, W * nwin, batch=nwin)
    barrier(); te0 = time.perf_counter()
    hits, done = e.run(s.window_point(first), s.pubaddbig, K * nwin, batch=nwin)
    torch.cuda.synchronize(); te = (time.perf_counter() - te0) * 1e3
    barrier()
    if world > 1:
        tt = torch.tensor([ms, te], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, te = float(tt[0]), float(tt[1])
    steps_per_launch = 2 * N * nwin
    value = steps_per_launch * K * world / (ms * 1e-3)
    e2e = steps_per_launch * K * world / (te * 1e-3)

    line = None
    if rank == 0:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        sm_max = (clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        ip, ip_how = imad_peak(sm_max)
        lanes = ip["imad_wide_u32"]["lanes_per_clk_per_sm"]
        sms = ip.get("sms", 148)
        peak_m32 = sms * lanes * sm_max * 1e6                        # M32/s at max clock
        ach_m32 = value / world * M32_PER_STEP
        load = w / (1 << htsz)
        bytes_per_step = 8 + 4 * load + 32                           # probe (offset pair + tags) + giant stream
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # ncu figures are only reported when they were captured on THIS build of the kernel (source hash stamp)
        tj = None if args.small else stamped_profile("r2_traffic.json", nwin)
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"] if tj else None
        mix = None if args.small else stamped_profile("r2_inst_mix.json", nwin)
        line = {
            "metric": "giant-steps/s", "value": value, "unit": "giant-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (256-bit Fp)",
            "data": ("synthetic: uniformly random 64-bit table keys (seed %d) in the htGPU layout" % args.seed if args.synthetic_table else
                     "real baby-step table (w points (j+1)G, generated and sorted on the GPU)") + "; real giant array; real puzzle#65 base points",
            "config": {"workload": wl["name"], "T": T, "B": B, "p": p, "w": w, "htsz": htsz, "windows_per_launch": nwin,
                       "giant_steps_per_launch": steps_per_launch, "keys_per_giant_step": 2 * w, "l2": "inputs (6.6 GB) larger than L2",
                       "grid": info["grid"], "block": info["block"], "points_per_thread": info["points_per_thread"],
                       "launches_in_flight": 2,
                       "parallelism": f"{world} x disjoint window ranges"},
            "keys_per_s": value * 2 * w,
            "e2e": {"value": e2e, "unit": "giant-steps/s", "h2d_bytes_per_step": 128 * nwin, "d2h_bytes_per_step": 16 + 64 * 12,
                    "api": "bsgs_run (C ABI, host buffers)", "ms_per_step": te / K},
            "gpu_launches": K,
            "roofline": {"bound": "imad", "achieved": ach_m32 / 1e12, "peak": peak_m32 / 1e12, "unit": "T mad32/s", "frac": ach_m32 / peak_m32,
                         "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/r2_traffic.json)" if tj else
                         "null: no ncu capture of this exact kernel source (stamp mismatch)", "per_unit": f"{M32_PER_STEP} mad32 per giant step", "peak_source": ip_how,
                         "imad_lanes_per_clk_per_sm": lanes, "sm_mhz_for_peak": sm_max, "kernel_src_sha16": kernel_src_hash(),
                         # the contract counts 256 algorithmic mad32 per step; what the kernel EXECUTES (ncu instruction mix) is fewer
                         "products_executed_per_step": mix["imad_wide_per_step"] if mix else None,
                         "warp_instructions_per_point_window": mix["inst_per_point_window"] if mix else None,
                         "frac_by_executed_products": (value / world * mix["imad_wide_per_step"] / peak_m32) if mix else None},
            "roofline_hbm": {"bound": "hbm", "achieved": value / world * bytes_per_step / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": value / world * bytes_per_step / 1e9 / hbm_peak, "per_unit": f"{bytes_per_step:.1f} B per giant step",
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "clocks": clocks, "setup_s": t_setup, "hits_in_e2e": len(hits),
        }
        if bcast:
            line["ht_broadcast"] = bcast
        if world == 1 and not args.no_solve and not args.synthetic_table and not args.small:
            line["solve"] = solve_puzzle65(b, e, s, T, B, p, w, htsz, nwin)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(e, T, B, p, htsz, pts[:8])
    e.close()
    if not args.small and not args.no_extras:
        # driver-visible numbers for the other BASELINE.json configurations, outside the timed
        # regions above, tables built on the device in this same run
        extras = extra_configs(args, torch, dist, b, ec, rank, world, local)
        if line is not None:
            line.update(extras)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return line


PUZZLE80_PUB = "037e1238f7b1ce757df94faa9a2eb261bf0aeb9f84dbf81212104e78931c2a19dc"
PUZZLE80_PK = 0x80000000000000000000


def measure_config(args, torch, dist, b, ec, rank, world, local, w, htsz, steps, probe_keys=0):
    """One BASELINE.json table configuration at -t 512 -b 148 -p 256: table generated on rank 0's GPU,
    one NCCL broadcast of the htGPU image (N > 1), `steps` timed launches of args.nwin windows on every
    rank (disjoint windows of the puzzle #80 range), optionally the probe-only micro-kernel."""
    T, B, p = 512, 148, 256
    N, nwin = T * B * p, args.nwin
    dev = f"cuda:{local}"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t[0])

    out = {"w": w, "htsz": htsz}
    e = b.Engine(T, B, p, w, htsz, device=local, max_windows=nwin)
    e.giants_generate()
    t0 = time.time()
    if world == 1:
        e.ht_generate()
    else:
        ptr, nbytes = e.ht_staging()
        if rank == 0:
            e.ht_generate()
        buf = torch.as_tensor(DevBuf(ptr, nbytes), device=dev)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); dist.broadcast(buf, src=0); ev1.record(); torch.cuda.synchronize()
        ms = maxr(ev0.elapsed_time(ev1))
        out["ht_broadcast"] = {"bytes": nbytes, "ms": ms, "GBps": nbytes / ms / 1e6, "how": "one ncclBroadcast of the htGPU image, CUDA events, max over ranks"}
        if rank != 0:
            e.ht_commit()
    barrier()
    out["table_ready_s"] = time.time() - t0
    out["htgpu_image_bytes"] = 4 * (1 << htsz) + 4 + 4 * w
    s = b.Search(T, B, p, w, htsz, PUZZLE80_PK, ec.parse_pubkey(PUZZLE80_PUB))
    first = rank * 4096                                               # disjoint windows per rank
    pts = [s.window_point(first + j) for j in range((steps + 3) * nwin)]
    e.time_launches(pts[:3 * nwin], nwin, 3)
    barrier()
    ms = e.time_launches(pts[3 * nwin:], nwin, steps)
    barrier()
    ms = maxr(ms)
    out["giant_steps_per_s"] = 2 * N * nwin * steps * world / (ms * 1e-3)
    out["ms_per_launch"] = ms / steps
    out["launches"] = steps
    if probe_keys:
        # probe-only micro-kernel: uniformly random 64-bit keys (seed 1) against the real table
        e.probe_bench(probe_keys, seed=1, iters=1)
        barrier()
        pms, found = e.probe_bench(probe_keys, seed=2, iters=3)
        pms = maxr(pms)
        pr = probe_keys * 3 / (pms * 1e-3)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm = peaks.get("hbm_gbs", 6650.0)
        out["probe"] = {"keys": probe_keys, "launches": 3, "probes_per_s_per_gpu": pr, "found": int(found),
                        "sector_GBps": pr * 32 / 1e9, "dram_fill_GBps": pr * 64 / 1e9, "hbm_gbs": hbm, "frac_of_copy_peak": pr * 64 / 1e9 / hbm,
                        "what": "one random 32-byte fingerprint sector per probe (64-byte DRAM fill); exact CSR lookup only on a fingerprint match"}
    e.close()
    return out


def bcast_ht_check(b, world):
    """The reference's own multi-GPU model -- ONE process, one context per GPU (1_9_7File.pb:2337-2350, 4776):
    bsgs_bcast_ht copies the table of context 0 to contexts on the other devices; the images must be equal
    and every device must report the same records for the same planted base point."""
    import numpy as np
    from bsgs_cuda_b200 import ec
    T, B, p, w, htsz = 64, 8, 16, 1 << 22, 20
    t0 = time.time()
    es = [b.Engine(T, B, p, w, htsz, device=d) for d in range(world)]
    try:
        for e in es:
            e.giants_generate()
        es[0].ht_generate()
        b.native.bcast_ht(es)
        img0 = es[0].ht_export("gpu")
        equal = all(np.array_equal(img0, e.ht_export("gpu")) for e in es[1:])
        k, bb = T * B * p - 3, 777                                    # P + Q_k = bb*G -> record (1, k)
        P = ec.mul(ec.G, (k + 1) * 2 * w + bb)
        recs = [sorted(e.step(P)) for e in es]
        return {"devices": world, "image_bytes": int(img0.nbytes), "images_equal": bool(equal), "records_equal": all(r == recs[0] for r in recs),
                "planted_found": (1, k) in recs[0], "s": time.time() - t0, "api": "bsgs_bcast_ht (one process, one context per device)"}
    finally:
        for e in es:
            e.close()


def extra_configs(args, torch, dist, b, ec, rank, world, local):
    out = {}
    w3 = ec.parse_w("30.5"); w5 = ec.parse_w("31.5")
    out["config3"] = dict(measure_config(args, torch, dist, b, ec, rank, world, local, w3, 29, steps=5),
                          workload="configs[2]: -t 512 -b 148 -p 256 -w 30.5 -htsz 29 (puzzle #80 range), %d GPU(s)" % world)
    out["config5"] = dict(measure_config(args, torch, dist, b, ec, rank, world, local, w5, 29, steps=5, probe_keys=1 << 28),
                          workload="configs[4]: -w 31.5 -htsz 29 max table, %d GPU(s)" % world)
    out["config3_steps_per_s"] = out["config3"]["giant_steps_per_s"]
    out["probe"] = out["config5"]["probe"]
    if world > 1:
        out["ht_broadcast_max"] = out["config5"]["ht_broadcast"]
        torch.cuda.synchronize()
        cpu_group = dist.new_group(backend="gloo")                   # the other ranks wait on the HOST while rank 0 drives every device
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                out["bcast_ht_check"] = bcast_ht_check(b, world)
            except Exception as ex:                                   # noqa: BLE001
                out["bcast_ht_check"] = {"error": repr(ex)}
        dist.barrier(group=cpu_group)
    return out


def solve_puzzle65(b, e, s, T, B, p, w, htsz, nwin):
    """Whole problem, outside the timed region: recover the puzzle #65 key over [2^64, 2^65) with the
    resident tables (GetJob windows through bsgs_run, records decoded by the checkerThread mirror)."""
    e.set_verify(True)                                               # records are verified on the device (no htCPU on the host)
    t0 = time.time()
    key, done, seen = b.solve(e, None, T, B, p, w, htsz, PUZZLE65_PK, PUZZLE65_PKE, s.target, batch=nwin)
    dt = time.time() - t0
    e.set_verify(False)
    return {"key": hex(key) if key else None, "expected": "0x1a838b13505b26867", "ok": key == 0x1a838b13505b26867, "search_s": dt,
            "windows": done, "records_reaching_host": seen, "verification": "device (k_verify_hits)"}


def cpu_baseline(engine, T, B, p, htsz, pts, budget_s=12.0, tables=None):
    """The oracle port (C restatement of one launch, batch 4096 like #helpsize) on all host cores,
    plus the reference's own single-core sanity figure (10,000 sequential affine additions,
    lib/Curve64.pb:3303-3320, checked against its known answer)."""
    from oracle import oracle as orc
    orc.build(); orc.lib()
    g2, ht = tables if tables else (engine.giants_export(), engine.ht_export("gpu"))
    cores = os.cpu_count() or 1
    N = T * B * p
    t0 = time.time(); steps = 0; launches = 0
    while time.time() - t0 < budget_s:
        orc.launch(g2, T, B, p, ht, htsz, pts[launches % len(pts)], batch=4096, nthreads=cores, cap=1 << 16)
        steps += 2 * N; launches += 1
    dt = time.time() - t0
    add_s, add_pt = orc.time_seq_adds(10000)
    return {"value": steps / dt, "unit": "giant-steps/s", "cores": cores, "kind": "port", "per_core": steps / dt / cores,
            "sample": f"{launches} full launches ({steps} giant steps) of the same workload in {dt:.1f} s, oracle/bsgs_oracle.c orc_launch",
            "seq_adds_10000": {"ms": add_s * 1e3, "adds_per_s": 10000 / add_s, "cores": 1,
                               "known_answer_ok": add_pt[0] == 0xdb7432110ba814bfe6371ddfd03ba554b558548aa90e81b8e1421321656065a8,
                               "what": "b = G + b ten thousand times with the Curve64-style affine addition (lib/Curve64.pb:3303-3320)"}}


def check_tables_against_oracle(orc, g2, ht, T, B, p, w, htsz, nbuckets=10000, ngiants=1000, seed=7):
    """The reference arm's tables come from this repository's device builders; before anything is timed
    they are cross-checked against the CPU oracle: sampled buckets of the htGPU image (offsets monotone,
    tags ascending, first and random baby keys (j+1)G found by orc_probe_htgpu, perturbed keys not found)
    and sampled giants (orc_g2_get == (k+1) * ADDPUBG)."""
    import random
    import numpy as np
    rnd = random.Random(seed)
    H, N = 1 << htsz, T * B * p
    off = ht[:4 * (H + 1)].view(np.uint32)
    tags = ht[4 * (H + 1):].view(np.uint32)
    ok = int(off[0]) == 0 and int(off[H]) == w
    hs = np.array([rnd.randrange(H) for _ in range(nbuckets)])
    ok = ok and bool(np.all(off[hs + 1] >= off[hs]))
    for h in hs[:2000]:
        seg = tags[int(off[h]):int(off[h + 1])]
        ok = ok and bool(np.all(seg[1:] >= seg[:-1]))
    first = orc.baby_keys(min(w, nbuckets))                           # keys of (j+1)G, j < 10^4: 10^4 more buckets
    ok = ok and all(orc.probe_htgpu(ht, htsz, int(k)) for k in first)
    js = [w - 1, w // 2] + [rnd.randrange(w) for _ in range(200)]
    far = [orc.pt_mul((orc.GX, orc.GY), j + 1)[0] & (2**64 - 1) for j in js]
    ok = ok and all(orc.probe_htgpu(ht, htsz, k) for k in far)
    miss = sum(1 for k in far if orc.probe_htgpu(ht, htsz, k ^ (1 << 47)))
    D = orc.pt_neg(orc.pt_mul((orc.GX, orc.GY), 2 * w))
    ks = [0, 1, N - 1] + [rnd.randrange(N) for _ in range(ngiants - 3)]
    gok = all(orc.g2_get(g2, T, B, p, k) == orc.pt_mul(D, k + 1) for k in ks)
    return {"buckets_checked": nbuckets + len(first), "baby_keys_found": len(first) + len(far), "perturbed_keys_found": miss,
            "giants_checked": len(ks), "ok": bool(ok and gok and miss == 0)}


def run_reference(args):
    """The reference's OWN kernel (`_test1`, PTX embedded in 1_9_7File.pb:5181-23979, assembled for
    sm_100a by oracle/Makefile) through oracle/ref_harness.c on the same box -- one kernel instance per
    GPU, each rank on its own device taking its own windows, as the reference runs one worker thread per
    GPU over a shared job counter (1_9_7File.pb:2077-2092, 4772-4815).  Falls back to the CPU oracle port
    when oracle/_ref was not built."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import numpy as np
    import bsgs_cuda_b200 as b
    from bsgs_cuda_b200 import ec
    from oracle import oracle as orc
    orc.build(); orc.lib()
    dist = None
    if world > 1:
        import torch.distributed as dist                              # host-side plumbing only (gloo): barrier + max over ranks
        dist.init_process_group("gloo")
    wl = workload(args)
    T, B, p, w, htsz = wl["T"], wl["B"], wl["p"], wl["w"], wl["htsz"]
    N = T * B * p
    K, W = args.steps, max(args.warmup, 3)
    target = ec.parse_pubkey(PUZZLE65_PUB)
    s = b.Search(T, B, p, w, htsz, PUZZLE65_PK, target)
    total_windows = s.windows_needed(PUZZLE65_PKE)
    per_rank = (total_windows + world - 1) // world
    pts = [s.window_point((rank * per_rank + j) % total_windows) for j in range(max(K, 8))]
    have_ref = os.path.exists(orc.REF_LIB) and os.path.exists(orc.REF_CUBIN)
    # tables: built by this repository's device builders (byte-identical to the oracle's at test sizes), cross-checked below
    e = b.Engine(T, B, p, w, htsz, device=local)
    e.giants_generate(); e.ht_generate()
    g2 = e.giants_export()
    line = None
    if rank == 0:
        ht = e.ht_export("gpu")
        line = {"impl": "reference", "metric": "giant-steps/s", "unit": "giant-steps/s", "n_gpus": world if have_ref else 0, "steps": K, "warmup": W,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (256-bit Fp)",
                "data": "real baby-step table and giant array; real puzzle#65 base points",
                "config": {"workload": wl["name"], "T": T, "B": B, "p": p, "w": w, "htsz": htsz, "windows_per_launch": 1,
                           "parallelism": f"{world} x one reference kernel per GPU, disjoint window ranges"},
                "tables": dict(check_tables_against_oracle(orc, g2, ht, T, B, p, w, htsz),
                               source="bsgs_giants_generate / bsgs_ht_generate of this repository (the reference builds them on CPU threads for hours); "
                                      "sampled against oracle/bsgs_oracle.c before timing")}
        if world == 1 or not have_ref:
            line["cpu_baseline"] = cpu_baseline(e, T, B, p, htsz, pts[:8], budget_s=10.0, tables=(g2, ht))
        del ht
    if have_ref:
        ref = orc.RefKernel(T, B, p, w, htsz, dev=local)
        ref.upload_giants(g2)
        ptr, nbytes = e.ht_device_image(); ref.upload_ht_dev(ptr, nbytes)
        # parity, outside the timed region: the product's multi-window launch against one reference launch per window
        nw = min(16, len(pts))
        mine = sorted((wi, c, 0 if c == 5 else i) for c, i, wi in e.step_multi(pts[:nw]))
        theirs = sorted((wi, c, i) for wi in range(nw) for c, i in ref.launch(pts[wi])[0])
        parity = {"windows": nw, "records": len(theirs), "equal": mine == theirs, "what": "sorted (window, code, idx) records: bsgs_step_multi vs one _test1 launch per window, same tables"}
        dropin_cubin = os.path.join(ROOT, "bsgs-cuda_b200", "dropin_test1.cubin")
        mine_k = None
        if args.with_dropin and rank == 0 and os.path.exists(dropin_cubin):   # boundary option A: our module behind the reference's own interface
            mine_k = orc.RefKernel(T, B, p, w, htsz, dev=local, cubin=dropin_cubin)
            mine_k.upload_giants(g2); mine_k.upload_ht_dev(ptr, nbytes)
        e.close()
        del g2

        def barrier():
            if dist is not None:
                dist.barrier()

        def timed(k):
            k.time_launches(pts[0], W)
            barrier()
            ms = sum(k.time_launches(pts[i % len(pts)], 1) for i in range(K))     # CUDA events around each launch, synchronised
            barrier()
            return ms

        def timed_e2e(k):
            """cuda()'s loop as it is (1_9_7File.pb:2435-2509): 64 B base point up, launch, synchronise, 4 B counter
            (+ records) down, host buffers throughout; wall clock."""
            barrier(); t0 = time.perf_counter()
            for i in range(K):
                k.launch(pts[i % len(pts)])
            dt = (time.perf_counter() - t0) * 1e3
            barrier()
            return dt
        ms = timed(ref); te = timed_e2e(ref)
        if dist is not None:
            import torch
            tt = torch.tensor([ms, te], dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms, te = float(tt[0]), float(tt[1])
            pe = torch.tensor([1.0 if parity["equal"] else 0.0]); dist.all_reduce(pe, op=dist.ReduceOp.MIN); parity["equal_on_all_ranks"] = bool(pe[0] == 1.0)
        ref.close()
        if line is not None:
            value = 2 * N * K * world / (ms * 1e-3)
            line.update({"value": value, "ms_per_step": ms / K, "gpu_launches": K, "parity": parity,
                         "reference_kind": "the reference's own GPU kernel _test1 (PTX from 1_9_7File.pb, ptxas -arch=sm_100a) on this box, one instance per GPU; "
                                           "the reference has no CPU giant-step loop, see cpu_baseline (N=1) for the CPU restatement",
                         "e2e": {"value": 2 * N * K * world / (te * 1e-3), "unit": "giant-steps/s", "h2d_bytes_per_step": 64, "d2h_bytes_per_step": 4,
                                 "api": "oracle/ref_harness.c ref_launch = cuda()'s per-launch sequence (1_9_7File.pb:2435-2509), host buffers", "ms_per_step": te / K}})
            if mine_k is not None:
                ms2 = timed(mine_k) if dist is None else None
                if ms2:
                    line["dropin_module"] = {"value": 2 * N * K / (ms2 * 1e-3), "unit": "giant-steps/s", "ms_per_step": ms2 / K,
                                             "what": "bsgs-cuda_b200/dropin_test1.cubin (entry _test1, const _A, the reference memory map) in the same harness, same tables, one window per launch"}
                mine_k.close()
    else:
        e.close()
        if line is not None:
            cpu = line["cpu_baseline"]
            line.update({"value": cpu["value"], "ms_per_step": None, "gpu_launches": 0, "n_gpus": 0,
                         "reference_kind": "oracle port on host cores (oracle/_ref not built)",
                         "e2e": {"value": cpu["value"], "unit": "giant-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)      # at least 50 launches after warm-up
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nwin", type=int, default=16, help="GetJob windows per launch")
    ap.add_argument("--blocks-per-sm", type=int, default=3)
    ap.add_argument("--block", type=int, default=256, help="threads per CTA: 256 (x blocks-per-sm), 320 (x 2), 128 (x 6) or 64 (x 12)")
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (32|64|128), 0 = leave")
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--synthetic-table", action="store_true", help="uniform random keys instead of the real baby-step table")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-3 / config-5 / probe / broadcast measurements")
    ap.add_argument("--with-dropin", action="store_true", help="reference arm only: also time bsgs-cuda_b200/dropin_test1.cubin in the same harness (extra key, off by default)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there meanwhile (NCCL's version
    # banner under NCCL_DEBUG=VERSION, for one) is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
