"""Command line with the reference's flag surface (getprogparam, 1_9_7File.pb:875-1042):

    python -m bsgs_cuda_b200 -t 512 -b 148 -p 256 -w 30 -htsz 28 -pb <pubkey> -pk 0x1.. -pke 0x1f.. [-d 0,1]
                             [-infile pubs.txt] [-wl currentwork.txt] [-wt 180]

What it keeps from the reference: flag names and meanings, `-w` parsing (2^x truncated or decimal),
pubkey forms, the table file names in the working directory (loaded when present with the right
size, otherwise generated ON THE GPU and written so that the reference can load them), the per-target
banner lines, `win.txt` (1_9_7File.pb:5151-5157), `currentwork.txt` with its SHA-1 settings
fingerprint (1_9_7File.pb:3913-3921, 4634-4636), sequential `-infile` processing with tables loaded
once (1_9_7File.pb:4995-5168).  What it drops: the interactive "press ENTER" prompts
(1_9_7File.pb:4400, 754-764) and the OpenCL environment variables.  One engine per `-d` device, each
driven by its own host thread like CreateThread(@cuda(), i) (1_9_7File.pb:4776); windows are dealt
round-robin (rank r of R owns j = r mod R), the saved counter is the minimum over the devices
(1_9_7File.pb:3904-3911)."""
from __future__ import annotations
import hashlib
import math
import os
import sys
import threading
import time
import numpy as np
from . import ec, formats
from . import native
from .native import Engine
from .search import Search

DEFAULTS = dict(t=256, b=132, p=400, w=1 << 25, htsz=25, pk=0x1, pke=0x1ffffffffffffffff, d="0", wt=180,
                pb="036d05521c67b9cc1c0ef906b42215c7120c7302c34d9316a2726199bedac50936")     # 1_9_7File.pb:181-210, 168


class UsageError(Exception):
    pass


def cut_hex(s: str) -> str:                                          # cutHex, 1_9_7File.pb:864-873
    s = s.strip().lower()
    return s[2:] if s.startswith("0x") else s


def parse_args(argv):
    o = dict(DEFAULTS)
    o.update(infile=None, wl=None, sf=1, pb_given=False, pk_str="01", pke_str="1ffffffffffffffff", nosave=False, onlygen=False)
    i = 0
    takes = {"-t": ("t", int), "-b": ("b", int), "-p": ("p", int), "-d": ("d", str), "-htsz": ("htsz", int), "-infile": ("infile", str),
             "-wl": ("wl", str), "-sf": ("sf", int)}
    while i < len(argv):
        a = argv[i].lower()
        val = argv[i + 1] if i + 1 < len(argv) and not (argv[i + 1].startswith("-") and len(argv[i + 1]) > 1 and not argv[i + 1][1].isdigit()) else None
        if a == "-h":
            raise UsageError(__doc__)
        if a == "-nosave":                                           # extension: keep the generated tables in HBM/RAM only
            o["nosave"] = True; i += 1; continue
        if a == "-onlygen":                                          # the role of onlygen1_9_6File.pb: write the table files and stop
            o["onlygen"] = True; i += 1; continue
        if a in takes:
            if val is not None:
                o[takes[a][0]] = takes[a][1](val); i += 1
        elif a == "-pb":
            if val is not None:
                o["pb"] = cut_hex(val); o["pb_given"] = True; i += 1
        elif a == "-pk":
            if val is not None:
                o["pk"] = int(cut_hex(val), 16); o["pk_str"] = cut_hex(val); i += 1
        elif a == "-pke":
            if val is not None:
                o["pke"] = int(cut_hex(val), 16); o["pke_str"] = cut_hex(val); i += 1
        elif a == "-w":
            if val is not None:
                o["w"] = ec.parse_w(val); i += 1
        elif a == "-wt":
            if val is not None:
                o["wt"] = max(int(val), 30); i += 1                  # 1_9_7File.pb:935-937
        else:
            raise UsageError(f"Unknown parameter [{argv[i]}]")      # 1_9_7File.pb:1033
        i += 1
    return o


def validate(o):
    """Structural limits of the formats (1_9_7File.pb:4412-4472, 4616-4618, 4904-4922)."""
    if o["w"] >= 3069485951:
        raise UsageError("-w should be less than 3069485951")
    if o["htsz"] > 31:
        raise UsageError("-htsz should be less or equal to 31")
    if o["p"] % 2:
        raise UsageError("vaule -p must be a multiple of 2")
    if o.get("onlygen") and (o["t"] % 2 or o["b"] % 2):
        raise UsageError("vaule -t and -b must be a multiple of 2")   # onlygen1_9_6File.pb:3140-3145
    if o["pk"] == 0:
        raise UsageError("Range begin should be more than 0")
    if o["pke"] <= o["pk"]:
        raise UsageError("Range end should be more than range begin")
    lim = {27: 1331331443, 28: 1777178603}.get(o["htsz"])
    warn = []
    if lim and o["w"] > lim:
        warn.append(f"UNSAFE: with htsz {o['htsz']} value -w should be <= {lim}")
    if int(math.log2(o["w"])) - o["htsz"] > 3:
        warn.append("Warning: hashtable is small for this -w, probes will be slow")
    return warn


def fingerprint(o) -> str:
    """SHA-1 of Str(T)+Str(B)+Str(p)+Str(w)+pk+pke+Str(htsz) (1_9_7File.pb:4634-4636); pk/pke are the
    argument strings after cutHex + LCase (1_9_7File.pb:990-1004).  The bytes hashed are those of an
    ASCII build of the reference (PureBasic 5.31 default); a unicode build would hash UTF-16."""
    s = f"{o['t']}{o['b']}{o['p']}{o['w']}{o['pk_str']}{o['pke_str']}{o['htsz']}"
    return hashlib.sha1(s.encode("ascii")).hexdigest()


def load_or_build_tables(o, engines, log):
    """htGPU + G2 on every device, htCPU on the host (the checker's table), files as the reference names them."""
    T, B, p, w, htsz = o["t"], o["b"], o["p"], o["w"], o["htsz"]
    f_cpu, f_gpu, f_g2 = formats.htcpu_name(w, htsz), formats.htgpu_name(w, htsz), formats.g2_name(T, B, p, w)
    e0 = engines[0]
    t0 = time.time()
    have = lambda f, n: os.path.exists(f) and os.path.getsize(f) == n                       # noqa: E731
    if have(f_cpu, formats.htcpu_bytes(w, htsz)) and have(f_gpu, formats.htgpu_bytes(w, htsz)):
        log(f"Load hashtable from {f_gpu}")
        htcpu = formats.load(f_cpu)
        e0.upload_htcpu(htcpu)                                   # tags + payloads: enables device-side verification
        if len(engines) > 1:
            native.bcast_ht(engines)
    else:
        log(f"Generate {w} baby points on GPU#{o['devices'][0]}")
        e0.ht_generate()
        htcpu = None                                             # records are verified on the device: no host copy needed
        if not o["nosave"]:
            htcpu = e0.ht_export("cpu")
            formats.save(f_cpu, htcpu); formats.save(f_gpu, e0.ht_export("gpu"))
        if len(engines) > 1:
            native.bcast_ht(engines)                             # device to device, not 7 uploads from the host
    for e in engines:
        e.set_verify(True)
    if have(f_g2, formats.g2_bytes(T, B, p)):
        log(f"Load giants from {f_g2}")
        g2 = formats.load(f_g2)
        for e in engines:
            e.upload_giants(g2)
    else:
        log(f"Generate {T * B * p} giant points on GPU")
        for e in engines:
            e.giants_generate()
        if not o["nosave"]:
            formats.save(f_g2, e0.giants_export())
    log(f"Tables ready in {time.time() - t0:.1f}s")
    return htcpu


def read_pubs(o):
    if o["infile"]:
        if not os.path.exists(o["infile"]):
            raise UsageError(f"File[{o['infile']}] not found!")
        return [l.strip() for l in open(o["infile"]) if l.strip()]
    return [o["pb"]]


def list_devices(log):
    """The start-up device listing with the parameter suggestion (1_9_7File.pb:778-860, Tune 324-431)."""
    from . import native, tune
    for d in native.devices():
        log(f"Cuda device:{d['name']} ({d['free_bytes'] / 1048576:.3f}/{d['total_bytes'] // 1048576}MB)")
        cores = tune.cores_for(d["cc"][0], d["cc"][1], d["sm_count"])
        log(f"Device have: MP:{d['sm_count']} Cores+{cores}")
        if d["sm_count"] and cores:
            ref = tune.tune_reference(d["free_bytes"], d["sm_count"], cores, ncpu=os.cpu_count() or 1)
            if ref:
                log(tune.format_try(ref) + "   (reference Tune)")
            own = tune.tune_b200(d["free_bytes"], d["sm_count"])
            if own:
                log(tune.format_try(own) + "   (this build)")
        log("---------------")


def main(argv=None, out=sys.stdout) -> int:
    log = lambda s: (out.write(s + "\n"), out.flush())               # noqa: E731
    t_program = time.time()                                          # begintime, 1_9_7File.pb:5135 (Tt:)
    try:
        o = parse_args(list(sys.argv[1:] if argv is None else argv))
        for wmsg in validate(o):
            log(wmsg)
        pubs = read_pubs(o)
    except UsageError as ex:
        log(str(ex)); return 2
    o["devices"] = [int(x) for x in o["d"].split(",") if x != ""]
    try:
        list_devices(log)
    except Exception as ex:                                            # noqa: BLE001 -- the listing is informational
        log(f"device listing unavailable: {ex}")
    T, B, p, w, htsz = o["t"], o["b"], o["p"], o["w"], o["htsz"]
    fp = fingerprint(o)
    log(f"Current config hash[{fp}]")
    rec = None
    if o["wl"]:
        if not os.path.exists(o["wl"]):
            log(f"Recovery file [{o['wl']}] does not exist"); return 2
        lines = [l.strip() for l in open(o["wl"])][:4]
        if len(lines) < 4 or not all(lines):
            log("Empty line in recovery file"); return 2
        rec = dict(pos=int(lines[0]), pub=lines[1], cnt=int(lines[2], 16), fp=lines[3])
        if rec["fp"] != fp:
            log("Current and recovery configuration are different"); return 2
    engines = [Engine(T, B, p, w, htsz, device=d) for d in o["devices"]]
    R = len(engines)
    try:
        if o["onlygen"]:                                             # onlygen1_9_6File.pb:3182-3205
            o["nosave"] = False
            load_or_build_tables(o, engines, log)
            log("finished ok")
            return 0
        htcpu = load_or_build_tables(o, engines, log)
        found_total = 0
        for listpos, pub in enumerate(pubs, start=1):
            target = ec.parse_pubkey(pub)
            full = format(target[0], "064x") + format(target[1], "064x")
            if rec and listpos != rec["pos"]:
                continue                                             # 1_9_7File.pb:5020-5024
            if rec and rec["pub"] not in (full, pub.lower()):
                log("Find position but the keys are different"); return 2
            start_key = rec["cnt"] if rec else 1
            rec = None
            log(""); log(f"Findpubkey  : {ec.compress(target)}")
            searches = [Search(T, B, p, w, htsz, o["pk"], target, start_key=start_key, rank=r, world=R) for r in range(R)]
            log(f"Searchpubkey: {ec.compress(searches[0].findpub)}")
            state = dict(key=None, stop=False, done=[0] * R, steps=[0] * R, lock=threading.Lock())
            t_begin = time.time()

            def worker(r):
                s, e = searches[r], engines[r]
                nwin = s.windows_needed(o["pke"])
                done = 0
                while done < nwin and not state["stop"]:
                    chunk = min(nwin - done, 128)                  # 8 launches of 16 windows, two in flight inside bsgs_run
                    hits, d = e.run(s.window_point(done), s.D, chunk, batch=min(chunk, 16), stop_after_hits=1)
                    for (code, idx, win), payload in zip(hits, list(e.payloads)):
                        if payload != 0xFFFFFFFF:
                            k = s.check_verified(code, idx, done + win, payload)
                        elif htcpu is not None:
                            k = s.check_hit(htcpu, code, idx, done + win)
                        else:
                            raise RuntimeError("unverified record and no htCPU table on the host")
                        if k is not None:
                            with state["lock"]:
                                state["key"], state["stop"] = k, True
                    done += d
                    with state["lock"]:
                        state["done"][r] = done; state["steps"][r] += 2 * s.N * d

            def saver():                                             # saveCurentCNT, 1_9_7File.pb:3897-3931
                last = time.time()
                while not state["stop"]:
                    time.sleep(0.5)
                    if time.time() - last >= o["wt"]:
                        save_checkpoint(listpos, full, min(searches[r].window_key(state["done"][r]) for r in range(R)), fp)
                        last = time.time()

            threads = [threading.Thread(target=worker, args=(r,)) for r in range(R)]
            sv = threading.Thread(target=saver, daemon=True); sv.start()
            for th in threads:
                th.start()
            last_log = time.time(); last_steps = [0] * R
            while True:
                alive = [th for th in threads if th.is_alive()]
                if not alive:
                    break
                alive[0].join(0.05)                                  # returns at once when the worker ends (no polling quantum per target)
                if time.time() - last_log > 2:                       # status line, 1_9_7File.pb:5119-5141
                    with state["lock"]:
                        st = list(state["steps"])
                    dt = time.time() - last_log
                    rates = [(a - c) / dt for a, c in zip(st, last_steps)]; last_steps = st; last_log = time.time()
                    cnt = min(searches[r].window_key(state["done"][r]) for r in range(R))
                    log(format_status(cnt, len(alive), rates, w, time.time() - t_begin, time.time() - t_program))
            state["stop"] = True
            if state["key"] is not None:
                tag = f"KEY[{listpos}]: "
                lines = [f"{tag}0x{state['key']:064x}", " " * (len(tag) - 5) + "Pub: " + ec.compress(target)]
                for l in lines:
                    log(l)
                with open("win.txt", "a") as f:                      # #WINFILE, 1_9_7File.pb:19, 5151-5157
                    f.write("\n".join(lines) + "\n")
                log(f"Job time {time.strftime('%H:%M:%S', time.gmtime(time.time() - t_begin))}s")
                found_total += 1
            else:
                log("Reached end of space")
        log(f"finished, found {found_total} of {len(pubs)}")
        return 0
    finally:
        for e in engines:
            e.close()


def _hms(sec):
    sec = int(sec)
    return f"{sec // 3600 % 24:02d}:{sec // 60 % 60:02d}:{sec % 60:02d}"


def format_status(cnt, running, rates, w, job_s, total_s):
    """The status line of the main loop (1_9_7File.pb:5119-5141): counter, workers alive, giant steps/s of
    every GPU and in total in units of 2^20, the keys each step covers (2w), job time and total time."""
    total = sum(int(r) for r in rates)
    wald = math.log2(2 * w)
    hashd = math.log2(total) if total else 0
    perf = "".join(f"{int(r) >> 20} " for r in rates)
    return (f"Cnt:{cnt:x} [{running}][ {perf}] = {total >> 20} MKeys/s x2^{wald:.1f}=2^{hashd + wald:.2f} "
            f"Jt:{_hms(job_s)} Tt:{_hms(total_s)}")


def save_checkpoint(listpos, pub_full_hex, cnt, fp, path="currentwork.txt"):
    with open("currentwork.temp", "w") as f:
        f.write(f"{listpos}\n{pub_full_hex}\n{cnt:064x}\n{fp}\n")
    os.replace("currentwork.temp", path)
