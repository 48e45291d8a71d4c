#!/usr/bin/env python3
"""Turn the captures of scripts/r2_profile.sh (gpurun_out/final.ncu-rep, gpurun_out/launches_final.csv) into the
tracked files under profiles/: the ncu summary, the launch list, and the two stamped JSON files bench.py reads
(DRAM traffic and executed-instruction mix of one 16-window launch).  The stamp is the hash of the kernel sources
the captures were taken on: run this right after the capture, before touching giant_step.cuh / fp.cuh."""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
tag = sys.argv[1] if len(sys.argv) > 1 else "final"
rep = os.path.join(ROOT, "gpurun_out", "final.ncu-rep")
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(ROOT, "profiles", f"r2_ncu_{tag}_w16.txt"), "w").write(summ)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
def val(k):
    v = float(d[k].replace(",", "")); unit = u.get(k, "")
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(unit, 1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h = rows[1]; ix = {x: i for i, x in enumerate(h)}
op = collections.Counter()
for r in rows[2:]:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    if m:
        op[m.group(2)] += int(r[ix["Instructions Executed"]])
T, B, p, nwin = 512, 148, 256, 16
grid, block = 444, 256
PP = -(-T * B * p // (grid * block))
warp_pw = grid * block // 32 * PP * nwin                      # point-windows per launch, in warps
inst = sum(op.values()); wide = sum(v for k, v in op.items() if k.startswith("IMAD.WIDE"))
stamp = bench.kernel_src_hash()
how = "ncu --set full --import-source on --clock-control none -k regex:k_giant_step -s 3 -c 1 python bench.py --steps 1 --warmup 3 (scripts/r2_profile.sh); one launch timed alone by ncu"
json.dump({"kernel": "gs::k_giant_step<256,3,0>", "windows_per_launch": nwin, "workload": "configs[1]", "kernel_src_sha16": stamp,
           "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
           "giant_steps_per_launch": 2 * T * B * p * nwin, "dram_bytes_per_giant_step": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / (2 * T * B * p * nwin),
           "how": how}, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
json.dump({"kernel": "gs::k_giant_step<256,3,0>", "windows_per_launch": nwin, "workload": "configs[1]", "kernel_src_sha16": stamp,
           "warp_instructions_executed": inst, "inst_per_point_window": inst / warp_pw, "imad_wide_per_point_window": wide / warp_pw,
           "imad_wide_per_step": wide / warp_pw / 2, "other_per_point_window": (inst - wide) / warp_pw,
           "opcode_mix_per_point_window": {k: round(v / warp_pw, 2) for k, v in op.most_common(24)},
           "fmaheavy_active_pct": float(d["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
           "alu_active_pct": float(d["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
           "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
           "gpu_time_ms_alone": float(d["gpu__time_duration.sum"].replace(",", "")) / (1e6 if u.get("gpu__time_duration.sum") == "ns" else 1), "how": how},
          open(os.path.join(ROOT, "profiles", "r2_inst_mix.json"), "w"), indent=1)
# launch list
lf = os.path.join(ROOT, "gpurun_out", "launches_final.csv")
if os.path.exists(lf):
    rows = [r for r in csv.reader(open(lf)) if len(r) > 8]
    h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki][:90], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", f"r2_launches_{tag}.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1 --no-cpu --no-extras\n"
                "(table generation + 5 timed/warm-up launches + the end-to-end run + the whole puzzle #65 solve; launches serialised and cold under ncu)\n"
                "launches   total ms   share  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n:8d} {t/1e6:10.3f} {100*t/tot:6.1f}%  {k}\n")
print(open(os.path.join(ROOT, "profiles", "r2_inst_mix.json")).read())
