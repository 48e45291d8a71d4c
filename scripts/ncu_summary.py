#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): pipes, stalls, DRAM bytes, opcode mix, top stall sites.
usage: scripts/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/name.txt]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
for k in keys:
    if k in d:
        print(f"{k:75s} {d[k]:>18s} {u.get(k,'')}")
print("\nstall reasons (warp-cycles per issued instruction):")
st = {k: float(d[k]) for k in hdr if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k)}
for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
    print(f"  {k.split('stalled_')[1].split('_per_issue')[0]:22s} {v:7.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]; ix = {x: i for i, x in enumerate(h)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
op = collections.Counter(); smp = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    if m:
        op[m.group(2)] += int(r[ix["Instructions Executed"]]); smp[m.group(2)] += int(r[ix["# Samples"]])
te = sum(op.values())
print(f"\nexecuted warp instructions: {te}   static instructions: {len(data)}   stall samples: {tot}")
for k, v in op.most_common(16):
    print(f"  {k:22s} {v:14d} {100*v/te:5.1f}%   samples {100*smp[k]/max(tot,1):5.1f}%")
stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
print("\ntop stall sites:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:14]:
    why = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) * 5 > int(r[ix["# Samples"]])}
    print(f"  {100*int(r[ix['# Samples']])/max(tot,1):5.2f}%  {r[ix['Source']].strip()[:64]:64s} {why}")
