#!/usr/bin/env python3
"""BASELINE.json configs[3] / SURVEY 8d config 4: -infile with 16 synthetic pubkeys over one 2^64 range.
Keys pk + r_i, r_i = first 16 outputs of splitmix64(seed 20261017) (a fixed-seed 64-bit PRNG), range start
0x49dccfd96dc5df56487436f5a1b18c4f5d34f65ddb48cb5e0000000000000000 (README.md:47 of the reference);
one compressed pubkey per line; processed sequentially through the command line with the tables of
config 2 (-t 512 -b 148 -p 256 -w 30 -htsz 28) built once.  Prints per-target time and a JSON summary."""
import io, json, os, re, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsgs_cuda_b200 as b
from bsgs_cuda_b200 import ec, cli

PK = 0x49dccfd96dc5df56487436f5a1b18c4f5d34f65ddb48cb5e0000000000000000
M64 = (1 << 64) - 1


def splitmix64(state):
    while True:
        state = (state + 0x9E3779B97F4A7C15) & M64
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        yield z ^ (z >> 31)


gen = splitmix64(20261017)
secrets = [PK + next(gen) for _ in range(16)]
work = tempfile.mkdtemp(prefix="cfg4_")
os.chdir(work)
with open("pubs.txt", "w") as f:
    for s in secrets:
        f.write(ec.compress(ec.mul(ec.G, s)) + "\n")


class Tee(io.StringIO):
    def __init__(self): super().__init__(); self.stamps = []
    def write(self, s):
        if s.startswith("KEY[") or s.startswith("Findpubkey") or s.startswith("Tables ready"):
            self.stamps.append((time.time(), s.strip()))
        return super().write(s)


out = Tee()
t0 = time.time()
rc = cli.main(f"-t 512 -b 148 -p 256 -w 30 -htsz 28 -pk {PK:x} -pke {PK + M64:x} -infile pubs.txt -nosave".split(), out=out)
total = time.time() - t0
text = out.getvalue()
found = [int(m, 16) for m in re.findall(r"KEY\[\d+\]: 0x([0-9a-f]{64})", text)]
per = []
start = None
for t, s in out.stamps:
    if s.startswith("Findpubkey"):
        start = t
    elif s.startswith("KEY[") and start is not None:
        per.append(round(t - start, 3)); start = None
tables = next((s for _, s in out.stamps if s.startswith("Tables ready")), "")
N, w = 512 * 148 * 256, 1 << 30
windows = [((s - PK - 1) // (4 * N * w)) + 1 for s in secrets]
steps = sum(windows) * 2 * N
print(json.dumps({"rc": rc, "targets": 16, "all_found": found == secrets, "per_target_s": per, "search_s_total": sum(per), "wall_s": total,
                  "tables": tables, "windows_per_target": windows, "aggregate_giant_steps_per_s_lower_bound": steps / max(sum(per), 1e-9)}))
