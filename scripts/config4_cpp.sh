# BASELINE.json configs[3] through the C++ host: 16 synthetic targets over one 2^64 range, tables built once
set -e
work=$(mktemp -d); cd "$work"
python - <<'PY'
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from bsgs_cuda_b200 import ec
PK = 0x49dccfd96dc5df56487436f5a1b18c4f5d34f65ddb48cb5e0000000000000000
M64 = (1 << 64) - 1
def splitmix64(state):
    while True:
        state = (state + 0x9E3779B97F4A7C15) & M64
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        yield z ^ (z >> 31)
g = splitmix64(20261017)
secrets = [PK + next(g) for _ in range(16)]
open("pubs.txt", "w").write("".join(ec.compress(ec.mul(ec.G, s)) + "\n" for s in secrets))
open("expect.txt", "w").write("".join(f"KEY[{i}]: 0x{s:064x}\n" for i, s in enumerate(secrets, 1)))
PY
exe=${GRAFT_REPO_ROOT:-/root/repo}/bsgs-cuda_b200/bsgs_b200
start=$(date +%s.%N)
$exe -t 512 -b 148 -p 256 -w 30 -htsz 28 -pk 49dccfd96dc5df56487436f5a1b18c4f5d34f65ddb48cb5e0000000000000000 -pke 49dccfd96dc5df56487436f5a1b18c4f5d34f65ddb48cb5effffffffffffffff -infile pubs.txt -nosave > out.txt
end=$(date +%s.%N)
grep -c "^KEY\[" out.txt; grep "^KEY\[" out.txt | diff - expect.txt && echo "all 16 keys correct"
grep "Tables ready" out.txt; tail -1 out.txt
python -c "print('wall_s', round($end - $start, 3))"
