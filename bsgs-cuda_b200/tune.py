"""Parameter suggestion (`Tune`, 1_9_7File.pb:324-431) and the per-architecture core table it is fed
from (1_9_7File.pb:816-846).

Two functions:

* `tune_reference(memsize, sm, cores)` restates the reference procedure line by line and returns the
  values of its `Try -t .. -b .. -p .. -w .. -htsz ..` line.  The reference's core table stops at
  compute capability 8; `cores_for()` continues it (9.x and 10.x: 128 FP32 lanes per SM) so that the
  procedure has an input on Hopper/Blackwell at all.
* `tune_b200(free_bytes, sm)` is the suggestion this build prints: the reference procedure knows
  neither the kernel's structural limits (8*N < 2^31 in the reference PTX, N < 2^31 here, w < 2^32,
  htsz <= 31) nor this library's memory model (giants 64 B + 2 x 32 B of chain per window per giant (two launches in flight);
  table = htGPU image + one 32-byte fingerprint sector per bucket + 4 B payload per entry when the
  table is generated on the device), and on a 180 GB part it proposes an N the reference kernel
  cannot index.
"""
import math

ALIGN_SIZE = 128          # #align_size, 1_9_7File.pb:113
HELPSIZE = 4096           # #helpsize, 1_9_7File.pb:112
MB = 1048576


def cores_for(major: int, minor: int, mp: int) -> int:
    """`cores` of the device listing (1_9_7File.pb:816-846), continued past compute capability 8."""
    if major == 2:
        return mp * (48 if minor == 1 else 32)
    if major == 3:
        return mp * 192
    if major == 5:
        return mp * 128
    if major in (6, 7):
        return mp * 64
    if major >= 8:                      # 8: Ampere/Ada (reference); 9 Hopper, 10 Blackwell: 128 FP32 lanes per SM
        return mp * 128
    return 0


def _pb_int(x: float) -> int:
    """PureBasic assigns a float expression to an integer variable by rounding to nearest."""
    return int(math.floor(x + 0.5))


def tune_reference(memsize: int, sm: int, cores: int, ncpu: int = 1):
    """1_9_7File.pb:324-431.  Returns None where the reference prints nothing (memsize <= gsize)."""
    blockN = sm * 2
    treadN = 2 * cores // sm
    pparamN = 140
    gsize = treadN * blockN * pparamN * 96 + 160 + ALIGN_SIZE
    if memsize - gsize <= 0:
        return None
    k = _pb_int(math.log(memsize - gsize) / math.log(2))
    if k > 31:
        k = 31
    k += 1
    htszN = 0
    while True:                                                   # Repeat ... Until, 1_9_7File.pb:341-358
        k -= 1
        htszN = {31: 29, 30: 27, 29: 27, 28: 27, 27: 25}.get(k, htszN)
        httotsize = (1 << k) * 4 + (1 << htszN) * 4
        if gsize + httotsize < memsize or k < 25:
            break
    while True:                                                   # 1_9_7File.pb:359-378
        frmem = memsize - httotsize
        frmem = (frmem - 160 - ALIGN_SIZE) // 96
        frmem //= blockN
        frmem //= treadN
        if frmem > 512:
            if treadN < 512:
                treadN += 32
            else:
                blockN += sm
        if frmem < 512 or blockN // sm >= 4:
            break
    frmem = memsize - httotsize
    frmem = (frmem - 160 - ALIGN_SIZE - 96) // 96
    frmem = frmem // blockN // treadN
    if frmem & 1:
        frmem += 1
    pparamN = frmem
    gsize = treadN * blockN * pparamN * 96 + 160 + ALIGN_SIZE
    if gsize + httotsize > memsize:
        pparamN -= 2
        gsize = treadN * blockN * pparamN * 96 + 160 + ALIGN_SIZE
    # host RAM needed while generating (1_9_7File.pb:394-422): the largest of the three phases
    r1 = HELPSIZE * 96 * ncpu + 8 * (1 << k) + ncpu * HELPSIZE * 8 + (1 << htszN) * 8 + ALIGN_SIZE + (1 << htszN) * 8
    r2 = HELPSIZE * 96 * ncpu + 8 * (1 << k) + (1 << htszN) * 8 + ALIGN_SIZE + (1 << htszN) * 8 + (1 << htszN) * 4 + ALIGN_SIZE + (1 << k) * 8 + 4
    r3 = HELPSIZE * 96 * ncpu + (1 << htszN) * 4 + ALIGN_SIZE + (1 << k) * 4 + 4 + 2 * (treadN * blockN * pparamN + 1) * 64
    return {"t": treadN, "b": blockN, "p": pparamN, "w": k, "htsz": htszN,
            "device_mb": (gsize + httotsize) / MB, "gen_ram_mb": max(r1, r2, r3) // MB}


def table_bytes(w: int, htsz: int, on_device_build: bool = True) -> int:
    """Device bytes of the baby-step table in this library."""
    H = 1 << htsz
    return 4 * (H + 1) + 4 * w + 32 * H + (4 * w if on_device_build else 0)


def tune_b200(free_bytes: int, sm: int, windows: int = 16, t: int = 512):
    """Suggestion for this library: one CTA column per SM (`-b sm`), the largest table that leaves
    room for a useful giant array, bucket load between 4 and 8, N as large as the remaining memory
    and the structural limit allow (the per-launch work and the inversion amortisation grow with N)."""
    best = None
    for wexp2 in (63, 62, 61, 60, 59, 58, 57, 56, 54, 52, 50, 48):          # -w 31.5, 31, 30.5, ... (half steps, then whole)
        wexp = wexp2 / 2
        w = int(2 ** wexp)
        if w >= 1 << 32:
            continue
        htsz = min(31, int(math.floor(wexp)) - 2)
        tb = table_bytes(w, htsz)
        if tb > 0.5 * free_bytes:
            continue
        room = int((free_bytes - tb) * 0.9)
        n_max = min(room // (64 + 2 * 32 * windows), (1 << 28) - 1)                 # 8*N < 2^31: stay loadable by the reference's own files
        p = (n_max // (t * sm)) & ~1
        if p < 2:
            continue
        best = {"t": t, "b": sm, "p": p, "w": wexp, "w_items": w, "htsz": htsz, "windows_per_launch": windows,
                "device_mb": (tb + t * sm * p * (64 + 2 * 32 * windows)) / MB}
        break
    return best


def format_try(s) -> str:
    """The reference's line (1_9_7File.pb:425)."""
    w = s["w"]
    wtxt = str(int(w)) if float(w).is_integer() else str(w)
    line = f"Try -t {s['t']} -b {s['b']} -p {s['p']} -w {wtxt} -htsz {s['htsz']} [{s['device_mb']:.3f} MB]"
    if "gen_ram_mb" in s:
        line += f" Gen RAM[{s['gen_ram_mb']} MB]"
    return line
