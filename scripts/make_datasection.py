#!/usr/bin/env python3
"""Write a PureBasic DataSection holding a PTX module in the reference's own embedding, so that
bsgs-cuda_b200/dropin_test1.ptx can replace the kernel text of 1_9_7File.pb:5181-23979 in place:

    every byte XOR 93 (undone at start-up by `PokeC(addr+i, PeekC(addr+i) ! 93)`, 1_9_7File.pb:4739-4741),
    packed little-endian into `Data.q $....` quadwords between the labels
    `BSGS4_cuda_quad_htchangeble_v2:` and `BSGS4_cuda_quad_htchangeble_v2end:`; the text is padded with
    XOR-ed zero bytes to a multiple of 8 (the reference's own tail is `\\0` + padding), which
    cuModuleLoadData treats as the end of the string.

usage: scripts/make_datasection.py bsgs-cuda_b200/dropin_test1.ptx out.pbi [--check]
(--check decodes the output again with the rule oracle/decode_ref_ptx.py uses and compares.)"""
import sys


def encode(ptx: bytes, per_line: int = 8) -> str:
    data = ptx + b"\0"
    data += b"\0" * (-len(data) % 8)
    x = bytes(b ^ 93 for b in data)
    quads = [int.from_bytes(x[i:i + 8], "little") for i in range(0, len(x), 8)]
    lines = ["DataSection", "  BSGS4_cuda_quad_htchangeble_v2:"]
    for i in range(0, len(quads), per_line):
        lines.append("  Data.q " + ",".join("$%016X" % q for q in quads[i:i + per_line]))
    lines += ["  BSGS4_cuda_quad_htchangeble_v2end:", "EndDataSection", ""]
    return "\n".join(lines)


def decode(text: str) -> bytes:
    out = bytearray()
    on = False
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("BSGS4_cuda_quad_htchangeble_v2end:"):
            break
        if s.startswith("BSGS4_cuda_quad_htchangeble_v2:"):
            on = True
            continue
        if on and s.startswith("Data.q"):
            for tok in s[6:].split(","):
                out += int(tok.strip().lstrip("$"), 16).to_bytes(8, "little")
    return bytes(b ^ 93 for b in out)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    ptx = open(src, "rb").read()
    text = encode(ptx)
    open(dst, "w").write(text)
    if "--check" in sys.argv:
        back = decode(text)
        assert back.rstrip(b"\0") == ptx.rstrip(b"\0"), "round trip failed"
    print(f"{dst}: {len(ptx)} bytes of PTX -> {text.count('Data.q')} Data.q lines")


if __name__ == "__main__":
    main()
