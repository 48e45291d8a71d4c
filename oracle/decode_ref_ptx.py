#!/usr/bin/env python3
"""Decode the reference's embedded CUDA kernel into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference ships its kernel `_test1` as PTX text, XOR-ed with 0x5D and stored as `Data.q`
quadwords between the labels BSGS4_cuda_quad_htchangeble_v2 / ...v2end of 1_9_7File.pb
(1_9_7File.pb:5181-23979); the program undoes the XOR at start-up (1_9_7File.pb:4739-4741) and
hands the text to cuModuleLoadData (1_9_7File.pb:2194).  This script performs the same decode at
BUILD time, reading the reference where it lies, and writes the PTX only under oracle/_ref/
(git-ignored; never committed).  oracle/Makefile then assembles it with ptxas for sm_100a.
"""
import struct, sys, os

def decode(pb_path: str, label: str = "BSGS4_cuda_quad_htchangeble_v2") -> str:
    lines = open(pb_path, "r", errors="replace").read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == label + ":")
    end = next(i for i, l in enumerate(lines) if l.strip() == label + "end:")
    raw = bytearray()
    for l in lines[start + 1:end]:
        l = l.strip()
        if not l.startswith("Data.q"):
            continue
        for tok in l[6:].split(","):
            tok = tok.strip()
            if tok.startswith("$"):
                raw += struct.pack("<Q", int(tok[1:], 16))
    dec = bytes(b ^ 0x5D for b in raw)          # PokeC(addr+i, PeekC(addr+i) ! 93)
    dec = dec[:dec.rfind(b"}") + 1]              # drop the "\0]]]]" padding tail
    return dec.decode("latin1").replace("\r", "") + "\n"

if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/1_9_7File.pb"
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ref_test1.ptx")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    txt = decode(src)
    assert ".entry _test1" in txt and "_A[30]" in txt, "decoded text does not look like the kernel"
    open(out, "w").write(txt)
    print(f"decoded {len(txt)} bytes of PTX -> {out}")
