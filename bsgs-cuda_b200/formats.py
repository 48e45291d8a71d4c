"""The reference's on-disk artefacts, byte for byte.

  <Gxhex>_<w>_<H>_htCPUv0.BIN : u32 off[H] (exclusive prefix), u32 total, total x {u32 tag, u32 j}
                                 packHTFile, 1_9_7File.pb:3232-3335
  <Gxhex>_<w>_<H>_htGPUv0.BIN : same header, total x u32 tag       packHTGPUFile, 1_9_7File.pb:3337-3444
  <T>_<B>_<p>_<w>_g2.BIN       : N*32 bytes X block | N*32 bytes Y block, transposed word order
                                 Save_Load_Giants / Writeint, 1_9_7File.pb:1905-2058, 1831-1903
File names follow Save_HTpacked (1_9_7File.pb:3652-3655) and Save_Load_Giants (1_9_7File.pb:1916).
"""
from __future__ import annotations
import os
import numpy as np
from . import ec

GX_HEX = format(ec.G[0], "064x")


def htcpu_name(w, htsz):
    return f"{GX_HEX}_{w}_{1 << htsz}_htCPUv0.BIN"


def htgpu_name(w, htsz):
    return f"{GX_HEX}_{w}_{1 << htsz}_htGPUv0.BIN"


def g2_name(T, B, p, w):
    return f"{T}_{B}_{p}_{w}_g2.BIN"


def htcpu_bytes(w, htsz):
    return (1 << htsz) * 4 + 4 + w * 8


def htgpu_bytes(w, htsz):
    return (1 << htsz) * 4 + 4 + w * 4


def g2_bytes(T, B, p):
    return T * B * p * 64


def save(path, arr: np.ndarray):
    tmp = path + ".tmp"
    np.ascontiguousarray(arr).view(np.uint8).tofile(tmp)
    os.replace(tmp, path)


def load(path, expect_bytes=None) -> np.ndarray:
    a = np.fromfile(path, dtype=np.uint8)
    if expect_bytes is not None and a.nbytes != expect_bytes:
        raise ValueError(f"{path}: {a.nbytes} bytes, expected {expect_bytes}")
    return a


def ht_header(img: np.ndarray, htsz):
    """-> (off[H+1] as u32 view, payload view)"""
    H = 1 << htsz
    return img[:4 * (H + 1)].view(np.uint32), img[4 * (H + 1):]


def htcpu_lookup(htcpu: np.ndarray, htsz, key64):
    """All payloads j whose (bucket, tag) equal those of key64 (ReadHTpack + compare, 1_9_7File.pb:3028-3054)."""
    off, payload = ht_header(htcpu, htsz)
    items = payload.view(np.uint32).reshape(-1, 2)
    h = (key64 & 0xFFFFFFFF) & ((1 << htsz) - 1)
    tag = (key64 >> 32) & 0xFFFFFFFF
    lo, hi = int(off[h]), int(off[h + 1])
    seg = items[lo:hi]
    return [int(j) for t, j in seg if int(t) == tag]


def g2_word_index(T, B, p, k, word):
    """u32 index of `word` (0 = most significant) of giant k inside the X (or Y) block."""
    TB = T * B
    blk, thr, r = k // (T * p), (k % (T * p)) // p, k % p
    return r * 8 * TB + (blk * T + thr) + word * TB


def g2_get(g2: np.ndarray, T, B, p, k):
    N = T * B * p
    X = g2[:N * 32].view(np.uint32)
    Y = g2[N * 32:].view(np.uint32)
    x = y = 0
    for i in range(8):
        x = (x << 32) | int(X[g2_word_index(T, B, p, k, i)])
        y = (y << 32) | int(Y[g2_word_index(T, B, p, k, i)])
    return (x, y)
