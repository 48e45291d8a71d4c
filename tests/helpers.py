"""Shared builders for the parity tests: synthetic tables with planted hits (TEST INFRASTRUCTURE)."""
import random
import numpy as np
from tests import pyec

M64 = 2**64 - 1


def random_point(rnd):
    return pyec.mul(pyec.G, rnd.randrange(1, pyec.N))


def giant(w, k):
    """Q_k = (k+1) * (-(2w)G)"""
    return pyec.neg(pyec.mul(pyec.G, (k + 1) * 2 * w))


def candidate_x(P, w, k, code):
    """X coordinate the kernel computes for record (code, k) and base point P."""
    Q = giant(w, k)
    if code == 1:
        return pyec.add(P, Q)[0]
    if code == 2:
        return pyec.add(P, pyec.neg(Q))[0]
    if code == 4:
        return pyec.add(P, P)[0]
    if code == 5:
        return P[0]
    raise ValueError(code)


def synthetic_keys(w, planted_x, seed):
    """w 64-bit table keys: uniformly random ones plus the low 64 bits of the planted X values."""
    rs = np.random.RandomState(seed)
    keys = rs.randint(0, 2**63, size=w, dtype=np.int64).astype(np.uint64) * np.uint64(2) + rs.randint(0, 2, size=w).astype(np.uint64)
    for i, x in enumerate(planted_x):
        keys[(i * 7919 + 13) % w] = np.uint64(x & M64)
    return keys


def norm_hits(hits):
    """sorted multiset of (code, idx) with the idx of code 5 forced to 0 (never written by the reference)."""
    return sorted((c, 0 if c == 5 else i) for c, i in hits)


def build_case(orc, T, B, p, w, htsz, seed, nplant=24):
    rnd = random.Random(seed)
    N = T * B * p
    Ppt = random_point(rnd)
    plant = []
    for i in range(nplant):
        k = rnd.randrange(N) if i > 3 else (0, N - 1, N // 2, 1 % N)[i]
        plant.append((1 + (i & 1), k))
    xs = [candidate_x(Ppt, w, k, c) for c, k in plant]
    keys = synthetic_keys(w, xs, seed)
    cpu, gpu = orc.build_ht(keys, htsz)
    g2 = orc.build_g2(T, B, p, w)
    return Ppt, plant, cpu, gpu, g2
