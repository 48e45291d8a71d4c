"""ctypes front-end of the CPU oracle and of the reference-kernel harness.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (bsgs-cuda_b200/) never imports this.
"""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_LIB = os.path.join(REF_DIR, "libref_harness.so")
REF_CUBIN = os.path.join(REF_DIR, "ref_test1.cubin")

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8


def build(force: bool = False) -> None:
    """Compile the checker (liboracle.so, and oracle/_ref when the reference tree is present)."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "bsgs_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    try:
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        pass


class Fe(C.Structure):
    _fields_ = [("v", C.c_uint64 * 4)]

    @staticmethod
    def of(x: int) -> "Fe":
        f = Fe()
        for i in range(4):
            f.v[i] = (x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF
        return f

    def int(self) -> int:
        return sum(int(self.v[i]) << (64 * i) for i in range(4))


class Pt(C.Structure):
    _fields_ = [("x", Fe), ("y", Fe)]

    @staticmethod
    def of(x: int, y: int) -> "Pt":
        p = Pt()
        p.x = Fe.of(x)
        p.y = Fe.of(y)
        return p

    def tup(self):
        return (self.x.int(), self.y.int())


class Rec96(C.Structure):
    _fields_ = [("x", Fe), ("y", Fe), ("chain", Fe)]


class Hit(C.Structure):
    _fields_ = [("code", C.c_uint32), ("idx", C.c_uint32)]


class Search(C.Structure):
    _fields_ = [("T", C.c_uint32), ("B", C.c_uint32), ("p", C.c_uint32), ("htsz", C.c_uint32), ("w", C.c_uint64),
                ("pk", Fe), ("target", Pt), ("findpub", Pt), ("addpubg", Pt), ("center", Pt), ("pubaddbig", Pt),
                ("prkaddbig", Fe), ("centerbig", Fe), ("globkey", Fe), ("globpub", Pt)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        pf, pp = C.POINTER(Fe), C.POINTER(Pt)
        for name in ("orc_fe_add", "orc_fe_sub", "orc_fe_mul"):
            getattr(L, name).argtypes = [pf, pf, pf]
        for name in ("orc_fe_neg", "orc_fe_sqr", "orc_fe_inv"):
            getattr(L, name).argtypes = [pf, pf]
        for name in ("orc_mod_add", "orc_mod_sub"):
            getattr(L, name).argtypes = [pf, pf, pf, pf]
        L.orc_mod_inv.argtypes = [pf, pf, pf]
        L.orc_raw_add.argtypes = [pf, pf, pf]; L.orc_raw_add.restype = C.c_int
        L.orc_raw_sub.argtypes = [pf, pf, pf]; L.orc_raw_sub.restype = C.c_int
        L.orc_pt_dbl.argtypes = [pp, pp]
        L.orc_pt_add.argtypes = [pp, pp, pp]
        L.orc_pt_mul.argtypes = [pp, pp, pf]
        L.orc_y_from_x.argtypes = [pf, pf]
        L.orc_begin_batch_add.argtypes = [pf, C.c_size_t, pp, C.POINTER(Rec96)]
        L.orc_complete_batch_add_with_double.argtypes = [pp, C.c_size_t, pp, C.POINTER(Rec96), pf]
        L.orc_fillarray_n.argtypes = [C.POINTER(Rec96), C.c_size_t, pp]
        L.orc_baby_keys.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
        L.orc_htcpu_bytes.argtypes = [C.c_uint64, C.c_uint32]; L.orc_htcpu_bytes.restype = C.c_size_t
        L.orc_htgpu_bytes.argtypes = [C.c_uint64, C.c_uint32]; L.orc_htgpu_bytes.restype = C.c_size_t
        L.orc_build_ht.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_build_g2.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]
        L.orc_g2_get.argtypes = [pp, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_probe_htgpu.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64]; L.orc_probe_htgpu.restype = C.c_int
        L.orc_lookup_htcpu.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32]
        L.orc_lookup_htcpu.restype = C.c_uint32
        L.orc_launch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, pp,
                                 C.POINTER(Hit), C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, C.c_uint64]
        L.orc_launch.restype = C.c_uint32
        L.orc_search_init.argtypes = [C.POINTER(Search), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, pf, pp, pf]
        L.orc_get_job.argtypes = [C.POINTER(Search), pf, pp]
        L.orc_check_hit.argtypes = [C.POINTER(Search), C.c_void_p, C.POINTER(Hit), pf, pp, pf]
        L.orc_check_hit.restype = C.c_int
        L.orc_parse_w.argtypes = [C.c_char_p]; L.orc_parse_w.restype = C.c_uint64
        L.orc_time_seq_adds.argtypes = [C.c_uint32, C.POINTER(Pt)]; L.orc_time_seq_adds.restype = C.c_double
        _lib = L
    return _lib


# ------------------------------------------------------------------ thin pythonic helpers
def fe2(fn, a: int, b: int) -> int:
    r = Fe(); fa, fb = Fe.of(a), Fe.of(b)
    getattr(lib(), fn)(C.byref(r), C.byref(fa), C.byref(fb))
    return r.int()


def fe1(fn, a: int) -> int:
    r = Fe(); fa = Fe.of(a)
    getattr(lib(), fn)(C.byref(r), C.byref(fa))
    return r.int()


def pt_add(a, b):
    r = Pt(); pa, pb = Pt.of(*a), Pt.of(*b)
    lib().orc_pt_add(C.byref(r), C.byref(pa), C.byref(pb)); return r.tup()


def pt_dbl(a):
    r = Pt(); pa = Pt.of(*a)
    lib().orc_pt_dbl(C.byref(r), C.byref(pa)); return r.tup()


def pt_mul(a, k: int):
    r = Pt(); pa = Pt.of(*a); fk = Fe.of(k)
    lib().orc_pt_mul(C.byref(r), C.byref(pa), C.byref(fk)); return r.tup()


def pt_neg(a):
    return (a[0], (P - a[1]) % P)


def y_from_x(x: int) -> int:
    return fe1("orc_y_from_x", x)


def parse_pubkey(s: str):
    """Pubkey forms accepted by the reference: 128 hex, '04'+128, '02/03'+64 (1_9_7File.pb:5006-5018)."""
    s = s.strip().lower()
    if s.startswith("0x"):
        s = s[2:]
    if len(s) == 130 and s.startswith("04"):
        s = s[2:]
    if len(s) == 128:
        return (int(s[:64], 16), int(s[64:], 16))
    if len(s) == 66 and s[:2] in ("02", "03"):
        x = int(s[2:], 16)
        y = y_from_x(x)
        if (y & 1) != (int(s[:2], 16) & 1):      # parity fix, 1_9_7File.pb:287-289
            y = P - y
        return (x, y)
    raise ValueError("bad pubkey")


def parse_w(arg: str) -> int:
    return int(lib().orc_parse_w(arg.encode()))


def time_seq_adds(n: int = 10000):
    """Curve64.pb:3303-3320: n sequential affine additions b = G + b on one core -> (seconds, final point)."""
    out = Pt()
    dt = lib().orc_time_seq_adds(n, C.byref(out))
    return dt, out.tup()


def baby_keys(w: int, nthreads: int = 8) -> np.ndarray:
    keys = np.empty(w, dtype=np.uint64)
    lib().orc_baby_keys(keys.ctypes.data, w, nthreads)
    return keys


def build_ht(keys: np.ndarray, htsz: int):
    w = len(keys)
    cpu = np.empty(lib().orc_htcpu_bytes(w, htsz), dtype=np.uint8)
    gpu = np.empty(lib().orc_htgpu_bytes(w, htsz), dtype=np.uint8)
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    lib().orc_build_ht(keys.ctypes.data, w, htsz, cpu.ctypes.data, gpu.ctypes.data)
    return cpu, gpu


def build_g2(T: int, B: int, p: int, w: int, nthreads: int = 8) -> np.ndarray:
    out = np.empty(T * B * p * 64, dtype=np.uint8)
    lib().orc_build_g2(out.ctypes.data, T, B, p, w, nthreads)
    return out


def g2_get(g2: np.ndarray, T: int, B: int, p: int, k: int):
    q = Pt(); lib().orc_g2_get(C.byref(q), g2.ctypes.data, T, B, p, k); return q.tup()


def probe_htgpu(htgpu: np.ndarray, htsz: int, key64: int) -> bool:
    return bool(lib().orc_probe_htgpu(htgpu.ctypes.data, htsz, key64))


def lookup_htcpu(htcpu: np.ndarray, htsz: int, key64: int):
    js = (C.c_uint32 * 64)()
    n = lib().orc_lookup_htcpu(htcpu.ctypes.data, htsz, key64, js, 64)
    return [int(js[i]) for i in range(n)]


def launch(g2, T, B, p, htgpu, htsz, P_xy, batch=0, nthreads=8, cap=4096, k_begin=0, k_end=None):
    """One reference launch on the CPU -> sorted list of (code, idx)."""
    out = (Hit * cap)(); pt = Pt.of(*P_xy)
    if k_end is None:
        k_end = T * B * p
    n = lib().orc_launch(g2.ctypes.data, T, B, p, htgpu.ctypes.data, htsz, C.byref(pt), out, cap, batch, nthreads, k_begin, k_end)
    assert n <= cap, "oracle record buffer too small"
    return sorted((int(out[i].code), int(out[i].idx)) for i in range(n))


class SearchState:
    """Window enumeration + hit decoding restated from 1_9_7File.pb (GetJob, checkerThread)."""

    def __init__(self, T, B, p, w, htsz, pk: int, target_xy, start_key: int | None = None):
        self.s = Search(); fpk = Fe.of(pk); tgt = Pt.of(*target_xy)
        sk = Fe.of(start_key) if start_key is not None else None
        lib().orc_search_init(C.byref(self.s), T, B, p, w, htsz, C.byref(fpk), C.byref(tgt), C.byref(sk) if sk else None)

    def get_job(self):
        k = Fe(); P_ = Pt()
        lib().orc_get_job(C.byref(self.s), C.byref(k), C.byref(P_))
        return k.int(), P_.tup()

    def check_hit(self, htcpu: np.ndarray, code: int, idx: int, window_key: int, P_xy):
        h = Hit(code, idx); k = Fe.of(window_key); pt = Pt.of(*P_xy); out = Fe()
        ok = lib().orc_check_hit(C.byref(self.s), htcpu.ctypes.data, C.byref(h), C.byref(k), C.byref(pt), C.byref(out))
        return out.int() if ok else None


# ------------------------------------------------------------------ reference-kernel harness (GPU box only)
class RefKernel:
    """The reference's own `_test1` cubin driven through oracle/ref_harness.c."""

    def __init__(self, T, B, p, w, htsz, dev=0, cubin=None):
        """cubin=None: the reference's kernel; otherwise any module with the same interface (entry
        `_test1`, const `_A`, the reference memory map), e.g. bsgs-cuda_b200/dropin_test1.cubin."""
        cubin = cubin or REF_CUBIN
        if not (os.path.exists(REF_LIB) and os.path.exists(cubin)):
            raise FileNotFoundError(f"{REF_LIB} / {cubin} not built (the reference cubin needs the reference tree at build time)")
        self.L = C.CDLL(REF_LIB)
        self.L.ref_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32]
        self.L.ref_upload_giants.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self.L.ref_upload_ht.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        self.L.ref_launch.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        self.L.ref_destroy.argtypes = [C.c_void_p]
        self.L.ref_upload_ht_dev.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t]
        self.L.ref_time_launches.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_float)]
        self.h = C.c_void_p()
        rc = self.L.ref_create(C.byref(self.h), dev, cubin.encode(), T, B, p, w, htsz)
        if rc:
            raise RuntimeError(f"ref_create failed: {rc}")

    def upload(self, g2: np.ndarray, htgpu: np.ndarray):
        rc = self.L.ref_upload_giants(self.h, g2.ctypes.data, g2.nbytes)
        assert rc == 0, rc
        rc = self.L.ref_upload_ht(self.h, htgpu.ctypes.data, htgpu.nbytes)
        assert rc == 0, rc

    def upload_giants(self, g2: np.ndarray):
        rc = self.L.ref_upload_giants(self.h, g2.ctypes.data, g2.nbytes)
        assert rc == 0, rc

    def upload_ht_dev(self, dev_ptr: int, nbytes: int):
        rc = self.L.ref_upload_ht_dev(self.h, dev_ptr, nbytes)
        assert rc == 0, rc

    def time_launches(self, P_xy, iters: int) -> float:
        ms = C.c_float()
        rc = self.L.ref_time_launches(self.h, P_xy[0].to_bytes(32, "little"), P_xy[1].to_bytes(32, "little"), iters, C.byref(ms))
        assert rc == 0, rc
        return ms.value

    def launch(self, P_xy, cap=240):
        recs = (C.c_uint32 * (2 * cap))(); n = C.c_uint32(); ms = C.c_float()
        px = P_xy[0].to_bytes(32, "little"); py = P_xy[1].to_bytes(32, "little")
        rc = self.L.ref_launch(self.h, px, py, recs, cap, C.byref(n), C.byref(ms))
        assert rc == 0, rc
        m = min(n.value, cap, 240)
        out = []
        for i in range(m):
            code, idx = int(recs[2 * i]), int(recs[2 * i + 1])
            out.append((code, 0 if code == 5 else idx))       # idx of code 5 is never written (PTX197:L104-109)
        return sorted(out), n.value, ms.value

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
