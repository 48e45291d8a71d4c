"""ctypes binding of libbsgs_b200.so (C ABI declared in include/bsgs_b200.h)."""
from __future__ import annotations
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_WINDOWS = 16


def lib_path() -> str:
    """The in-tree CUDA library (BSGS_B200_LIB overrides the path: used to A/B experimental builds)."""
    return os.environ.get("BSGS_B200_LIB") or os.path.join(_HERE, "libbsgs_b200.so")


class BsgsError(RuntimeError):
    def __init__(self, status, what, cuda=0):
        super().__init__(f"{what}: status {status} ({_strerror(status)}), cuda error {cuda}")
        self.status = status
        self.cuda = cuda


class Hit(C.Structure):
    _fields_ = [("code", C.c_uint32), ("idx", C.c_uint32), ("window", C.c_uint32), ("reserved", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("grid", C.c_uint32), ("block", C.c_uint32), ("points_per_thread", C.c_uint32), ("sm_count", C.c_uint32),
                ("n_giants", C.c_uint64), ("bytes_giants", C.c_uint64), ("bytes_chain", C.c_uint64), ("bytes_table", C.c_uint64)]


class DevInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("cc_major", C.c_int32), ("cc_minor", C.c_int32), ("sm_count", C.c_int32), ("reserved", C.c_int32),
                ("free_bytes", C.c_uint64), ("total_bytes", C.c_uint64)]


EXPORTS = ["bsgs_create", "bsgs_destroy", "bsgs_strerror", "bsgs_last_cuda_error", "bsgs_upload_giants", "bsgs_upload_ht", "bsgs_upload_htcpu",
           "bsgs_giants_staging", "bsgs_giants_commit", "bsgs_ht_staging", "bsgs_ht_commit", "bsgs_step", "bsgs_step_multi",
           "bsgs_giants_generate", "bsgs_giants_export", "bsgs_ht_build_from_keys", "bsgs_ht_generate", "bsgs_ht_build_synthetic", "bsgs_ht_export", "bsgs_ht_device_image",
           "bsgs_run", "bsgs_time_launches", "bsgs_probe_bench", "bsgs_probe_keys", "bsgs_get_info", "bsgs_set_option", "bsgs_debug_field", "bsgs_debug_key", "bsgs_debug_trace", "bsgs_device_count", "bsgs_device_query", "bsgs_bcast_ht"]

_lib = None


def load():
    """Load the CUDA library.  Fails loudly when it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C bsgs-cuda_b200` (or __graft_entry__.build()); "
                          "the giant-step path has no CPU fallback")
    L = C.CDLL(path)
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.bsgs_create.argtypes = [C.POINTER(vp), C.c_int, u32, u32, u32, u64, u32]
    L.bsgs_destroy.argtypes = [vp]; L.bsgs_destroy.restype = None
    L.bsgs_strerror.argtypes = [C.c_int]; L.bsgs_strerror.restype = C.c_char_p
    L.bsgs_last_cuda_error.argtypes = [vp]
    L.bsgs_upload_giants.argtypes = [vp, vp, C.c_size_t]
    L.bsgs_upload_ht.argtypes = [vp, vp, C.c_size_t]
    L.bsgs_upload_htcpu.argtypes = [vp, vp, C.c_size_t]
    L.bsgs_giants_staging.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.bsgs_giants_commit.argtypes = [vp]
    L.bsgs_ht_staging.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.bsgs_ht_commit.argtypes = [vp]
    L.bsgs_giants_generate.argtypes = [vp]
    L.bsgs_giants_export.argtypes = [vp, vp, C.c_size_t]
    L.bsgs_ht_build_from_keys.argtypes = [vp, vp, u64]
    L.bsgs_ht_generate.argtypes = [vp]
    L.bsgs_ht_build_synthetic.argtypes = [vp, u64]
    L.bsgs_ht_export.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.bsgs_ht_device_image.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.bsgs_step.argtypes = [vp, C.c_char_p, C.c_char_p, C.POINTER(Hit), u32, C.POINTER(u32)]
    L.bsgs_step_multi.argtypes = [vp, u32, C.c_char_p, C.POINTER(Hit), u32, C.POINTER(u32)]
    L.bsgs_run.argtypes = [vp, C.c_char_p, C.c_char_p, u64, u32, u32, C.POINTER(Hit), u32, C.POINTER(u32), C.POINTER(u64)]
    L.bsgs_time_launches.argtypes = [vp, u32, C.c_char_p, u32, C.POINTER(C.c_float)]
    L.bsgs_probe_bench.argtypes = [vp, u64, u64, u32, C.POINTER(C.c_float), C.POINTER(u64)]
    L.bsgs_probe_keys.argtypes = [vp, vp, u64, vp]
    L.bsgs_get_info.argtypes = [vp, C.POINTER(Info)]
    L.bsgs_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.bsgs_debug_field.argtypes = [C.c_int, C.c_int, vp, vp, vp, u32]
    L.bsgs_debug_key.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, vp, u32, C.c_int]
    L.bsgs_bcast_ht.argtypes = [C.POINTER(vp), C.c_int]
    L.bsgs_debug_trace.argtypes = [vp, vp, u32, C.POINTER(u32)]
    L.bsgs_device_count.argtypes = [C.POINTER(C.c_int)]
    L.bsgs_device_query.argtypes = [C.c_int, C.POINTER(DevInfo)]
    _lib = L
    return L


def devices():
    """The device listing of the reference's start-up (1_9_7File.pb:778-860): name, compute
    capability, multiprocessors, free/total memory of every CUDA device."""
    L = load()
    n = C.c_int()
    rc = L.bsgs_device_count(C.byref(n))
    if rc:
        raise BsgsError(rc, "bsgs_device_count")
    out = []
    for i in range(n.value):
        d = DevInfo()
        rc = L.bsgs_device_query(i, C.byref(d))
        if rc:
            raise BsgsError(rc, "bsgs_device_query")
        out.append({"index": i, "name": d.name.decode(), "cc": (d.cc_major, d.cc_minor), "sm_count": d.sm_count,
                    "free_bytes": d.free_bytes, "total_bytes": d.total_bytes})
    return out


def bcast_ht(engines):
    """Copy the table of engines[0] to the others device to device (bsgs_bcast_ht)."""
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    rc = load().bsgs_bcast_ht(arr, len(engines))
    if rc:
        raise BsgsError(rc, "bsgs_bcast_ht")


def debug_field(op: int, a, b, device=0):
    """Run one fp.cuh operation element-wise on the device (diagnostics / parity tests)."""
    n = len(a)
    A = np.array([[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for x in a], dtype=np.uint32)
    Bm = np.array([[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)] for x in b], dtype=np.uint32)
    R = np.zeros((n, 8), dtype=np.uint32)
    rc = load().bsgs_debug_field(device, op, A.ctypes.data, Bm.ctypes.data, R.ctypes.data, n)
    if rc:
        raise BsgsError(rc, "bsgs_debug_field")
    return [sum(int(R[j, i]) << (32 * i) for i in range(8)) for j in range(n)]


def _limbs(vals):
    a = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in vals), dtype=np.uint32)
    return np.ascontiguousarray(a.reshape(len(vals), 8))


def debug_key(ins, As, Qs, wide=False, force=False, device=0):
    """The kernel's key extraction (candidate_key in giant_step.cuh): ins = lambdas (wide=False) or arbitrary
    512-bit values standing for lambda^2 (wide=True); returns (keys, exact_flags) as numpy arrays."""
    n = len(ins)
    if wide:
        I_ = np.ascontiguousarray(np.frombuffer(b"".join(int(x).to_bytes(64, "little") for x in ins), dtype=np.uint32).reshape(n, 16))
    else:
        I_ = _limbs(ins)
    A_, Q_ = _limbs(As), _limbs(Qs)
    keys = np.zeros(n, dtype=np.uint64); ex = np.zeros(n, dtype=np.uint8)
    rc = load().bsgs_debug_key(device, 1 if wide else 0, I_.ctypes.data, A_.ctypes.data, Q_.ctypes.data, keys.ctypes.data, ex.ctypes.data, n, 1 if force else 0)
    if rc:
        raise BsgsError(rc, "bsgs_debug_key")
    return keys, ex


def _strerror(status):
    try:
        return load().bsgs_strerror(status).decode()
    except Exception:
        return "?"


def _le32(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def _pxy(points) -> bytes:
    return b"".join(_le32(x) + _le32(y) for x, y in points)


class Engine:
    """One GPU context: the role of the reference's per-GPU worker `cuda()` (1_9_7File.pb:2095-2553)."""

    def __init__(self, T, B, p, w, htsz, device=0, **options):
        self.L = load()
        self.h = C.c_void_p()
        self.T, self.B, self.p, self.w, self.htsz = T, B, p, w, htsz
        self.N = T * B * p
        rc = self.L.bsgs_create(C.byref(self.h), device, T, B, p, w, htsz)
        if rc:
            self.h = C.c_void_p()
            raise BsgsError(rc, "bsgs_create")
        for k, v in options.items():
            self._ck(self.L.bsgs_set_option(self.h, k.encode(), int(v)), f"bsgs_set_option({k})")
        self.cap = 1 << 16
        self._hits = (Hit * self.cap)()
        self.verify = bool(options.get("verify", 0))
        self.payloads = []                                   # parallel to the last returned records (verify=1): j or 0xFFFFFFFE

    def _ck(self, rc, what, allow=()):
        if rc and rc not in allow:
            raise BsgsError(rc, what, self.L.bsgs_last_cuda_error(self.h))
        return rc

    # ---- tables
    def upload_giants(self, g2: np.ndarray):
        g2 = np.ascontiguousarray(g2)
        self._ck(self.L.bsgs_upload_giants(self.h, g2.ctypes.data, g2.nbytes), "bsgs_upload_giants")

    def upload_ht(self, htgpu: np.ndarray):
        htgpu = np.ascontiguousarray(htgpu)
        self._ck(self.L.bsgs_upload_ht(self.h, htgpu.ctypes.data, htgpu.nbytes), "bsgs_upload_ht")

    def upload_htcpu(self, htcpu: np.ndarray):
        """The htCPU file image: same table plus the payloads j (needed by verify=1)."""
        htcpu = np.ascontiguousarray(htcpu)
        self._ck(self.L.bsgs_upload_htcpu(self.h, htcpu.ctypes.data, htcpu.nbytes), "bsgs_upload_htcpu")

    def set_verify(self, on: bool):
        """Device-side verification of hit records: false positives dropped, payload j attached."""
        self._ck(self.L.bsgs_set_option(self.h, b"verify", 1 if on else 0), "bsgs_set_option(verify)")
        self.verify = bool(on)

    def giants_staging(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.bsgs_giants_staging(self.h, C.byref(p), C.byref(n)), "bsgs_giants_staging")
        return p.value, n.value

    def giants_commit(self):
        self._ck(self.L.bsgs_giants_commit(self.h), "bsgs_giants_commit")

    def ht_staging(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.bsgs_ht_staging(self.h, C.byref(p), C.byref(n)), "bsgs_ht_staging")
        return p.value, n.value

    def ht_commit(self):
        self._ck(self.L.bsgs_ht_commit(self.h), "bsgs_ht_commit")

    def giants_generate(self):
        self._ck(self.L.bsgs_giants_generate(self.h), "bsgs_giants_generate")

    def giants_export(self) -> np.ndarray:
        out = np.empty(self.N * 64, dtype=np.uint8)
        self._ck(self.L.bsgs_giants_export(self.h, out.ctypes.data, out.nbytes), "bsgs_giants_export")
        return out

    def ht_build_from_keys(self, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        self._ck(self.L.bsgs_ht_build_from_keys(self.h, keys.ctypes.data, len(keys)), "bsgs_ht_build_from_keys")

    def ht_generate(self):
        """Real baby-step table, generated and sorted on the device."""
        self._ck(self.L.bsgs_ht_generate(self.h), "bsgs_ht_generate")

    def ht_build_synthetic(self, seed=1):
        self._ck(self.L.bsgs_ht_build_synthetic(self.h, seed), "bsgs_ht_build_synthetic")

    def ht_export(self, kind="gpu") -> np.ndarray:
        H = 1 << self.htsz
        n = 4 * H + 4 + self.w * (4 if kind == "gpu" else 8)
        out = np.empty(n, dtype=np.uint8)
        self._ck(self.L.bsgs_ht_export(self.h, 0 if kind == "gpu" else 1, out.ctypes.data, n), "bsgs_ht_export")
        return out

    def ht_device_image(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.bsgs_ht_device_image(self.h, C.byref(p), C.byref(n)), "bsgs_ht_device_image")
        return p.value, n.value

    # ---- hot path
    def _take(self, n):
        n = min(n, self.cap)
        self.payloads = [int(self._hits[i].reserved) for i in range(n)]
        return [(int(self._hits[i].code), int(self._hits[i].idx), int(self._hits[i].window)) for i in range(n)]

    def step(self, P_xy):
        """One reference launch -> list of (code, idx)."""
        n = C.c_uint32()
        self._ck(self.L.bsgs_step(self.h, _le32(P_xy[0]), _le32(P_xy[1]), self._hits, self.cap, C.byref(n)), "bsgs_step")
        return [(c, i) for c, i, _ in self._take(n.value)]

    def step_multi(self, points):
        """len(points) consecutive windows in one launch -> list of (code, idx, window)."""
        n = C.c_uint32()
        self._ck(self.L.bsgs_step_multi(self.h, len(points), _pxy(points), self._hits, self.cap, C.byref(n)), "bsgs_step_multi")
        return self._take(n.value)

    def run(self, P0, D, nwindows, batch=16, stop_after_hits=0):
        n, done = C.c_uint32(), C.c_uint64()
        self._ck(self.L.bsgs_run(self.h, _pxy([P0]), _pxy([D]), nwindows, batch, stop_after_hits, self._hits, self.cap,
                                 C.byref(n), C.byref(done)), "bsgs_run")
        return self._take(n.value), done.value

    # ---- measurement
    def time_launches(self, points, nwin, iters):
        """Device time (ms) of `iters` launches of `nwin` windows each; points holds iters*nwin base points."""
        assert len(points) == nwin * iters
        ms = C.c_float()
        self._ck(self.L.bsgs_time_launches(self.h, nwin, _pxy(points), iters, C.byref(ms)), "bsgs_time_launches")
        return ms.value

    def probe_bench(self, nkeys, seed=1, iters=1):
        ms, f = C.c_float(), C.c_uint64()
        self._ck(self.L.bsgs_probe_bench(self.h, nkeys, seed, iters, C.byref(ms), C.byref(f)), "bsgs_probe_bench")
        return ms.value, f.value

    def probe_keys(self, keys: np.ndarray) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(len(keys), dtype=np.uint8)
        self._ck(self.L.bsgs_probe_keys(self.h, keys.ctypes.data, len(keys), out.ctypes.data), "bsgs_probe_keys")
        return out.astype(bool)

    def trace(self):
        """(nwarps, 4) array of {SM id, t_start, t_pass1, t_end} (ns) of the last launch (option trace=1)."""
        n = self.info()["grid"] * self.info()["block"] // 32
        out = np.zeros((n, 4), dtype=np.uint64); got = C.c_uint32()
        self._ck(self.L.bsgs_debug_trace(self.h, out.ctypes.data, n, C.byref(got)), "bsgs_debug_trace")
        return out[:got.value]

    def info(self):
        i = Info()
        self._ck(self.L.bsgs_get_info(self.h, C.byref(i)), "bsgs_get_info")
        return {k: int(getattr(i, k)) for k, _ in Info._fields_}

    def close(self):
        if self.h:
            self.L.bsgs_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
