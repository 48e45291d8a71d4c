"""The C-ABI library loads on a machine without a GPU and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re
import pytest
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(bsgs_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_boundary():
    fns = declared_functions()
    for must in ("bsgs_create", "bsgs_destroy", "bsgs_upload_giants", "bsgs_upload_ht", "bsgs_step", "bsgs_run", "bsgs_strerror"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    import bsgs_cuda_b200 as b
    path = b.lib_path()
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "bsgs-cuda_b200"), "libbsgs_b200.so"])
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert sorted(b.native.EXPORTS) == declared_functions()


def test_no_compute_without_gpu_but_clean_errors():
    """Without a device the library must return an error status, never crash or fall back."""
    import bsgs_cuda_b200 as b
    import torch
    if torch.cuda.is_available():
        return
    L = b.native.load()
    assert L.bsgs_strerror(0) == b"ok" and L.bsgs_strerror(-5) != b"ok"
    try:
        b.Engine(8, 2, 2, 100, 4)
    except b.BsgsError as e:
        assert e.status in (-2, -3)
    else:
        raise AssertionError("bsgs_create succeeded without a GPU")


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "bsgs-cuda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert "oracle" not in open(os.path.join(root, f), errors="replace").read().replace("CPU oracle", "").replace("oracle's", "").lower().replace("the oracle", ""), f


def test_dropin_module_exports_the_reference_kernel_interface():
    """Option A of the boundary: the module must export `.entry _test1(.param .u64)` and a 120-byte
    `.const _A` (cuModuleGetFunction / cuModuleGetGlobal_v2 by these names, 1_9_7File.pb:2198, 2278)."""
    import re
    ptx = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bsgs-cuda_b200", "dropin_test1.ptx")
    if not os.path.exists(ptx):
        pytest.skip("dropin_test1.ptx not built")
    text = open(ptx).read()
    assert re.search(r"\.visible \.entry _test1\(\s*\.param \.u64[^,)]*\)", text)
    assert re.search(r"\.const \.align \d+ \.b8 _A\[120\];", text)
    assert ".target sm_100a" in text


def test_datasection_writer_round_trip(tmp_path):
    """scripts/make_datasection.py embeds the module the way the reference embeds its own kernel
    (XOR 93, Data.q quadwords, 1_9_7File.pb:4739-4741, 5181-23979); decoding gives the PTX back."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_datasection", os.path.join(root, "scripts", "make_datasection.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    ptx = b"//\n.version 8.8\n.target sm_100a\n.visible .entry _test1(.param .u64 p){ret;}\n"
    text = m.encode(ptx)
    assert text.startswith("DataSection") and "BSGS4_cuda_quad_htchangeble_v2:" in text
    assert m.decode(text).rstrip(b"\0") == ptx


def test_purebasic_binding_lists_only_declared_functions():
    """include/bsgs_b200.pbi (the binding a PureBasic maintainer includes) names only functions the
    header declares, the whole reference-facing set among them, and the same status codes."""
    pbi = open(os.path.join(ROOT, "include", "bsgs_b200.pbi")).read()
    code = "\n".join(l.split(";")[0] for l in pbi.splitlines())
    listed = set(re.findall(r"\b(bsgs_[a-z0-9_]+)\s*\(", code))
    assert listed <= set(declared_functions())
    for must in ("bsgs_create", "bsgs_destroy", "bsgs_upload_giants", "bsgs_upload_ht", "bsgs_step", "bsgs_step_multi", "bsgs_run",
                 "bsgs_giants_generate", "bsgs_ht_generate", "bsgs_ht_export", "bsgs_bcast_ht", "bsgs_set_option"):
        assert must in listed
    hdr = open(os.path.join(ROOT, "include", "bsgs_b200.h")).read()
    for name, val in re.findall(r"#(BSGS_[A-Z_]+)\s*=\s*(-?\d+)", code):
        if name == "BSGS_MAX_WINDOWS":
            assert re.search(r"#define BSGS_MAX_WINDOWS\s+%s\b" % val, hdr)
        else:
            assert re.search(r"\b%s\s*=\s*%s\b" % (name, val), hdr), name


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 with no C++ or CUDA types in any signature."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "bsgs_b200.h"\nint main(void) { return bsgs_strerror(0) == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
