"""bsgs-cuda_b200 -- B200-native giant-step engine for secp256k1 BSGS (drop-in for the GPU path of
Etayson/BSGS-cuda).  The compute path is libbsgs_b200.so (hand-written sm_100a CUDA behind the C ABI
of include/bsgs_b200.h); this package is the thin host side: ctypes binding, the reference's file
formats, window enumeration (GetJob) and hit decoding (checkerThread) mirrored from 1_9_7File.pb.
There is no CPU fallback: importing `native` without the built library raises."""
from . import ec, formats, tune                             # noqa: F401
from .native import Engine, BsgsError, lib_path               # noqa: F401
from .search import Search, solve                            # noqa: F401
from . import cli                                            # noqa: F401

__all__ = ["Engine", "BsgsError", "Search", "solve", "ec", "formats", "tune", "cli", "lib_path"]
