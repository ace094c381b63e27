"""CPU: libcontourist_b200.so loads and exports every entry point include/contourist_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "contourist_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"CTR_API\s+[\w\s\*]+?\b(ctr_\w+)\s*\(", text)))


def test_header_declares_the_api():
    syms = declared_symbols()
    for s in ("ctr_create", "ctr_destroy", "ctr_last_error", "ctr_mt3d_run", "ctr_mt3d_fetch", "ctr_mt2d_run",
              "ctr_mt2d_fetch", "ctr_mp4d_run", "ctr_mp4d_fetch", "ctr_set_stream", "ctr_stage_times"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from contourist_b200 import build, engine
    build.build()
    lib = ctypes.CDLL(engine.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing symbol " + s


def test_bindings_resolve():
    from contourist_b200 import engine
    lib = engine.load_library()
    assert lib.ctr_mt3d_run.restype is ctypes.c_int
    # a null context is rejected without touching CUDA
    assert lib.ctr_mt3d_run(None, None, None) == -1
    assert lib.ctr_kernel_launches(None) == 0
