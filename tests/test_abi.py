"""The C-ABI shared library loads and exports every symbol include/ribbit_scan.h declares (no compute calls: CPU box)."""
import ctypes
import os
import re

from ribbit_b200 import build, scan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    path = build.build_cuda()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "ribbit_scan.h")).read()
    declared = set(re.findall(r"\b(rb_[a-z_]+)\s*\(", hdr))
    assert declared == set(scan.EXPORTS), declared ^ set(scan.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert scan.load_library().rb_abi_version() == 1


def test_struct_layouts_match_header():
    assert scan.REC_DTYPE.itemsize == 16
    assert ctypes.sizeof(scan.RbParams) == 16
    assert ctypes.sizeof(scan.RbTiming) == 32
    assert ctypes.sizeof(scan.RbStreams) == 8 + 24 + 24 + 24


def test_create_fails_loudly_without_gpu_or_with_bad_params():
    import torch
    lib = scan.load_library()
    bad = scan.RbParams(5, 2, 0, 0)
    assert not lib.rb_create(0, ctypes.byref(bad))
    assert b"bad parameters" in lib.rb_last_error(None)
    if not torch.cuda.is_available():
        ok = scan.RbParams(2, 100, 0, 0)
        assert not lib.rb_create(0, ctypes.byref(ok))
        assert b"no CPU path" in lib.rb_last_error(None)
