"""The C-ABI library loads and exports every symbol include/circuitmap_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "circuitmap_b200.h")).read()
    return sorted(set(re.findall(r"CM_API\s+[\w\s\*]+?\b(cm_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for n in ["cm_version", "cm_last_error", "cm_nwd_create", "cm_nwd_forward", "cm_nwd_destroy",
              "cm_caviar_workspace_bytes", "cm_caviar_fit", "cm_last_launch_count", "cm_last_main_kernel_ms"]:
        assert n in names


def test_library_exports_every_declared_symbol():
    from circuitmap_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in _declared():
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == _declared()
    lib.cm_version.restype = ctypes.c_int
    assert lib.cm_version() == 100
    lib.cm_caviar_workspace_bytes.restype = ctypes.c_size_t
    lib.cm_caviar_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int]
    small = lib.cm_caviar_workspace_bytes(1, 100, 2000, 20000, 0)
    big = lib.cm_caviar_workspace_bytes(4, 100, 2000, 20000, 0)
    assert 0 < small < big and lib.cm_caviar_workspace_bytes(0, 1, 1, 1, 0) == 0


def test_struct_layout_matches_header():
    """ctypes mirrors of cm_caviar_options / cm_caviar_args have the C sizes (checked by compiling a probe)."""
    import subprocess, tempfile
    from circuitmap_b200 import _lib
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "p.c")
        open(src, "w").write('#include <stdio.h>\n#include "circuitmap_b200.h"\nint main(){printf("%zu %zu %zu\\n",'
                             'sizeof(cm_caviar_options),sizeof(cm_caviar_args),sizeof(cm_sim_options));return 0;}')
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        a, b, c = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(_lib.CaviarOptions) == a and ctypes.sizeof(_lib.CaviarArgs) == b
    assert ctypes.sizeof(_lib.SimOptions) == c
