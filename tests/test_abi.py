"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly what
include/meshrcnn_b200.h declares, with the ctypes signatures the Python host code uses (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "meshrcnn_b200.h")


def _parse_header():
    from meshrcnn_b200 import _lib
    return _lib.parse_header(HEADER)


def test_header_matches_ctypes_table():
    from meshrcnn_b200 import _lib
    protos = _parse_header()
    n_decl = len(re.findall(r"\bmrb_\w+\s*\(", re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)))
    assert len(protos) == n_decl >= 25          # every declaration was parsed
    assert protos == _lib.SIGNATURES
    assert protos["mrb_cubify_emit"] == ("i", "iiiipplllpppppppp")
    assert protos["mrb_sgemm"] == ("i", "iiiiipipifpip")


def test_library_builds_loads_and_exports_every_symbol(lib):
    from meshrcnn_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _parse_header():
        assert hasattr(raw, name), "missing export " + name
    assert lib.mrb_version() == 100
    assert lib.mrb_cubify_workspace_bytes(2, 8, 8, 8) > 0
    assert lib.mrb_cubify_workspace_bytes(0, 8, 8, 8) == -1


def test_sm100a_only():
    """The shared object carries sm_100a SASS and nothing else (no multi-arch fatbin, no PTX JIT fallback)."""
    import shutil
    import subprocess
    from meshrcnn_b200 import _lib, build
    build.build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback():
    import torch
    from meshrcnn_b200.layers import Cubify
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Cubify(0.5)(torch.rand(1, 4, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mesh_r-cnn_computer_vision_project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_fastcall_shim_calls_the_same_entry_points():
    """csrc/host/fastcall.c: all-integer / pointer entry points called through a plain cast give the ctypes results (version,
    error code + message of a rejected call -- no device work, so this runs without a GPU)."""
    from meshrcnn_b200 import _lib, build
    build.build()
    lib = _lib.load()
    if _lib._fast_call is None:
        pytest.skip("call shim not built on this platform")
    addr = lambda name: ctypes.cast(getattr(lib, name), ctypes.c_void_p).value
    assert _lib._fast_call(addr("mrb_version"), ()) == lib.mrb_version() == 100
    assert "mrb_segment_ids" in _lib._FAST and "mrb_sgemm" not in _lib._FAST            # float arguments stay on ctypes
    rc_fast = _lib._fast_call(_lib._FAST["mrb_segment_ids"], (None, 3, 7, None, None))   # null pointers: rejected before any launch
    msg_fast = lib.mrb_last_error()
    rc_ct = lib.mrb_segment_ids(None, 3, 7, None, None)
    assert rc_fast == rc_ct != 0 and msg_fast == lib.mrb_last_error()
    with pytest.raises(TypeError):
        _lib._fast_call(addr("mrb_version"), (1.5,))
