"""The C-ABI library builds, loads, and exports every function include/bevfront_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "bevfront_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bevf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bevfusion_3d_object_detection_b200 import _lib

    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert L.bevf_abi_version() >= 1
    assert L.bevf_compiled_arch() == 100


def test_host_only_helpers_and_error_text():
    from bevfusion_3d_object_detection_b200 import _lib

    L = _lib.lib()
    grid = (ctypes.c_int * 3)()
    assert L.bevf_voxel_grid_size(_lib.f32_array([0.075, 0.075, 0.2]), _lib.f32_array([-54, -54, -5, 54, 54, 3]),
                                  grid) == 0
    assert list(grid) == [1440, 1440, 40]
    # argument validation happens before any CUDA call, so it is testable on CPU
    rc = L.bevf_dynamic_voxelize(None, 10, 2, None, _lib.f32_array([1, 1, 1]), _lib.f32_array([0, 0, 0, 1, 1, 1]), 3,
                                 None)
    assert rc == 1 and b"C>=3" in L.bevf_last_error()
    assert L.bevf_hard_voxelize_workspace_bytes(320000, 10, 160000) > 0


def test_sass_is_sm100a_only():
    """The shipped library carries sm_100a SASS and nothing else (no multi-arch fat binary)."""
    import shutil
    import subprocess

    from bevfusion_3d_object_detection_b200 import build

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest

        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_tensor_core_variant_knob_is_host_side():
    """bevf_spconv_tc_variant only flips a process-wide setting (no CUDA call): query, set, restore."""
    from bevfusion_3d_object_detection_b200 import _lib

    L = _lib.lib()
    cur = L.bevf_spconv_tc_variant(-1)
    assert cur in (0, 1, 2)
    assert L.bevf_spconv_tc_variant(0) == cur
    assert L.bevf_spconv_tc_variant(-1) == 0
    assert L.bevf_spconv_tc_variant(7) == 0          # out of range: ignored
    assert L.bevf_spconv_tc_variant(cur) == 0
    assert L.bevf_spconv_tc_variant(-1) == cur
    assert L.bevf_spconv_tc_supported(64, 64) == 1 and L.bevf_spconv_tc_supported(5, 16) == 1
    assert L.bevf_spconv_tc_supported(256, 64) == 0 and L.bevf_spconv_tc_cin_pad(5) == 16
