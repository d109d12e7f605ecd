"""Load pieces of the REAL reference (/root/reference) in this container -- TEST INFRASTRUCTURE ONLY.

mmengine / mmcv / mmdet / spconv are not installed, so the reference modules are loaded by file path with
minimal stubs for the imports they cannot resolve (nothing in the reference is edited or copied).  Used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by CPU tests that can see
/root/reference.  /root/reference does not exist on the GPU box: nothing marked ``gpu`` may call this.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("BEVFRONT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "projects", "BEVFusion"))


def _stub_registry():
    if "mmdet3d" not in sys.modules:
        m = types.ModuleType("mmdet3d")
        m.__path__ = []
        sys.modules["mmdet3d"] = m
    if "mmdet3d.registry" not in sys.modules:
        reg = types.ModuleType("mmdet3d.registry")

        class _Reg:
            def register_module(self, *a, **k):
                def deco(cls):
                    return cls
                return deco

        reg.MODELS = _Reg()
        reg.TASK_UTILS = _Reg()
        sys.modules["mmdet3d.registry"] = reg
        sys.modules["mmdet3d"].registry = reg


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def voxel_generator():
    """mmdet3d/models/task_modules/voxel/voxel_generator.py (numba): points_to_voxel, VoxelGenerator."""
    _stub_registry()
    return _load("_ref_voxel_generator", "mmdet3d/models/task_modules/voxel/voxel_generator.py")


def bev_pool_py():
    """projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py with a stub bev_pool_ext (QuickCumsum is pure torch)."""
    pkg = types.ModuleType("_ref_bev_pool_pkg")
    pkg.__path__ = []
    sys.modules["_ref_bev_pool_pkg"] = pkg
    ext = types.ModuleType("_ref_bev_pool_pkg.bev_pool_ext")
    sys.modules["_ref_bev_pool_pkg.bev_pool_ext"] = ext
    pkg.bev_pool_ext = ext
    mod = _load("_ref_bev_pool_pkg.bev_pool", "projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py")
    return mod, ext


def depth_lss(bev_pool_fn=None):
    """projects/BEVFusion/bevfusion/depth_lss.py with `.ops.bev_pool` stubbed by `bev_pool_fn`."""
    _stub_registry()
    pkg = types.ModuleType("_ref_bevfusion_pkg")
    pkg.__path__ = []
    sys.modules["_ref_bevfusion_pkg"] = pkg
    ops = types.ModuleType("_ref_bevfusion_pkg.ops")
    ops.bev_pool = bev_pool_fn if bev_pool_fn is not None else (lambda *a, **k: None)
    sys.modules["_ref_bevfusion_pkg.ops"] = ops
    pkg.ops = ops
    return _load("_ref_bevfusion_pkg.depth_lss", "projects/BEVFusion/bevfusion/depth_lss.py")
