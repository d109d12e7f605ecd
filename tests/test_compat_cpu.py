"""The install-name shim and call compatibility with the REAL reference Python files (loaded from /root/reference when it
is visible: CPU container only; the GPU box does not have it and these tests skip there).  No compute: signatures and
import resolution."""
import importlib
import importlib.util
import inspect
import os
import sys
import types

import pytest

from bevfusion_3d_object_detection_b200 import compat

REF = os.environ.get("BEVFRONT_REFERENCE_ROOT", "/root/reference")
OPS = os.path.join(REF, "projects", "BEVFusion", "bevfusion", "ops")
needs_ref = pytest.mark.skipif(not os.path.isdir(OPS), reason="/root/reference not visible")


def test_install_registers_the_reference_extension_names():
    compat.uninstall()
    names = compat.install()
    try:
        assert names == ["projects.BEVFusion.bevfusion.ops.bev_pool.bev_pool_ext",
                         "projects.BEVFusion.bevfusion.ops.voxel.voxel_layer"]   # projects/BEVFusion/setup.py:49-67
        ext = sys.modules[names[0]]
        vl = sys.modules[names[1]]
        assert callable(ext.bev_pool_forward) and callable(ext.bev_pool_backward)         # bev_pool.cpp:89-94
        for fn in ("hard_voxelize", "dynamic_voxelize", "dynamic_point_to_voxel_forward",
                   "dynamic_point_to_voxel_backward"):                                    # voxelization.cpp:6-11
            assert callable(getattr(vl, fn))
        assert compat.install() == []     # idempotent: existing entries are kept
    finally:
        compat.uninstall()


def _load_as(name, path, package_stub=True):
    """Load a reference file under the dotted name the reference itself would give it, parents stubbed as namespace
    packages (their real __init__ files import mmengine)."""
    parts = name.split(".")
    for i in range(1, len(parts)):
        pn = ".".join(parts[:i])
        if pn not in sys.modules:
            m = types.ModuleType(pn)
            m.__path__ = []
            sys.modules[pn] = m
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@needs_ref
def test_reference_python_ops_import_unedited_against_the_shim():
    """ops/bev_pool/bev_pool.py, ops/voxel/voxelize.py and scatter_points.py of the reference are executed AS THEY ARE
    under their own dotted names; their `from . import bev_pool_ext` / `from .voxel_layer import ...` lines bind this
    library through the sys.modules aliases."""
    created = [k for k in list(sys.modules) if k.startswith("projects")]
    compat.install(overwrite=True)
    try:
        bp = _load_as("projects.BEVFusion.bevfusion.ops.bev_pool.bev_pool", os.path.join(OPS, "bev_pool", "bev_pool.py"))
        vx = _load_as("projects.BEVFusion.bevfusion.ops.voxel.voxelize", os.path.join(OPS, "voxel", "voxelize.py"))
        sc = _load_as("projects.BEVFusion.bevfusion.ops.voxel.scatter_points", os.path.join(OPS, "voxel", "scatter_points.py"))
        from bevfusion_3d_object_detection_b200.ops.bev_pool import bev_pool_ext
        from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer

        assert bp.bev_pool_ext is bev_pool_ext
        assert vx.hard_voxelize is voxel_layer.hard_voxelize and vx.dynamic_voxelize is voxel_layer.dynamic_voxelize
        assert sc.dynamic_point_to_voxel_forward is voxel_layer.dynamic_point_to_voxel_forward
        # the reference's Python layer is intact on top: same public classes / functions
        for name in ("QuickCumsum", "QuickCumsumCuda", "QuickCumsumTrainingCuda", "bev_pool"):
            assert hasattr(bp, name)
        assert hasattr(vx, "Voxelization") and hasattr(sc, "DynamicScatter")
    finally:
        compat.uninstall()
        for k in [k for k in list(sys.modules) if k.startswith("projects") and k not in created]:
            sys.modules.pop(k, None)


@needs_ref
def test_call_compatibility_with_the_reference_call_sites():
    """Arity / parameter names of our entry points against how the reference calls them (bev_pool.py:60-66, 86-95,
    voxelize.py:51-64, scatter_points.py:27, 39-47) and against the public signatures of its Python wrappers."""
    from bevfusion_3d_object_detection_b200 import ops
    from bevfusion_3d_object_detection_b200.ops.bev_pool import bev_pool_ext
    from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer

    def params(fn):
        return list(inspect.signature(fn).parameters)

    # pybind functions are called positionally: arities and order
    assert params(bev_pool_ext.bev_pool_forward) == ["x", "geom_feats", "interval_lengths", "interval_starts", "b", "d", "h", "w"]
    assert params(bev_pool_ext.bev_pool_backward) == ["out_grad", "geom_feats", "interval_lengths", "interval_starts", "b", "d", "h", "w"]
    p = inspect.signature(voxel_layer.hard_voxelize).parameters
    assert list(p)[:8] == ["points", "voxels", "coors", "num_points_per_voxel", "voxel_size", "coors_range", "max_points", "max_voxels"]
    assert p["NDim"].default == 3 and p["deterministic"].default is True          # voxelization.h:58-66
    assert params(voxel_layer.dynamic_voxelize)[:4] == ["points", "coors", "voxel_size", "coors_range"]
    assert len(params(voxel_layer.dynamic_point_to_voxel_forward)) == 3
    assert len(params(voxel_layer.dynamic_point_to_voxel_backward)) == 7
    # python wrappers: the same parameter names and defaults as the reference classes
    compat.install(overwrite=True)
    created = [k for k in list(sys.modules) if k.startswith("projects")]
    try:
        bp = _load_as("projects.BEVFusion.bevfusion.ops.bev_pool.bev_pool", os.path.join(OPS, "bev_pool", "bev_pool.py"))
        vx = _load_as("projects.BEVFusion.bevfusion.ops.voxel.voxelize", os.path.join(OPS, "voxel", "voxelize.py"))
        sc = _load_as("projects.BEVFusion.bevfusion.ops.voxel.scatter_points", os.path.join(OPS, "voxel", "scatter_points.py"))
        assert params(ops.bev_pool) == params(bp.bev_pool)
        ref_v = inspect.signature(vx.Voxelization.__init__).parameters
        our_v = inspect.signature(ops.Voxelization.__init__).parameters
        assert list(ref_v) == list(our_v)
        for k in ref_v:
            assert ref_v[k].default == our_v[k].default, k
        assert list(inspect.signature(sc.DynamicScatter.__init__).parameters) == \
            list(inspect.signature(ops.DynamicScatter.__init__).parameters)
    finally:
        compat.uninstall()
        for k in [k for k in list(sys.modules) if k.startswith("projects") and k not in created]:
            sys.modules.pop(k, None)


def test_spconv_alias_and_registry():
    had = "spconv" in sys.modules
    mod = compat.register_spconv()
    try:
        from spconv.pytorch import SparseConvTensor, SparseSequential, SubMConv3d, SparseConv3d   # sparse_block.py:11-14

        assert SparseConvTensor is mod.SparseConvTensor and SparseSequential is mod.SparseSequential
        assert SubMConv3d is mod.SubMConv3d and SparseConv3d is mod.SparseConv3d
        import spconv

        assert tuple(int(v) for v in spconv.__version__.split("+")[0].split(".")[:2]) >= (2, 0)
    finally:
        if not had:
            sys.modules.pop("spconv", None)
            sys.modules.pop("spconv.pytorch", None)


def test_quickcumsum_onnx_symbolic_exports_the_autoware_op():
    """QuickCumsumCuda.symbolic (bev_pool.py:96-125 of the reference): op `autoware::QuickCumsumCuda` with the four
    integer attributes, exercised through a torch.onnx-style graph context stub (no CUDA needed)."""
    from bevfusion_3d_object_detection_b200.ops.bev_pool.bev_pool import QuickCumsumCuda

    calls = []

    class FakeType:   # a graph value whose static sizes are unknown: the symbolic must not need them
        def isSubtypeOf(self, other):
            return False

    class FakeValue:
        def __init__(self, name):
            self.name = name

        def type(self):
            return FakeType()

    class G:
        def op(self, name, *inputs, **attrs):
            calls.append((name, inputs, attrs))
            return "out"

    vals = [FakeValue(n) for n in ("x", "geom", "lengths", "starts")]
    out = QuickCumsumCuda.symbolic(G(), *vals, 1, 1, 360, 360)
    assert out == "out" and len(calls) == 1
    name, inputs, attrs = calls[0]
    assert name == "autoware::QuickCumsumCuda"                         # bev_pool.py:107-118 of the reference
    assert [v.name for v in inputs] == ["x", "geom", "lengths", "starts"]
    assert attrs == dict(batch_size_i=1, dimension_i=1, height_i=360, width_i=360, outputs=1)
