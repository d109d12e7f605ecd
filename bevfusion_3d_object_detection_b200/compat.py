"""Install-name shim: make the reference's own import lines resolve to this library, unedited.

The reference builds two pybind extensions under fixed dotted names (projects/BEVFusion/setup.py:49-67):

    projects.BEVFusion.bevfusion.ops.bev_pool.bev_pool_ext      imported by ops/bev_pool/bev_pool.py:4   (from . import bev_pool_ext)
    projects.BEVFusion.bevfusion.ops.voxel.voxel_layer          imported by ops/voxel/voxelize.py:7, scatter_points.py:5

`install()` registers this package's `bev_pool_ext` / `voxel_layer` modules in `sys.modules` under exactly those names
(and, for convenience, under any other parent package given).  Python's import machinery consults `sys.modules` before
looking for a file, so `from . import bev_pool_ext` inside the reference's `bev_pool.py` binds our module although no
compiled extension exists in the reference tree.  Nothing in the reference is edited:

    import bevfusion_3d_object_detection_b200.compat as compat
    compat.install()                       # before `import projects.BEVFusion.bevfusion`
    from projects.BEVFusion.bevfusion import BEVFusion     # its ops now run on libbevfront_b200

`register_spconv()` does the same for the sparse-conv classes: it registers SubMConv3d / SparseConv3d /
SparseConvTensor / SparseSequential in mmengine's MODELS (when mmengine is importable) under the names
mmdet3d/models/layers/spconv/overwrite_spconv/write_spconv2.py:21-38 uses, and aliases `spconv.pytorch` so that
`from spconv.pytorch import SparseConvTensor, SparseSequential` (sparse_block.py:11-14, sparse_encoder.py:7-10)
resolves here when the real spconv is absent.
"""
import sys
import types

REFERENCE_OPS_PACKAGE = "projects.BEVFusion.bevfusion.ops"


def install(parent=REFERENCE_OPS_PACKAGE, overwrite=False):
    """Alias `<parent>.bev_pool.bev_pool_ext` and `<parent>.voxel.voxel_layer` to this library's modules.
    Returns the list of names registered.  Existing entries are kept unless `overwrite`."""
    from .ops.bev_pool import bev_pool_ext
    from .ops.voxel import voxel_layer

    names = {parent + ".bev_pool.bev_pool_ext": bev_pool_ext, parent + ".voxel.voxel_layer": voxel_layer}
    done = []
    for name, mod in names.items():
        if overwrite or name not in sys.modules:
            sys.modules[name] = mod
            done.append(name)
    return done


def uninstall(parent=REFERENCE_OPS_PACKAGE):
    for name in (parent + ".bev_pool.bev_pool_ext", parent + ".voxel.voxel_layer"):
        sys.modules.pop(name, None)


def register_spconv(alias_spconv_package=True):
    """Registry + `spconv.pytorch` alias for the sparse-conv surface (see module docstring)."""
    from . import registry, spconv as our_spconv

    registry.register_all()
    if alias_spconv_package and "spconv" not in sys.modules:
        pkg = types.ModuleType("spconv")
        pkg.__path__ = []
        pkg.__version__ = "2.3.6+bevfront_b200"     # mmdet3d/models/layers/spconv/__init__.py:9 checks >= 2.0.0
        pkg.pytorch = our_spconv
        sys.modules["spconv"] = pkg
        sys.modules["spconv.pytorch"] = our_spconv
    return our_spconv
