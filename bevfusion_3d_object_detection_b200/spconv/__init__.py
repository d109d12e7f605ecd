"""spconv-2.x compatible surface (the names the reference imports from `spconv.pytorch`) on libbevfront_b200."""
from .conv import (OFF_PATH_CLASSES, SparseConv2d, SparseConv3d, SparseConv4d, SparseConvolution, SparseConvTranspose2d,
                   SparseConvTranspose3d, SparseInverseConv2d, SparseInverseConv3d, SubMConv2d, SubMConv3d, SubMConv4d,
                   get_default_precision, set_default_precision)
from .core import CoordIndex, IndicePair, SparseConvTensor
from .modules import SparseModule, SparseSequential

__version__ = "2.3.6+b200"
__all__ = ["SparseConvTensor", "SparseModule", "SparseSequential", "SparseConvolution", "SubMConv3d", "SparseConv3d",
           "IndicePair", "CoordIndex", "set_default_precision", "get_default_precision", "SparseConv2d", "SparseConv4d",
           "SubMConv2d", "SubMConv4d", "SparseConvTranspose2d", "SparseConvTranspose3d", "SparseInverseConv2d",
           "SparseInverseConv3d", "OFF_PATH_CLASSES"]
