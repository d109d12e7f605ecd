"""spconv-2.x compatible surface (the names the reference imports from `spconv.pytorch`) on libbevfront_b200."""
from .conv import (SparseConv3d, SparseConvolution, SubMConv3d, get_default_precision, set_default_precision)
from .core import CoordIndex, IndicePair, SparseConvTensor
from .modules import SparseModule, SparseSequential

__version__ = "2.3.6+b200"
__all__ = ["SparseConvTensor", "SparseModule", "SparseSequential", "SparseConvolution", "SubMConv3d", "SparseConv3d",
           "IndicePair", "CoordIndex", "set_default_precision", "get_default_precision"]
