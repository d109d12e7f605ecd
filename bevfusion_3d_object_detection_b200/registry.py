"""Registration shim: the reference builds its sparse layers through mmengine's MODELS registry
(`build_conv_layer(dict(type='SubMConv3d', indice_key=...), ...)`, mmdet3d/models/layers/sparse_block.py:201-217)
after overwriting the names with the spconv-2.x classes (write_spconv2.py:21-38).  `register_all()` does the same
for the B200 classes -- into mmengine's registry when mmengine is installed (force=True, exactly like the
reference), and always into the local registry below so the package works without mmengine / mmcv.
"""
from torch import nn

_LOCAL = {}


def register_module(name=None, force=False):
    def deco(cls):
        key = name or cls.__name__
        if key in _LOCAL and not force and _LOCAL[key] is not cls:
            raise KeyError(f"{key} is already registered")
        _LOCAL[key] = cls
        return cls
    return deco


def get(name):
    return _LOCAL[name]


def build_conv_layer(cfg, *args, **kwargs):
    """mmcv.cnn.build_conv_layer for the sparse conv types (cfg = dict(type=..., indice_key=...))."""
    cfg = dict(cfg)
    layer_type = cfg.pop("type")
    if layer_type not in _LOCAL:
        raise KeyError(f"Cannot find {layer_type} in the registry")
    return _LOCAL[layer_type](*args, **kwargs, **cfg)


def build_norm_layer(cfg, num_features, postfix=""):
    """mmcv.cnn.build_norm_layer for BN1d / BN: returns (name, layer)."""
    cfg = dict(cfg)
    layer_type = cfg.pop("type")
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    if layer_type not in ("BN1d", "BN", "naiveSyncBN1d"):
        raise KeyError(f"norm layer {layer_type} is not available without mmcv")
    layer = nn.BatchNorm1d(num_features, **cfg)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return "bn" + str(postfix), layer


def register_all():
    """Register the B200 classes under the reference's names (write_spconv2.py:21-38 registers ten conv classes: the two on
    the BEVFusion path are built here, the other eight are registered too and raise NotImplementedError when constructed);
    returns True if mmengine's registry was updated."""
    from . import spconv
    from .sparse_encoder import BEVFusionSparseEncoder

    classes = [(spconv.SubMConv3d, "SubMConv3d"), (spconv.SparseConv3d, "SparseConv3d"),
               (BEVFusionSparseEncoder, "BEVFusionSparseEncoder")] + [(c, c.__name__) for c in spconv.OFF_PATH_CLASSES]
    for cls, nm in classes:
        _LOCAL[nm] = cls
    spconv.SparseModule._version = 2      # write_spconv2.py:38 (the checkpoint shim lives in SparseModule._load_from_state_dict)
    try:
        from mmengine.registry import MODELS
    except Exception:
        return False
    for cls, nm in classes:
        MODELS._register_module(cls, nm, force=True)
    return True
