"""SparseModule / SparseSequential with the spconv-2.x surface (imports at
mmdet3d/models/layers/sparse_block.py:11-14), including the checkpoint weight-layout shim the reference installs
on SparseModule (mmdet3d/models/layers/spconv/overwrite_spconv/write_spconv2.py:38, 43-104).
"""
import itertools
from collections import OrderedDict

import torch
from torch import nn

from .core import SparseConvTensor


def is_spconv_module(module):
    return isinstance(module, SparseModule)


class SparseModule(nn.Module):
    """Marker base class: SparseSequential hands these the SparseConvTensor itself.

    `_version = 2` and `_load_from_state_dict` reproduce the reference shim: a checkpoint whose metadata version
    is not 2 stores conv kernels in the mmcv/spconv-1.x layout (kD, kH, kW, Cin, Cout) and is permuted to the
    spconv-2.x layout (Cout, kD, kH, kW, Cin) while loading."""
    _version = 2

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        version = local_metadata.get("version", None)
        for hook in self._load_state_dict_pre_hooks.values():
            hook(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        local_state = {k: v.data for k, v in itertools.chain(self._parameters.items(), self._buffers.items())
                       if v is not None}
        for name, param in local_state.items():
            key = prefix + name
            if key not in state_dict:
                if strict:
                    missing_keys.append(key)
                continue
            value = state_dict[key]
            if param.dim() == 0 and value.dim() == 1:
                value = value[0]
            if version != 2 and value.dim() > 1:
                value = value.permute(value.dim() - 1, *range(value.dim() - 1))
            if value.shape != param.shape:
                error_msgs.append(f"size mismatch for {key}: copying a param with shape {tuple(value.shape)} from "
                                  f"checkpoint, the shape in current model is {tuple(param.shape)}.")
                continue
            if isinstance(value, nn.Parameter):
                value = value.data
            try:
                param.copy_(value)
            except Exception:
                error_msgs.append(f'While copying the parameter named "{key}", whose dimensions in the model are '
                                  f"{param.size()} and whose dimensions in the checkpoint are {value.size()}.")
        if strict:
            for key in state_dict.keys():
                if key.startswith(prefix):
                    child = key[len(prefix):].split(".", 1)[0]
                    if child not in self._modules and child not in local_state:
                        unexpected_keys.append(key)


def _fold_bn(bn):
    """eval-mode BatchNorm1d as y = x * scale + shift (cached on the module until a parameter / buffer changes)."""
    key = (bn.running_var._version, bn.running_mean._version, bn.running_var.data_ptr(),
           None if bn.weight is None else (bn.weight._version, bn.weight.data_ptr()),
           None if bn.bias is None else (bn.bias._version, bn.bias.data_ptr()), bn.eps)
    hit = getattr(bn, "_bevf_folded", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    folded = _fold_bn_now(bn)
    bn._bevf_folded = (key, folded)
    return folded


def _fold_bn_now(bn):
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    w = bn.weight.float() if bn.weight is not None else torch.ones_like(inv)
    b = bn.bias.float() if bn.bias is not None else torch.zeros_like(inv)
    scale = (w * inv).contiguous()
    shift = (b - bn.running_mean.float() * scale).contiguous()
    return scale, shift


def bn_is_foldable(m):
    return (isinstance(m, nn.BatchNorm1d) and not m.training and m.track_running_stats
            and m.running_mean is not None)


def wants_grad(x, *modules):
    """True when autograd will need a graph through `modules` applied to the sparse tensor `x` (model.eval() without
    no_grad, frozen-BN fine-tuning, input gradients): the fused BN / ReLU / residual epilogues are inference-only, so
    the callers then take the reference's unfused module sequence (spconv modules just run in that situation)."""
    if not torch.is_grad_enabled():
        return False
    f = getattr(x, "_features", None)
    if f is not None and f.requires_grad:
        return True
    return any(p.requires_grad for m in modules for p in m.parameters())


class SparseSequential(SparseModule):
    """nn.Sequential that routes a SparseConvTensor: sparse modules receive the tensor, dense modules (BatchNorm1d,
    ReLU, ...) are applied to `.features`.  In eval mode a [conv, BatchNorm1d, ReLU] run is executed as ONE kernel
    (BN folded into the conv epilogue) -- same values, two fewer passes over the features."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError(f"index {idx} is out of range")
        if idx < 0:
            idx += len(self)
        return next(itertools.islice(self._modules.values(), idx, None))

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        from .conv import SparseConvolution

        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, SparseConvolution) and isinstance(input, SparseConvTensor) and m.fuse_epilogue:
                bn = mods[i + 1] if i + 1 < len(mods) else None
                if bn is not None and bn_is_foldable(bn) and not wants_grad(input, m, bn):
                    relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                    scale, shift = _fold_bn(bn)
                    input = m(input, bn_scale=scale, bn_shift=shift, relu=relu)
                    i += 3 if relu else 2
                    continue
            if (isinstance(m, SparseConvolution) and isinstance(input, SparseConvTensor) and i + 1 < len(mods)
                    and isinstance(mods[i + 1], nn.BatchNorm1d) and mods[i + 1].training):
                # training: conv, then batch-statistics BatchNorm1d + ReLU as one fused op that also emits the bf16
                # operand copy of the next conv (csrc/bn_train.cu)
                from . import bn_train

                out = m(input)
                if bn_train.usable(mods[i + 1], out._features):
                    relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                    y, yb = bn_train.bn_act_train(out._features, mods[i + 1], None, relu,
                                                  want_bf16=m._resolve_precision() == "bf16")
                    input = out.replace_feature(y)
                    input._bf16 = yb
                    i += 3 if relu else 2
                    continue
                input = out
                i += 1
                continue
            if is_spconv_module(m):
                assert isinstance(input, SparseConvTensor)
                input = m(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input = input.replace_feature(m(input.features))
            else:
                input = m(input)
            i += 1
        return input
