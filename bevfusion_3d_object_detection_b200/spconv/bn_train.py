"""Train-mode BatchNorm1d on the rows of a sparse tensor fused with ReLU / residual add and the bf16 operand copy of the
next conv (csrc/bn_train.cu).  Same statistics, running-statistics update and gradients as torch.nn.BatchNorm1d followed
by the separate relu / += identity of the reference (mmdet3d/models/layers/sparse_block.py:137-154); used by
SparseSequential / SparseBasicBlock when the norm layer is a BatchNorm1d in training mode (training split, configs[2]).
"""
import ctypes
import os

import torch
from torch import nn

from .._lib import check, cur_stream, lib, ptr

ENABLED = os.environ.get("BEVFRONT_FUSED_BN_TRAIN", "1") == "1"


def usable(bn, feats):
    return (ENABLED and isinstance(bn, nn.BatchNorm1d) and bn.training and feats is not None and feats.is_cuda
            and feats.dtype == torch.float32 and feats.dim() == 2 and feats.shape[0] > 1 and feats.shape[1] % 4 == 0
            and feats.shape[1] <= 256)


class _BnActTrain(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, bn, relu, want_bf16):
        x = x.contiguous()
        n, c = x.shape
        dev = x.device
        if residual is not None:
            residual = residual.contiguous().float()
        y = torch.empty_like(x)
        y_bf16 = torch.empty((n, c), dtype=torch.bfloat16, device=dev) if want_bf16 else None
        mean = torch.empty(c, dtype=torch.float32, device=dev)
        invstd = torch.empty(c, dtype=torch.float32, device=dev)
        sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
        rm = rv = None
        momentum = 0.0
        if bn.track_running_stats and bn.running_mean is not None:
            bn.num_batches_tracked.add_(1)
            rm, rv = bn.running_mean, bn.running_var
            momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item())
        with torch.cuda.device(dev):
            check(lib().bevf_bn_train_forward(ptr(x), ptr(residual), ptr(gamma), ptr(beta), int(n), int(c),
                                              ctypes.c_float(float(bn.eps)), ctypes.c_float(float(momentum)), int(bool(relu)), ptr(rm), ptr(rv), ptr(mean), ptr(invstd),
                                              ptr(sums), ptr(y), ptr(y_bf16), cur_stream(dev)))
        ctx.save_for_backward(x, y, gamma, mean, invstd)
        ctx.relu, ctx.has_res, ctx.has_beta, ctx.want_bf16 = bool(relu), residual is not None, beta is not None, want_bf16
        if y_bf16 is not None:
            ctx.mark_non_differentiable(y_bf16)
        return y, y_bf16

    @staticmethod
    def backward(ctx, dy, _unused=None):
        x, y, gamma, mean, invstd = ctx.saved_tensors
        n, c = x.shape
        dev = x.device
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        dx_bf16 = torch.empty((n, c), dtype=torch.bfloat16, device=dev) if ctx.want_bf16 else None
        d_res = torch.empty_like(x) if ctx.has_res else None
        dgamma = torch.empty(c, dtype=torch.float32, device=dev) if gamma is not None else None
        dbeta = torch.empty(c, dtype=torch.float32, device=dev) if ctx.has_beta else None
        sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib().bevf_bn_train_backward(ptr(x), ptr(dy), ptr(y), ptr(gamma), ptr(mean), ptr(invstd), int(n), int(c),
                                               int(ctx.relu), ptr(sums), ptr(dx), ptr(dx_bf16), ptr(d_res), ptr(dgamma),
                                               ptr(dbeta), cur_stream(dev)))
        if dx_bf16 is not None:
            dx._bevf_bf16 = dx_bf16     # the upstream conv's backward reads its operand copy from here (no cast pass)
        return dx, dgamma, dbeta, d_res, None, None, None


def bn_act_train(feats, bn, residual=None, relu=True, want_bf16=False):
    """-> (y fp32 [n, C], y_bf16 or None)."""
    return _BnActTrain.apply(feats, bn.weight, bn.bias, residual, bn, relu, want_bf16)
