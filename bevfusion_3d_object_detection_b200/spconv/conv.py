"""SubMConv3d / SparseConv3d with the spconv-2.x constructor and parameter layout
(`SparseConvolution.__init__`, mirrored by the reference at projects/SparseConvolution/sparse_conv.py:288-360;
registered into mmengine MODELS under the same names as mmdet3d/models/layers/spconv/overwrite_spconv/
write_spconv2.py:21-38 does for the spconv classes).
"""
import math
import os

import torch
from torch import nn

from . import functional as Fsp
from .core import SparseConvTensor
from .modules import SparseModule

def lib_cin_pad(c):
    from .._lib import lib

    return int(lib().bevf_spconv_tc_cin_pad(int(c)))


_DEFAULT_PRECISION = os.environ.get("BEVFRONT_SPCONV_PRECISION", "fp32")


def set_default_precision(precision):
    """'fp32' (FFMA, 1e-5 parity with the fp32 reference) or 'bf16' (tcgen05 tensor cores, 2e-2)."""
    global _DEFAULT_PRECISION
    assert precision in ("fp32", "bf16")
    _DEFAULT_PRECISION = precision


def get_default_precision():
    return _DEFAULT_PRECISION


class _SparseConvFunction(torch.autograd.Function):
    """Forward and backward on libbevfront_b200 (training, SURVEY configs[2]).  The data gradient is the forward
    gather-GEMM over the inverse rulebook with transposed weights (same fp32 / tcgen05 kernels as the forward), the
    weight gradient a tiled fp32 outer-product reduction (csrc/spconv_bwd.cu); no host round trips."""

    @staticmethod
    def forward(ctx, features, weight, bias, pair_fwd, n_out, packed, precision, datas=None, features_bf16=None):
        cout, cin = weight.shape[0], weight.shape[-1]
        kv = weight.numel() // (cout * cin)
        if precision != "bf16" or features_bf16 is None or features_bf16.shape[1] != lib_cin_pad(cin):
            features_bf16 = None
        out, _ = Fsp.implicit_gemm(features, pair_fwd, n_out, packed, kv, cin, cout, precision=precision, bias=bias,
                                   features_bf16=features_bf16)
        ctx.features_bf16 = features_bf16   # operand copy written by the producer (fused BN): reused by the weight gradient
        ctx.save_for_backward(features, weight, pair_fwd)
        ctx.has_bias = bias is not None
        ctx.precision = precision
        ctx.n_out = n_out
        ctx.datas = datas     # the rulebook object: layers that share it share its inverse (built once per step)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, weight, pair_fwd = ctx.saved_tensors
        cout, cin = weight.shape[0], weight.shape[-1]
        kv = pair_fwd.shape[0]
        n_in, n_out = features.shape[0], ctx.n_out
        grad_out = grad_out.contiguous().float()
        g_feat = g_w = None
        bf16 = ctx.precision == "bf16" and n_in > 0 and n_out > 0
        tc = bf16 and Fsp.tc_supported(cout, cin)       # data gradient = a (Cout -> Cin) convolution
        tc_w = bf16 and Fsp.tc_supported(cin, cout)     # weight gradient: same channel pair as the forward
        # bf16 operand copy of the output gradient, shared by the data and the weight gradient
        g_bf16 = None
        if tc or tc_w:
            g_bf16 = getattr(grad_out, "_bevf_bf16", None)   # left by the fused BN backward: no cast pass
            if g_bf16 is None or g_bf16.shape != (n_out, lib_cin_pad(cout)):
                g_bf16 = Fsp.cast_features_bf16(grad_out, lib_cin_pad(cout))
        if ctx.needs_input_grad[0]:
            if n_in == 0 or n_out == 0:
                g_feat = torch.zeros_like(features)
            else:
                # W^T per tap in the parameter layout of a (Cout -> Cin) convolution: [Cin, kD, kH, kW, Cout]
                wt = weight.detach().reshape(cout, kv, cin).permute(2, 1, 0).contiguous().view(
                    cin, *weight.shape[1:-1], cout)
                packed_t = Fsp.pack_weight_bf16(wt) if tc else Fsp.pack_weight_f32(wt)
                pb = getattr(ctx.datas, "pair_bwd_cache", None) if ctx.datas is not None else None
                if pb is None or pb.shape[1] != n_in:
                    pb = Fsp.pair_bwd(pair_fwd, n_out, n_in)
                    if ctx.datas is not None:
                        ctx.datas.pair_bwd_cache = pb
                g_feat, _ = Fsp.implicit_gemm(None if tc else grad_out, pb, n_in, packed_t, kv, cout, cin,
                                              precision="bf16" if tc else "fp32", features_bf16=g_bf16,
                                              timing_tag="data_gradient")
        if ctx.needs_input_grad[1]:
            timing = Fsp.GEMM_TIMING
            if timing is not None:   # profiling pass: events around the launch, useful flops from the rulebook
                pairs = int((pair_fwd[:, :n_out] >= 0).sum().item())
                ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
                      2.0 * pairs * cin * cout, "weight_gradient"]
                ev[0].record()
            if tc_w:   # tcgen05, both operands MN-major (csrc/spconv_wgrad_tc.cu)
                f_bf16 = ctx.features_bf16
                if f_bf16 is None:
                    f_bf16 = Fsp.cast_features_bf16(features.detach(), lib_cin_pad(cin))
                g_w = Fsp.wgrad_bf16(f_bf16, g_bf16, pair_fwd, n_out, kv, cin, cout).view(weight.shape)
            else:
                g_w = Fsp.wgrad_f32(features, grad_out, pair_fwd, n_out, kv, cin, cout).view(weight.shape)
            if timing is not None:
                ev[1].record()
                timing.append(tuple(ev))
        g_b = grad_out.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return g_feat, g_w, g_b, None, None, None, None, None, None


class SparseConvolution(SparseModule):
    __constants__ = ["stride", "padding", "dilation", "groups", "bias", "subm", "inverse", "transposed",
                     "output_padding"]

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None, algo=None,
                 fp32_accum=None, record_voxel_count=False, act_type=None, act_alpha=0, act_beta=0, large_kernel_fast_algo=False,
                 name=None, precision=None):
        super().__init__()
        if ndim != 3:
            raise NotImplementedError("only 3-D sparse convolutions are on the BEV front-end hot path")
        if groups != 1:
            raise NotImplementedError("groups != 1 is not supported")
        if transposed or inverse:
            raise NotImplementedError("transposed / inverse sparse convolutions are not on the BEV front-end hot path")
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = list(Fsp._triple(kernel_size))
        self.stride = list(Fsp._triple(stride))
        self.padding = list(Fsp._triple(padding))
        self.dilation = list(Fsp._triple(dilation))
        self.output_padding = list(Fsp._triple(output_padding))
        self.conv1x1 = all(k == 1 for k in self.kernel_size)
        self.transposed, self.inverse, self.groups, self.subm = transposed, inverse, groups, subm
        self.indice_key = indice_key
        self.algo = algo
        self.fp32_accum = fp32_accum
        self.record_voxel_count = record_voxel_count
        self.name = name
        self.precision = precision      # None -> module default
        self.fuse_epilogue = True       # SparseSequential may fold eval BatchNorm1d + ReLU into this conv
        self.need_f32 = True            # False: on the tensor-core path write only the bf16 copy (the encoder sets
                                        # this on its inner layers; `.features` is then materialised on demand)
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self._packed = {}
        self.reset_parameters()

    def extra_repr(self):
        s = "{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}"
        if self.padding != [0] * 3:
            s += ", padding={padding}"
        if self.dilation != [1] * 3:
            s += ", dilation={dilation}"
        if self.bias is None:
            s += ", bias=False"
        if self.indice_key is not None:
            s += ", indice_key={indice_key}"
        return s.format(**self.__dict__)

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    # ---- weights re-packed for the kernels, cached per parameter version --------------------------------
    def _packed_weight(self, precision):
        key = (precision, self.weight._version, self.weight.data_ptr(), self.weight.device)
        hit = self._packed.get(precision)
        if hit is not None and hit[0] == key:
            return hit[1]
        packed = Fsp.pack_weight_f32(self.weight) if precision == "fp32" else Fsp.pack_weight_bf16(self.weight)
        self._packed[precision] = (key, packed)
        return packed

    def _resolve_precision(self):
        p = self.precision or _DEFAULT_PRECISION
        if p == "bf16" and not Fsp.tc_supported(self.in_channels, self.out_channels):
            raise RuntimeError(f"bf16 tensor-core path does not support {self.in_channels}->{self.out_channels} "
                               "channels; use precision='fp32' for this layer")
        return p

    def _rulebook(self, input):
        ksize, stride = tuple(self.kernel_size), tuple(self.stride)
        padding, dilation = tuple(self.padding), tuple(self.dilation)
        key = self.indice_key
        datas = input.find_indice_pair(key)
        if datas is not None:
            if self.subm:
                assert datas.subm and datas.ksize == ksize and datas.dilation == dilation, \
                    f"indice_key {key!r} is shared by layers with different geometry"
            else:
                assert not datas.subm and datas.matches(ksize, stride, padding, dilation, False), \
                    f"indice_key {key!r} is shared by layers with different geometry"
            return datas
        if self.subm and key is None:
            # the reference passes indice_key=None to every SubM conv of SparseBasicBlock
            # (middle_encoders/sparse_encoder.py:171,222-226) so spconv rebuilds the same rulebook 16 times;
            # equal geometry on the same sites gives the same rulebook, so cache it under a derived key
            auto = ("__subm__", ksize, dilation)
            datas = input.indice_dict.get(auto)
            if datas is not None and datas.out_indices.data_ptr() == input.indices.data_ptr():
                return datas
            datas = Fsp.get_indice_pairs(input, ksize, stride, padding, dilation, True)
            input.indice_dict[auto] = datas
            return datas
        datas = Fsp.get_indice_pairs(input, ksize, stride, padding, dilation, self.subm)
        if key is not None:
            input.indice_dict[key] = datas
        return datas

    def forward(self, input, bn_scale=None, bn_shift=None, residual=None, relu=False):
        assert isinstance(input, SparseConvTensor)
        in_ch = input._features.shape[1] if input._features is not None else self.in_channels
        assert in_ch == self.in_channels, "channel size mismatch"
        precision = self._resolve_precision()
        datas = self._rulebook(input)
        out = input.shadow_copy()
        out._bf16 = None
        if not self.subm:
            out.indices = datas.out_indices
            out.spatial_shape = list(datas.out_spatial_shape)
            out._index = datas.out_index
            out._sorted_rows = True
        n_out = datas.n_out
        needs_grad = torch.is_grad_enabled() and (self.weight.requires_grad or (
            input._features is not None and input._features.requires_grad))
        fused = bn_scale is not None or residual is not None or relu
        if needs_grad:
            feats = _SparseConvFunction.apply(input.features, self.weight, self.bias, datas.pair_fwd, n_out,
                                              self._packed_weight(precision), precision, datas, input._bf16)
            if fused:   # the fused epilogues are inference-only: same arithmetic as explicit (differentiable) torch ops
                if bn_scale is not None:
                    feats = feats * bn_scale + bn_shift
                if residual is not None:
                    feats = feats + residual.to(feats.dtype)
                if relu:
                    feats = torch.relu(feats)
        else:
            kv = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
            use_bf16_in = precision == "bf16" and input._bf16 is not None
            feats, fb = Fsp.implicit_gemm(None if use_bf16_in else input.features, datas.pair_fwd, n_out,
                                          self._packed_weight(precision), kv, self.in_channels, self.out_channels,
                                          precision=precision, bias=self.bias, bn_scale=bn_scale, bn_shift=bn_shift,
                                          residual=residual, relu=relu,
                                          features_bf16=input._bf16 if use_bf16_in else None,
                                          want_bf16=(precision == "bf16"), want_f32=self.need_f32)
            out._bf16 = fb
        out._features = feats
        return out


class SubMConv3d(SparseConvolution):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, large_kernel_fast_algo=False, name=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key, algo=algo, fp32_accum=fp32_accum, name=name, **kwargs)


class SparseConv3d(SparseConvolution):

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, algo=None, fp32_accum=None, record_voxel_count=False, large_kernel_fast_algo=False,
                 name=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key, algo=algo, fp32_accum=fp32_accum,
                         record_voxel_count=record_voxel_count, name=name, **kwargs)


def _off_path(name, ndim, what):
    """The other conv classes spconv exports and mmdet3d registers (write_spconv2.py:21-38): same names, but constructing
    one says plainly that it is not on the BEV front-end hot path instead of failing later in a registry lookup."""

    class _OffPath(SparseModule):
        def __init__(self, *args, **kwargs):
            super().__init__()
            raise NotImplementedError(f"{name}: {what} sparse convolutions are not on the BEVFusion front-end hot path "
                                      f"(only SubMConv3d / SparseConv3d are built; DESIGN.md 7)")

    _OffPath.__name__ = _OffPath.__qualname__ = name
    _OffPath.ndim = ndim
    return _OffPath


SparseConv2d = _off_path("SparseConv2d", 2, "2-D")
SparseConv4d = _off_path("SparseConv4d", 4, "4-D")
SubMConv2d = _off_path("SubMConv2d", 2, "2-D")
SubMConv4d = _off_path("SubMConv4d", 4, "4-D")
SparseConvTranspose2d = _off_path("SparseConvTranspose2d", 2, "transposed")
SparseConvTranspose3d = _off_path("SparseConvTranspose3d", 3, "transposed")
SparseInverseConv2d = _off_path("SparseInverseConv2d", 2, "inverse")
SparseInverseConv3d = _off_path("SparseInverseConv3d", 3, "inverse")
OFF_PATH_CLASSES = (SparseConv2d, SparseConv4d, SubMConv2d, SubMConv4d, SparseConvTranspose2d, SparseConvTranspose3d,
                    SparseInverseConv2d, SparseInverseConv3d)
