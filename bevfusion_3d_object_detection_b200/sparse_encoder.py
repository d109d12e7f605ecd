"""LiDAR sparse encoder with the reference's class / config surface:
  make_sparse_convmodule, SparseBasicBlock   mmdet3d/models/layers/sparse_block.py:94-224
  SparseEncoder.make_encoder_layers           mmdet3d/models/middle_encoders/sparse_encoder.py:165-241
  BEVFusionSparseEncoder                      projects/BEVFusion/bevfusion/sparse_encoder.py:13-156
Module names (and therefore state-dict keys) follow the reference so its checkpoints load unchanged.
"""
import os

from torch import nn

from . import registry
from .spconv import SparseConvTensor, SparseModule, SparseSequential, SubMConv3d, SparseConv3d
from .spconv.modules import _fold_bn, bn_is_foldable, wants_grad

registry._LOCAL.setdefault("SubMConv3d", SubMConv3d)
registry._LOCAL.setdefault("SparseConv3d", SparseConv3d)


# bf16 path: keep the residual stream (block inputs / outputs) in fp32?  Measured at config A against the fp32 CPU
# oracle chain (scripts/parity_config_a.py, profiles/README.md r2a): max error of the encoder output relative to its
# scale 1.2e-2 with bf16 skip connections, 0.8e-2 with the fp32 stream -- both inside north_star's 2e-2 -- and the fp32
# stream costs 13 % GEMM time (1.16 vs 1.02 ms: 9 layers write fp32 rows and lose the TMA-staged epilogue).  Default
# off; env BEVFRONT_RESIDUAL_F32=1 or BEVFusionSparseEncoder.set_residual_f32(True) turns it on.
RESIDUAL_F32_DEFAULT = os.environ.get("BEVFRONT_RESIDUAL_F32", "0") == "1"


def replace_feature(out, new_features):
    return out.replace_feature(new_features)


class SparseBasicBlock(SparseModule):
    """sparse_block.py:94-154 (mmdet BasicBlock wiring): SubM conv - BN - ReLU - SubM conv - BN - (+identity) -
    ReLU.  Both convs are 3x3x3, padding 1, bias=False.  In eval mode each conv runs with its BatchNorm (and the
    residual add + ReLU) folded into the GEMM epilogue."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, indice_key=None, conv_cfg=None, norm_cfg=None):
        super().__init__()
        conv_cfg = dict(type="SubMConv3d") if conv_cfg is None else dict(conv_cfg)
        conv_cfg.setdefault("indice_key", indice_key)
        norm_cfg = dict(type="BN1d") if norm_cfg is None else norm_cfg
        self.norm1_name, norm1 = registry.build_norm_layer(norm_cfg, planes, postfix=1)
        self.norm2_name, norm2 = registry.build_norm_layer(norm_cfg, planes, postfix=2)
        self.conv1 = registry.build_conv_layer(conv_cfg, inplanes, planes, 3, stride=stride, padding=1, dilation=1,
                                               bias=False)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = registry.build_conv_layer(conv_cfg, planes, planes, 3, padding=1, bias=False)
        self.add_module(self.norm2_name, norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    @property
    def norm1(self):
        return getattr(self, self.norm1_name)

    @property
    def norm2(self):
        return getattr(self, self.norm2_name)

    def forward(self, x):
        if (self.downsample is None and bn_is_foldable(self.norm1) and bn_is_foldable(self.norm2)
                and not wants_grad(x, self)):
            s1, b1 = _fold_bn(self.norm1)
            s2, b2 = _fold_bn(self.norm2)
            # the skip connection reads whichever copy of the block input exists (fp32, else the bf16 operand copy)
            identity = x._features if x._features is not None else x._bf16
            out = self.conv1(x, bn_scale=s1, bn_shift=b1, relu=True)
            return self.conv2(out, bn_scale=s2, bn_shift=b2, residual=identity, relu=True)
        identity = x.features
        assert x.features.dim() == 2, f"x.features.dim()={x.features.dim()}"
        from .spconv import bn_train

        if self.downsample is None and bn_train.usable(self.norm1, identity) and bn_train.usable(self.norm2, identity):
            # training: each conv is followed by ONE fused op (batch-statistics BN, ReLU, skip connection, bf16 copy)
            bf16 = self.conv1._resolve_precision() == "bf16"
            out = self.conv1(x)
            y, yb = bn_train.bn_act_train(out.features, self.norm1, None, True, want_bf16=bf16)
            out = replace_feature(out, y)
            out._bf16 = yb
            out = self.conv2(out)
            y, yb = bn_train.bn_act_train(out.features, self.norm2, identity, True, want_bf16=bf16)
            out = replace_feature(out, y)
            out._bf16 = yb
            return out
        out = self.conv1(x)
        out = replace_feature(out, self.norm1(out.features))
        out = replace_feature(out, self.relu(out.features))
        out = self.conv2(out)
        out = replace_feature(out, self.norm2(out.features))
        if self.downsample is not None:
            identity = self.downsample(x).features
        out = replace_feature(out, out.features + identity)
        out = replace_feature(out, self.relu(out.features))
        return out


def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0,
                           conv_type="SubMConv3d", norm_cfg=None, order=("conv", "norm", "act"), **kwargs):
    """sparse_block.py:157-224: SparseSequential of conv / norm / act in `order` (conv has bias=False)."""
    assert isinstance(order, tuple) and len(order) <= 3
    assert set(order) | {"conv", "norm", "act"} == {"conv", "norm", "act"}
    conv_cfg = dict(type=conv_type, indice_key=indice_key)
    norm_cfg = dict(type="BN1d") if norm_cfg is None else norm_cfg
    layers = []
    for layer in order:
        if layer == "conv":
            layers.append(registry.build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                                    padding=padding, bias=False))
        elif layer == "norm":
            layers.append(registry.build_norm_layer(norm_cfg, out_channels)[1])
        elif layer == "act":
            layers.append(nn.ReLU(inplace=True))
    return SparseSequential(*layers)


class SparseEncoder(nn.Module):
    """Only the layer builder of mmdet3d's SparseEncoder is on the BEVFusion path (the project subclass overrides
    __init__ and forward)."""

    def make_encoder_layers(self, make_block, norm_cfg, in_channels, block_type="conv_module",
                            conv_cfg=dict(type="SubMConv3d")):
        assert block_type in ["conv_module", "basicblock"]
        self.encoder_layers = SparseSequential()
        for i, blocks in enumerate(self.encoder_channels):
            blocks_list = []
            for j, out_channels in enumerate(tuple(blocks)):
                padding = tuple(self.encoder_paddings[i])[j]
                first_of_stage = i != 0 and j == 0
                last_of_stage = j == len(blocks) - 1 and i != len(self.encoder_channels) - 1
                if block_type == "conv_module" and first_of_stage:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg, stride=2,
                                                  padding=padding, indice_key=f"spconv{i + 1}",
                                                  conv_type="SparseConv3d"))
                elif block_type == "basicblock" and last_of_stage:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg, stride=2,
                                                  padding=padding, indice_key=f"spconv{i + 1}",
                                                  conv_type="SparseConv3d"))
                elif block_type == "basicblock":
                    blocks_list.append(SparseBasicBlock(out_channels, out_channels, norm_cfg=norm_cfg,
                                                        conv_cfg=conv_cfg))
                else:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg, padding=padding,
                                                  indice_key=f"subm{i + 1}", conv_type="SubMConv3d"))
                in_channels = out_channels
            self.encoder_layers.add_module(f"encoder_layer{i + 1}", SparseSequential(*blocks_list))
        return out_channels


class BEVFusionSparseEncoder(SparseEncoder):
    """projects/BEVFusion/bevfusion/sparse_encoder.py:13-156.  Spatial order is (X, Y, Z); the output is
    dense() -> [N, C, X, Y, Z] -> permute(0, 1, 4, 2, 3) -> view(N, C*Z, X, Y)."""

    def __init__(self, in_channels, sparse_shape, order=("conv", "norm", "act"),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), base_channels=16, output_channels=128,
                 encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)), block_type="conv_module",
                 return_middle_feats=False):
        super().__init__()
        assert block_type in ["conv_module", "basicblock"]
        assert isinstance(order, tuple) and len(order) == 3 and set(order) == {"conv", "norm", "act"}
        self.sparse_shape = sparse_shape
        self.in_channels = in_channels
        self.order = order
        self.base_channels = base_channels
        self.output_channels = output_channels
        self.encoder_channels = encoder_channels
        self.encoder_paddings = encoder_paddings
        self.stage_num = len(self.encoder_channels)
        self.fp16_enabled = False
        self.return_middle_feats = return_middle_feats
        # the voxelizer emits voxels in first-appearance (point) order; the encoder's output is a dense map, so the
        # row order inside it is free: ascending cell order makes every conv's input rows one contiguous range per
        # kernel slab (what the tensor-core kernel's halo staging wants) and the level-0 index needs no permutation
        self.sort_voxels = True
        first_order = ("conv",) if self.order[0] != "conv" else ("conv", "norm", "act")
        self.conv_input = make_sparse_convmodule(in_channels, self.base_channels, 3, norm_cfg=norm_cfg, padding=1,
                                                 indice_key="subm1", conv_type="SubMConv3d", order=first_order)
        encoder_out_channels = self.make_encoder_layers(make_sparse_convmodule, norm_cfg, self.base_channels,
                                                        block_type=block_type)
        self.conv_out = make_sparse_convmodule(encoder_out_channels, self.output_channels, kernel_size=(1, 1, 3),
                                               stride=(1, 1, 2), norm_cfg=norm_cfg, padding=0,
                                               indice_key="spconv_down2", conv_type="SparseConv3d")
        self.set_residual_f32(RESIDUAL_F32_DEFAULT)

    def set_residual_f32(self, flag):
        """Which layers of the tensor-core (bf16) path also write an fp32 copy of their output.  Always: the last conv
        (dense() consumes it).  flag=True: additionally every layer whose output is a SparseBasicBlock's input -- the
        residual stream (conv_input, the second conv of every block, the strided convs) -- so the skip connections add
        in fp32 and bf16 rounding does not compound along the 8 residual blocks; the convolution operands stay bf16.
        flag=False: inner layers hand only their bf16 operand copy on (fastest; skip connections read bf16)."""
        self.residual_f32 = bool(flag)
        for m in self.modules():
            if hasattr(m, "need_f32") and hasattr(m, "indice_key"):
                m.need_f32 = m is self.conv_out[0]
        if not self.residual_f32:
            return
        feeds_block = None      # the conv whose output the next module reads
        for m in [self.conv_input] + [blk for stage in self.encoder_layers for blk in stage]:
            if isinstance(m, SparseBasicBlock):
                if feeds_block is not None:
                    feeds_block.need_f32 = True
                feeds_block = m.conv2
            else:
                feeds_block = m[0]

    def forward(self, voxel_features, coors, batch_size):
        """voxel_features [M, C] fp32, coors [M, 4] (batch, x, y, z) -> [B, C_out * Z_out, X_out, Y_out]."""
        coors = coors.int()
        x = SparseConvTensor(voxel_features, coors, self.sparse_shape, batch_size)
        if self.sort_voxels and coors.shape[0] > 0:
            x = _sorted_by_cell(x)
        x = self.conv_input(x)
        encode_features = []
        for encoder_layer in self.encoder_layers:
            x = encoder_layer(x)
            encode_features.append(x)
        out = self.conv_out(encode_features[-1])
        spatial_features = out.dense_bev()  # == dense().permute(0,1,4,2,3).contiguous().view(N, C*D, H, W)
        if self.return_middle_feats:
            return spatial_features, encode_features
        return spatial_features


def _sorted_by_cell(x):
    """Same sparse tensor with rows in ascending linear (batch, x, y, z) order; reuses the coordinate index."""
    from .spconv.core import CoordIndex

    index = x.coord_index()
    n = x.indices.shape[0]
    if index.perm is None:
        return x
    perm = index.perm[:n].long()
    out = SparseConvTensor(x.features.index_select(0, perm), x.indices.index_select(0, perm), x.spatial_shape,
                           x.batch_size)
    out._index = CoordIndex(out.indices, x.batch_size, x.spatial_shape, mem=index.mem)
    out._sorted_rows = True
    return out


NUSCENES_ENCODER_CFG = dict(  # projects/BEVFusion/configs/nuscenes/bevfusion_lidar_voxel0075...py:56-65
    in_channels=5, sparse_shape=[1440, 1440, 41], order=("conv", "norm", "act"),
    norm_cfg=dict(type="BN1d", eps=0.001, momentum=0.01),
    encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
    encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, (1, 1, 0)), (0, 0)), block_type="basicblock")
