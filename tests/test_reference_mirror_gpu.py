"""The reference's OWN tests for the sparse-conv rows of the hot path, restated on this library with the same fixtures
(tests/test_models/test_layers/test_spconv/test_spconv_module.py: test_SparseBasicBlock :15-49,
test_make_sparse_convmodule :52-106; tests/test_models/test_middle_encoders/test_sparse_encoders.py:9-29).

The reference asserts module structure and output SHAPES only (spconv is a third-party dependency).  Here the
same calls must pass the same assertions, and the VALUES are additionally checked against the CPU oracle (rulebook +
gather-GEMM restated in oracle/bevfront_oracle.c, then torch's own BatchNorm1d / ReLU on the CPU)."""
import numpy as np
import pytest
import torch

from bevfusion_3d_object_detection_b200 import spconv
from bevfusion_3d_object_detection_b200.sparse_encoder import (BEVFusionSparseEncoder, SparseBasicBlock,
                                                               make_sparse_convmodule)

pytestmark = pytest.mark.gpu

# fixtures of test_spconv_module.py:18-27 / :58-67, verbatim values
VOXEL_FEATURES = [[6.56126, 0.9648336, -1.7339306, 0.315],
                  [6.8162713, -2.480431, -1.3616394, 0.36],
                  [11.643568, -4.744306, -1.3580885, 0.16],
                  [23.482342, 6.5036807, 0.5806964, 0.35]]
COORDINATES = [[0, 12, 819, 131], [0, 16, 750, 136], [1, 16, 705, 232], [1, 35, 930, 469]]
SPATIAL_SHAPE = [41, 1600, 1408]
BATCH = 2


def _input():
    f = torch.tensor(VOXEL_FEATURES, dtype=torch.float32).cuda()
    c = torch.tensor(COORDINATES, dtype=torch.int32).cuda()
    return spconv.SparseConvTensor(f, c, SPATIAL_SHAPE, BATCH)


def _oracle_subm(oracle_mod, feats, conv):
    idx = np.asarray(COORDINATES, np.int32)
    _, pair, _ = oracle_mod.spconv_rulebook(idx, SPATIAL_SHAPE, 3, 1, 1, 1, True)
    return oracle_mod.spconv_gemm(np.ascontiguousarray(feats, np.float32), conv.weight.detach().cpu().numpy(), pair)


def _bn_train_cpu(x, bn):
    ref = torch.nn.BatchNorm1d(bn.num_features, eps=bn.eps, momentum=bn.momentum)
    ref.load_state_dict({k: v.detach().cpu() for k, v in bn.state_dict().items()})
    ref.train()
    return ref(torch.from_numpy(x)).detach().numpy()


def test_SparseBasicBlock(oracle_mod):
    torch.manual_seed(0)
    x = _input()
    block = SparseBasicBlock(4, 4, conv_cfg=dict(type="SubMConv3d", indice_key="subm1"),
                             norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01)).cuda()
    for m in (block.conv1, block.conv2):
        m.precision = "fp32"
    # the reference's assertions (:37-49)
    assert isinstance(block.conv1, spconv.SubMConv3d)
    assert block.conv1.in_channels == 4
    assert block.conv1.out_channels == 4
    assert isinstance(block.conv2, spconv.SubMConv3d)
    assert block.conv2.out_channels == 4
    assert block.bn1.eps == 1e-3
    assert block.bn1.momentum == 0.01
    bn1_before = {k: v.detach().cpu().clone() for k, v in block.bn1.state_dict().items()}
    bn2_before = {k: v.detach().cpu().clone() for k, v in block.bn2.state_dict().items()}
    out = block(x)                                    # train mode, as in the reference test
    assert out.features.shape == torch.Size([4, 4])
    # values: conv - BN(batch statistics) - ReLU - conv - BN - (+identity) - ReLU, sparse_block.py:137-154
    feats = np.asarray(VOXEL_FEATURES, np.float32)

    class _Bn:   # the BatchNorm state the forward pass started from
        def __init__(self, bn, sd):
            self.num_features, self.eps, self.momentum, self._sd = bn.num_features, bn.eps, bn.momentum, sd

        def state_dict(self):
            return self._sd

    y = _oracle_subm(oracle_mod, feats, block.conv1)
    y = np.maximum(_bn_train_cpu(y, _Bn(block.bn1, bn1_before)), 0.0)
    y = _oracle_subm(oracle_mod, y, block.conv2)
    y = np.maximum(_bn_train_cpu(y, _Bn(block.bn2, bn2_before)) + feats, 0.0)
    got = out.features.detach().cpu().numpy()
    assert np.abs(got - y).max() <= 1e-4 * max(1.0, np.abs(y).max()), np.abs(got - y).max()


def test_make_sparse_convmodule(oracle_mod):
    torch.manual_seed(0)
    x = _input()
    block0 = make_sparse_convmodule(4, 16, 3, "test0", stride=1, padding=0, conv_type="SubMConv3d",
                                    norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01),
                                    order=("conv", "norm", "act")).cuda()
    block0[0].precision = "fp32"
    # the reference's assertions (:81-91)
    assert isinstance(block0[0], spconv.SubMConv3d)
    assert block0[0].in_channels == 4
    assert block0[0].out_channels == 16
    assert isinstance(block0[1], torch.nn.BatchNorm1d)
    assert block0[1].eps == 0.001
    assert block0[1].momentum == 0.01
    assert isinstance(block0[2], torch.nn.ReLU)
    bn_before = {k: v.detach().cpu().clone() for k, v in block0[1].state_dict().items()}
    out = block0(x)
    assert out.features.shape == torch.Size([4, 16])
    ref_bn = torch.nn.BatchNorm1d(16, eps=1e-3, momentum=0.01)
    ref_bn.load_state_dict(bn_before)
    ref_bn.train()
    y = _oracle_subm(oracle_mod, np.asarray(VOXEL_FEATURES, np.float32), block0[0])
    y = np.maximum(ref_bn(torch.from_numpy(y)).detach().numpy(), 0.0)
    got = out.features.detach().cpu().numpy()
    assert np.abs(got - y).max() <= 1e-4 * max(1.0, np.abs(y).max()), np.abs(got - y).max()
    # second half of the reference test (:93-106): order ('norm', 'act', 'conv') with an inverse conv.  Inverse /
    # transposed sparse convs are not on the BEVFusion path: the builder must say so, not build a wrong layer.
    with pytest.raises((NotImplementedError, KeyError)):
        make_sparse_convmodule(4, 16, 3, "test1", stride=1, padding=0, conv_type="SparseInverseConv3d",
                               norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), order=("norm", "act", "conv"))
    # the same order with a conv that IS on the path
    block1 = make_sparse_convmodule(4, 16, 3, "test1", stride=1, padding=0, conv_type="SubMConv3d",
                                    norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), order=("norm", "act", "conv"))
    assert isinstance(block1[0], torch.nn.BatchNorm1d)
    assert isinstance(block1[1], torch.nn.ReLU)
    assert isinstance(block1[2], spconv.SubMConv3d)


def test_sparse_encoder_shape_at_reference_test_size():
    """test_sparse_encoders.py:9-29 builds mmdet3d's SparseEncoder on 207 842 voxels / batch 4 and checks the dense output
    shape.  The class on the BEVFusion path is its subclass BEVFusionSparseEncoder (same layer builder, sparse_shape in
    (x, y, z) order, BEV tail [N, C * D, H, W]); the same size goes through it here.  The reference draws coordinates
    with torch.randint(0, 4): 256 distinct cells for 207 842 rows -- duplicates, whose treatment is hash-order dependent
    in spconv; distinct cells are drawn instead."""
    rng = np.random.default_rng(0)
    shape, batch, n = [1024, 1024, 41], 4, 207842
    cells = batch * shape[0] * shape[1] * shape[2]
    lin = rng.choice(cells, size=n, replace=False)
    coors = np.stack([lin // (shape[0] * shape[1] * shape[2]), (lin // (shape[1] * shape[2])) % shape[0],
                      (lin // shape[2]) % shape[1], lin % shape[2]], 1).astype(np.int32)
    enc = BEVFusionSparseEncoder(in_channels=5, sparse_shape=shape, order=("conv", "norm", "act"),
                                 encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                                 encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, (1, 1, 0)), (0, 0)),
                                 block_type="basicblock").cuda().eval()
    with torch.no_grad():
        ret = enc(torch.rand([n, 5]).cuda(), torch.from_numpy(coors).cuda(), batch)
    assert ret.shape == torch.Size([4, 256, 128, 128])
    assert torch.isfinite(ret).all()
