"""mmcv-style spelling of the voxel ops: what mmdet3d's own data preprocessor calls
(mmdet3d/models/data_preprocessors/voxelize.py:11-14, 17-97, 100-183, 186-326), on libbevfront_b200.

Differences from the project ops (ops/voxel/voxelize.py): voxel coordinates are **(z, y, x)**, voxel_size /
coors_range arrive as tensors, and the voxel count is returned through a 0-dim int64 tensor instead of a Python int.
`ext_module` below has the four entry points `mmcv.utils.ext_loader.load_ext('_ext', [...])` would provide, so
`Det3DDataPreprocessor.voxelize` and `deploy/export.py` work with `ext_module` swapped for this object.
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import voxel_layer


def _as_list(v, n):
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().tolist()
    v = [float(x) for x in v]
    assert len(v) == n, f"expected {n} values, got {len(v)}"
    return v


class _ExtModule:
    """The `_ext` functions of mmcv that mmdet3d binds (voxelize.py:11-14), zyx coordinate order."""

    @staticmethod
    def hard_voxelize_forward(points, voxel_size, coors_range, voxels, coors, num_points_per_voxel, voxel_num,
                              max_points=35, max_voxels=20000, NDim=3, deterministic=True):
        m = voxel_layer.hard_voxelize(points, voxels, coors, num_points_per_voxel, _as_list(voxel_size, 3),
                                      _as_list(coors_range, 6), max_points, max_voxels, NDim, deterministic)
        if m > 0:  # the library writes (x, y, z); mmcv's contract is (z, y, x)
            coors[:m] = coors[:m].flip(1)
        voxel_num.fill_(m)

    @staticmethod
    def dynamic_voxelize_forward(points, voxel_size, coors_range, coors, NDim=3):
        voxel_layer.dynamic_voxelize(points, coors, _as_list(voxel_size, 3), _as_list(coors_range, 6), NDim)
        # mmcv marks a point outside the range as (-1, -1, -1) and orders the columns (z, y, x)
        bad = (coors < 0).any(dim=1, keepdim=True)
        coors.copy_(torch.where(bad, torch.full_like(coors, -1), coors.flip(1)))

    dynamic_point_to_voxel_forward = staticmethod(voxel_layer.dynamic_point_to_voxel_forward)
    dynamic_point_to_voxel_backward = staticmethod(voxel_layer.dynamic_point_to_voxel_backward)


ext_module = _ExtModule()


class _Voxelization(Function):
    """data_preprocessors/voxelize.py:17-94."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000, deterministic=True):
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
            ext_module.dynamic_voxelize_forward(points, torch.tensor(voxel_size, dtype=torch.float),
                                                torch.tensor(coors_range, dtype=torch.float), coors, NDim=3)
            return coors
        voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
        coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
        num_points_per_voxel = points.new_zeros(size=(max_voxels,), dtype=torch.int)
        voxel_num = torch.zeros(size=(), dtype=torch.long)
        ext_module.hard_voxelize_forward(points, torch.tensor(voxel_size, dtype=torch.float),
                                         torch.tensor(coors_range, dtype=torch.float), voxels, coors,
                                         num_points_per_voxel, voxel_num, max_points=max_points,
                                         max_voxels=max_voxels, NDim=3, deterministic=deterministic)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class VoxelizationByGridShape(nn.Module):
    """data_preprocessors/voxelize.py:100-183: voxelization given either the voxel size or the grid shape."""

    def __init__(self, point_cloud_range, max_num_points, voxel_size=[], grid_shape=[], max_voxels=20000,
                 deterministic=True):
        super().__init__()
        if voxel_size and grid_shape:
            raise ValueError("voxel_size is mutually exclusive grid_shape")
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)
        self.deterministic = deterministic
        rng = torch.tensor(point_cloud_range, dtype=torch.float32)
        if voxel_size:
            self.voxel_size = voxel_size
            self.grid_shape = torch.round((rng[3:] - rng[:3]) / torch.tensor(voxel_size, dtype=torch.float32)
                                          ).long().tolist()
        elif grid_shape:
            self.grid_shape = grid_shape
            self.voxel_size = ((rng[3:] - rng[:3]) / (torch.tensor(grid_shape, dtype=torch.float32) - 1)).tolist()
        else:
            raise ValueError("must assign a value to voxel_size or grid_shape")

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels,
                            self.deterministic)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, grid_shape={self.grid_shape}, "
                f"point_cloud_range={self.point_cloud_range}, max_num_points={self.max_num_points}, "
                f"max_voxels={self.max_voxels}, deterministic={self.deterministic})")


class _DynamicScatter(Function):
    """data_preprocessors/voxelize.py:186-249 (adds `return_map` to the project version)."""

    @staticmethod
    def forward(ctx, feats, coors, reduce_type="max", return_map=False):
        voxel_feats, voxel_coors, point2voxel_map, voxel_points_count = \
            ext_module.dynamic_point_to_voxel_forward(feats, coors, reduce_type)
        ctx.reduce_type = reduce_type
        ctx.save_for_backward(feats, voxel_feats, point2voxel_map, voxel_points_count)
        ctx.mark_non_differentiable(voxel_coors)
        if return_map:
            ctx.mark_non_differentiable(point2voxel_map)
            return voxel_feats, voxel_coors, point2voxel_map
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None, grad_map=None):
        feats, voxel_feats, point2voxel_map, voxel_points_count = ctx.saved_tensors
        grad_feats = torch.zeros_like(feats)
        ext_module.dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats.contiguous(), feats, voxel_feats,
                                                   point2voxel_map, voxel_points_count, ctx.reduce_type)
        return grad_feats, None, None, None


dynamic_scatter_3d = _DynamicScatter.apply


class DynamicScatter3D(nn.Module):
    """data_preprocessors/voxelize.py:252-326: mean (average_points) or max reduction; batched coors are
    (batch, z, y, x) and reduced sample by sample."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points

    def forward_single(self, points, coors):
        return dynamic_scatter_3d(points.contiguous(), coors.contiguous(), "mean" if self.average_points else "max")

    def forward(self, points, coors):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        batch_size = int(coors[-1, 0]) + 1
        voxels, voxel_coors = [], []
        for i in range(batch_size):
            inds = torch.where(coors[:, 0] == i)
            voxel, voxel_coor = self.forward_single(points[inds], coors[inds][:, 1:])
            voxel_coors.append(nn.functional.pad(voxel_coor, (1, 0), mode="constant", value=i))
            voxels.append(voxel)
        return torch.cat(voxels, dim=0), torch.cat(voxel_coors, dim=0)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, "
                f"point_cloud_range={self.point_cloud_range}, average_points={self.average_points})")
