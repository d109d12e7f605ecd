"""Voxelization module / function with the reference's surface
(projects/BEVFusion/bevfusion/ops/voxel/voxelize.py:10-152): same constructor arguments, attributes,
return values ((x, y, z) int32 coordinates) and `__repr__`, running on libbevfront_b200.
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from . import voxel_layer


class _Voxelization(Function):
    """voxelize.py:10-74.  hard voxelization when max_points/max_voxels are set, dynamic when either is -1."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000, deterministic=True):
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros((points.size(0), 3), dtype=torch.int32)
            voxel_layer.dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
            return coors
        # reference contract: the caller allocates zero-filled outputs (voxelize.py:51-53)
        voxels = points.new_zeros((max_voxels, max_points, points.size(1)))
        coors = points.new_zeros((max_voxels, 3), dtype=torch.int32)
        num_points_per_voxel = points.new_zeros((max_voxels,), dtype=torch.int32)
        voxel_num = voxel_layer.hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range,
                                              max_points, max_voxels, 3, deterministic)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    """voxelize.py:80-152.  `max_voxels` is (training, testing); `deterministic` is accepted for config
    compatibility (this implementation is always deterministic and never O(N^2))."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)
        self.deterministic = deterministic

        pc_range = torch.tensor(point_cloud_range, dtype=torch.float32)
        vsize = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = torch.round((pc_range[3:] - pc_range[:3]) / vsize).long()
        self.grid_size = grid_size
        self.pcd_shape = [*grid_size[:2], 1]  # (x-len, y-len, 1): the fork keeps xyz order (voxelize.py:119-121)

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels,
                            self.deterministic)

    def forward_mean(self, points_list):
        """Extension: BEVFusion.voxelize with voxelize_reduce=True for a whole batch
        (bevfusion.py:227-255) -> (feats[M,C], coords[M,4] (b,x,y,z), sizes[M]) with ONE host sync at the
        end instead of one per sample and no padded [max_voxels, max_points, C] tensor."""
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        dev = points_list[0].device
        c = points_list[0].size(1)
        cap = sum(min(max_voxels, p.size(0)) for p in points_list)
        feats = torch.empty((cap, c), dtype=torch.float32, device=dev)
        coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
        sizes = torch.empty((cap,), dtype=torch.int32, device=dev)
        offset = torch.zeros(1, dtype=torch.int32, device=dev)
        for k, pts in enumerate(points_list):
            voxel_layer.voxelize_mean(pts.contiguous(), feats, coords, sizes, self.voxel_size, self.point_cloud_range,
                                      self.max_num_points, max_voxels, batch_idx=k, row_offset=offset)
        m = int(offset.item())
        return feats[:m], coords[:m], sizes[:m]

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, "
                f"point_cloud_range={self.point_cloud_range}, max_num_points={self.max_num_points}, "
                f"max_voxels={self.max_voxels}, deterministic={self.deterministic})")
