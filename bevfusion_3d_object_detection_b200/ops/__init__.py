"""Same public names as the reference's projects/BEVFusion/bevfusion/ops/__init__.py:1-4."""
from .bev_pool import BevPoolTables, bev_pool, bev_pool_fused
from .depth import depth_histogram, lidar_depth_image
from .voxel import DynamicScatter, Voxelization, dynamic_scatter, voxelization

__all__ = ["bev_pool", "Voxelization", "voxelization", "dynamic_scatter", "DynamicScatter", "bev_pool_fused",
           "BevPoolTables", "lidar_depth_image", "depth_histogram"]
