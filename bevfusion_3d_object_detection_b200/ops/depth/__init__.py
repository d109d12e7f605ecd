from .depth_prep import depth_histogram, lidar_depth_image

__all__ = ["lidar_depth_image", "depth_histogram"]
