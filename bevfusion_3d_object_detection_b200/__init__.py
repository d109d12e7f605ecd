"""B200-native BEV front end for BEVFusion (voxelization, bev_pool, sparse 3-D convolution).

Hand-written sm_100a CUDA behind a C ABI (include/bevfront_b200.h, built into lib/libbevfront_b200.so);
this package is the host-side mirror of the reference's operator surface.  There is no CPU fallback.
"""
__version__ = "0.1.0"
