"""Voxelize + mean on the frame (320 k points) and on the stress sweep (886 k points) for `ncu -k regex:vox_`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer as vl  # noqa: E402

dev = torch.device("cuda", 0)
rng = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
for pts_np, vox, cap in ((synthetic.lidar_sweeps(seed=0), synthetic.NUSCENES_VOXEL, 160000),
                         (synthetic.stress_sweep(seed=0, point_range=rng), [0.05, 0.05, 0.2], 600000)):
    pts = torch.from_numpy(pts_np).to(dev)
    f = torch.empty((cap, 5), device=dev)
    c = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    s = torch.empty((cap,), dtype=torch.int32, device=dev)
    for _ in range(3):
        n = vl.voxelize_mean(pts, f, c, s, vox, rng, 10, cap)
    torch.cuda.synchronize()
    print(pts.shape[0], int(n.item()))
