"""Device time of voxelize_mean (sync-free C-ABI form, CUDA graph over distinct sweeps): config A and the stress sweep."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer as vl  # noqa: E402

dev = torch.device("cuda", 0)
for name, voxel, cap, gen, nsw in (("config A", [0.075, 0.075, 0.2], 160000, lambda s: synthetic.lidar_sweeps(seed=s), 20),
                                   ("stress", [0.05, 0.05, 0.2], 600000, lambda s: synthetic.stress_sweep(seed=s), 8)):
    pts = [torch.from_numpy(gen(s)).to(dev) for s in range(nsw)]   # 20 x 6.4 MB / 8 x 17.7 MB: > L2
    f = torch.empty((cap, 5), device=dev)
    c = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    sz = torch.empty((cap,), dtype=torch.int32, device=dev)
    rng = synthetic.NUSCENES_RANGE
    num = vl.voxelize_mean(pts[0], f, c, sz, voxel, rng, 10, cap)
    torch.cuda.synchronize()
    m, n = int(num.item()), int(pts[0].shape[0])
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [vl.voxelize_mean(p, f, c, sz, voxel, rng, 10, cap) for p in pts]
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (10 * nsw)
    by = 20 * n + m * (20 + 16 + 4)
    print("%s: %d points -> %d voxels: %.1f us, %.0f GB/s = %.3f of 6546.6" % (name, n, m, 1e3 * ms, by / ms / 1e6,
                                                                               by / ms / 1e6 / 6546.6))
