"""Device time of the fused bev_pool forward (CUDA graph over distinct inputs > L2), batch 1 and batch 4, config A."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
for B in (1, 4):
    vt = frontend.BaseViewTransform(**frontend.NUSCENES_VIEW_CFG).to(dev)
    rig = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(6, (256, 704), B).items()}
    tab = vt.build_tables(vt.get_geometry(**rig))
    n_in = 8 if B == 1 else 2
    ins = []
    for i in range(n_in):
        d, c = synthetic.camera_features(6, 118, 80, (32, 88), B, seed=i)
        ins.append((torch.from_numpy(d).to(dev), torch.from_numpy(c).to(dev)))
    with torch.no_grad():
        for i in range(n_in):
            vt.pool_fused(*ins[i], tab)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = [vt.pool_fused(*ins[i], tab) for i in range(n_in)]
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (10 * n_in)
    by = 4 * B * 6 * 32 * 88 * (118 + 80) + 4 * tab.nk + 8 * tab.n_intervals + 4 * 80 * B * 360 * 360
    print("batch %d: %.1f us per call, %.0f GB/s = %.3f of 6546.6 (runs %d, intervals %d)" % (B, 1e3 * ms, by / ms / 1e6,
          by / ms / 1e6 / 6546.6, tab.n_runs, tab.n_intervals))
