"""One pass of voxelize_mean, fused bev_pool and the rulebook kernels on the bench workload (for ncu --set full)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = frontend.BEVFrontEnd(precision="bf16").to(dev).eval()
rig = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(6, (256, 704), 1).items()}
tables = model.set_calibration(rig)
pts = torch.from_numpy(synthetic.lidar_sweeps(seed=0)).to(dev)
depth, ctx = [torch.from_numpy(a).to(dev) for a in synthetic.camera_features(6, 118, 80, (32, 88), 1, seed=0)]
with torch.no_grad():
    for _ in range(3):
        feats, coords, _ = model.voxelize([pts])
        model.extract_img_bev(depth, ctx, tables)
        model.pts_middle_encoder(feats, coords, 1)
torch.cuda.synchronize()
print("ok")
