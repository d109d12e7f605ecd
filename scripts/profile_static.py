"""Three eager runs of the static plan on the bench workload (for an ncu launch list of exactly one frame:
`ncu -s 164 -c 82 ...` skips the first two frames)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200.static_frontend import StaticFrontEnd  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = frontend.BEVFrontEnd(precision="bf16").to(dev).eval()
rig = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(6, (256, 704), 1).items()}
tables = model.set_calibration(rig)
pts = torch.from_numpy(synthetic.lidar_sweeps(seed=0)).to(dev)
depth, ctx = [torch.from_numpy(a).to(dev) for a in synthetic.camera_features(6, 118, 80, (32, 88), 1, seed=0)]
plan = StaticFrontEnd(model, tables, dev, batch=1, max_points=int(pts.shape[0]) + 4096)
plan.load_inputs([pts], depth, ctx)
plan.run()
torch.cuda.synchronize()
for lv, n in zip(plan.levels, plan.counts()):
    lv.hint = n
torch.cuda.synchronize()
from bevfusion_3d_object_detection_b200._lib import lib  # noqa: E402

print("MARK")
n0 = lib().bevf_launch_count()
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
print("counts", plan.counts())
print("launches_per_frame", (lib().bevf_launch_count() - n0) // 3)
