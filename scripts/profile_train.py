"""A few training steps (configs[2], batch 4) for an ncu launch list: prints launches per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bevfusion_3d_object_detection_b200 import synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200.training import TrainStep  # noqa: E402

dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["train"]
B = int(os.environ.get("TRAIN_BATCH", cfg["batch"]))
model = bench.build_model(cfg, "bf16", dev)
rig = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(cfg["n_cams"], cfg["image"], B).items()}
tables = model.set_calibration(rig)
b = bench.batches_of(bench.make_frames(cfg, B, seed0=100), B)[0]
pts = [torch.from_numpy(p).to(dev) for p in b["points"]]
depth, ctx = torch.from_numpy(b["depth"]).to(dev), torch.from_numpy(b["ctx"]).to(dev)
step = TrainStep(model, tables, lr=1e-5)
for _ in range(2):
    step(pts, depth, ctx)
torch.cuda.synchronize()
print("MARK")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step(pts, depth, ctx)
e1.record()
torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1))
