"""Per-launch CUDA-event times of the 21 gather-GEMMs of a config-A frame on the static plan (bf16), best of 5."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200.static_frontend import StaticFrontEnd  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = frontend.BEVFrontEnd(precision="bf16").to(dev).eval()
synthetic.init_encoder_weights(model.pts_middle_encoder, 0)
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
layers = plan.profile(reps=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
tot = sum(r["ms"] for r in layers)
fl = sum(r["flops"] for r in layers)
print("layer_us", [round(1e3 * r["ms"], 1) for r in layers])
print("gemm_ms %.4f  useful TFLOP/s %.1f  (frac of 1647.3: %.4f)  tag=%s" % (tot, fl / tot / 1e9, fl / tot / 1e9 / 1647.3,
                                                                           os.environ.get("TAG", "")))
