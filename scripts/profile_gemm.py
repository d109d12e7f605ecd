"""Four passes of the sparse encoder on the bench workload (module path, bf16) for an `ncu --set full` capture of the
gather-GEMM kernels: `ncu -k regex:spconv_t -s <3 passes> -c <1 pass> ...`.  argv[1] = tc variant (default 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200._lib import lib  # noqa: E402

if len(sys.argv) > 1:
    lib().bevf_spconv_tc_variant(int(sys.argv[1]))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = frontend.BEVFrontEnd(precision="bf16").to(dev).eval()
pts = torch.from_numpy(synthetic.lidar_sweeps(seed=0)).to(dev)
feats, coords, _ = model.voxelize([pts])
with torch.no_grad():
    for _ in range(4):
        model.pts_middle_encoder(feats, coords, 1)
torch.cuda.synchronize()
print("ok")
