"""clock64 accounting of the TS gather-GEMM kernel per role (instrumented build: BEVFRONT_LIB=.../libbevfront_b200_prof.so)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200._lib import lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = frontend.BEVFrontEnd(precision="bf16").to(dev).eval()
pts = torch.from_numpy(synthetic.lidar_sweeps(seed=0)).to(dev)
feats, coords, _ = model.voxelize([pts])
buf = (ctypes.c_ulonglong * 64)()
with torch.no_grad():
    for _ in range(3):
        model.pts_middle_encoder(feats, coords, 1)
    lib().bevf_debug_ts_prof(buf)
    model.pts_middle_encoder(feats, coords, 1)
    lib().bevf_debug_ts_prof(buf)
names = ["b_tot", "b_halo", "b_lds", "b_empty", "b_st", "b_rot", "b_warps", "b_grp", "m_tot", "m_acc", "m_bfull", "m_full",
         "m_warps", "m_items", "e_wait", "e_work"]
for ci, c in enumerate((16, 32, 64, 128)):
    v = [buf[ci * 16 + i] for i in range(16)]
    bw, mw = max(v[6], 1), max(v[12], 1)
    print(f"cin {c}: builder warps {v[6]} MMA warps {v[12]} items {v[13]}")
    print("  builder per warp (kclk): " + ", ".join(f"{n}={v[i] / bw / 1e3:.1f}" for i, n in list(enumerate(names[:6])) + [(7, "b_grp")]))
    print("  mma per warp (kclk): " + ", ".join(f"{names[i]}={v[i] / mw / 1e3:.1f}" for i in (8, 9, 10, 11))
          + f", per item: tot={v[8] / max(v[13], 1):.0f} full_wait={v[11] / max(v[13], 1):.0f}")
    print(f"  epilogue warp0 per CTA (kclk): wait={v[14] / mw / 1e3:.1f} work={v[15] / mw / 1e3:.1f}")
