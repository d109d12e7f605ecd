"""Per-layer device time of the sparse encoder's GEMM launches on the bench workload (CUDA events), plus the
rulebook / index kernels.  Run on the GPU box: python scripts/profile_layers.py [bf16|fp32]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200.spconv import functional as Fsp  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = frontend.BEVFrontEnd(precision=precision).to(dev).eval()
pts = torch.from_numpy(synthetic.lidar_sweeps(seed=0)).to(dev)
feats, coords, _ = model.voxelize([pts])
enc = model.pts_middle_encoder
from bevfusion_3d_object_detection_b200._lib import lib  # noqa: E402
if len(sys.argv) > 2:
    lib().bevf_spconv_tc_variant(int(sys.argv[2]))   # 0 = SS kernel, 1 = default, 2 = TS wherever instantiated
print("tc variant", lib().bevf_spconv_tc_variant(-1))
with torch.no_grad():
    for _ in range(3):
        enc(feats, coords, 1)
    rows = None
    for rep in range(5):
        Fsp.GEMM_TIMING = []
        enc(feats, coords, 1)
        torch.cuda.synchronize()
        cur = [(ev[0].elapsed_time(ev[1]), ev[2]) for ev in Fsp.GEMM_TIMING]
        rows = cur if rows is None else [(min(r[0], c[0]), c[1]) for r, c in zip(rows, cur)]
    Fsp.GEMM_TIMING = None
convs = [m for m in enc.modules() if hasattr(m, "indice_key") and hasattr(m, "kernel_size")]
out = []
for m, (ms, fl) in zip(convs, rows):
    out.append(dict(cin=m.in_channels, cout=m.out_channels, subm=bool(m.subm), ms=ms, gflop=fl / 1e9,
                    tflops=fl / ms / 1e9))
    print(f"{m.in_channels:4d}->{m.out_channels:4d} subm={int(m.subm)} {ms*1e3:8.1f} us  {fl/1e9:7.2f} GF  {fl/ms/1e9:7.1f} TF/s")
print("total gemm ms", sum(r[0] for r in rows))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/layers_{precision}.json", "w"), indent=1)
