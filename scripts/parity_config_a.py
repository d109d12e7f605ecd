"""Config A (BASELINE configs[1]) differential: both BEV maps of the static plan against the CPU oracle chain
(oracle/cpu_frontend: reference C++ hard_voxelize_cpu + gather-mm-scatter sparse encoder in fp32 + the
outer-product / index_add_ bev_pool) on the same frame and weights, for
    fp32 | bf16 with the residual stream in fp32 | bf16 with bf16 skip connections.
Prints max / mean error relative to the reference map's scale (max |ref|) and the GEMM time of each mode."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from oracle import cpu_frontend  # noqa: E402
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200.static_frontend import StaticFrontEnd  # noqa: E402

oracle.build()
dev = torch.device("cuda", 0)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pts_np = synthetic.lidar_sweeps(seed=seed)
depth_np, ctx_np = synthetic.camera_features(6, 118, 80, (32, 88), 1, seed=seed)
rig_np = synthetic.camera_rig(6, (256, 704), 1)
res = {}
ref = None
for mode in ("fp32", "bf16_resid_f32", "bf16_resid_bf16"):
    torch.manual_seed(0)
    model = frontend.BEVFrontEnd(precision="fp32" if mode == "fp32" else "bf16").to(dev).eval()
    synthetic.init_encoder_weights(model.pts_middle_encoder, seed=0)
    model.pts_middle_encoder.set_residual_f32(mode != "bf16_resid_bf16")
    rig = {k: torch.from_numpy(v).to(dev) for k, v in rig_np.items()}
    tables = model.set_calibration(rig)
    if ref is None:
        torch.set_num_threads(os.cpu_count() or 1)
        plan_cpu = cpu_frontend.encoder_plan(model.pts_middle_encoder)
        vcfg = frontend.NUSCENES_VOXELIZE_CFG
        with torch.no_grad():
            f, c, _ = cpu_frontend.cpu_voxelize_mean(pts_np, vcfg["voxel_size"], vcfg["point_cloud_range"],
                                                     vcfg["max_num_points"], vcfg["max_voxels"][1])
            lid_ref = cpu_frontend.cpu_sparse_encoder(plan_cpu, f, c.numpy(), [1440, 1440, 41], 1).numpy()
            vt = frontend.BaseViewTransform(**frontend.NUSCENES_VIEW_CFG)
            geom = vt.get_geometry(**{k: torch.from_numpy(v) for k, v in rig_np.items()})
            gf, kept, _, indices = vt.bev_pool_aux(geom)
            nx = [int(v) for v in vt.nx]
            cam_ref = cpu_frontend.cpu_bev_pool(torch.from_numpy(depth_np), torch.from_numpy(ctx_np), kept, indices, gf,
                                                1, 6, nx[2], nx[0], nx[1]).numpy()
        ref = (lid_ref, cam_ref)
    plan = StaticFrontEnd(model, tables, dev, batch=1, max_points=int(pts_np.shape[0]) + 4096)
    plan.load_inputs([torch.from_numpy(pts_np).to(dev)], torch.from_numpy(depth_np).to(dev),
                     torch.from_numpy(ctx_np).to(dev))
    lid, cam = plan.run()
    torch.cuda.synchronize()
    lid, cam = lid.cpu().numpy(), cam.cpu().numpy()
    for lv, n in zip(plan.levels, plan.counts()):
        lv.hint = n
    layers = plan.profile(reps=5)
    s_l, s_c = np.abs(ref[0]).max(), np.abs(ref[1]).max()
    nz = ref[0] != 0
    res[mode] = dict(
        lidar_max_rel=float(np.abs(lid - ref[0]).max() / s_l), lidar_mean_rel=float(np.abs(lid - ref[0]).mean() / s_l),
        lidar_mean_rel_nonzero=float(np.abs(lid - ref[0])[nz].mean() / np.abs(ref[0])[nz].mean()),
        lidar_scale=float(s_l), lidar_same_support=bool(((lid != 0) == nz).all()),
        cam_max_rel=float(np.abs(cam - ref[1]).max() / s_c), cam_scale=float(s_c),
        gemm_ms=float(sum(r["ms"] for r in layers)), counts=plan.counts(),
        layer_us=[round(1e3 * r["ms"], 1) for r in layers])
    print(mode, json.dumps(res[mode]), flush=True)
    del plan, model
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/parity_config_a_seed%d.json" % seed, "w"), indent=1)
