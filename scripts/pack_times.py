"""Device time of the two pack kernels of the "rows" output alone (nothing else running): destination in pinned host
memory (zero-copy stores over the host link) and in device memory, plus a cudaMemcpy of the same bytes for comparison."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bevfusion_3d_object_detection_b200 import frontend, synthetic  # noqa: E402
from bevfusion_3d_object_detection_b200._lib import check, cur_stream, lib, ptr  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = frontend.BEVFrontEnd(precision="bf16").to(dev).eval()
rig = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(6, (256, 704), 1).items()}
tables = model.set_calibration(rig)
pts = torch.from_numpy(synthetic.lidar_sweeps(seed=0)).pin_memory()
depth, ctx = [torch.from_numpy(a).pin_memory() for a in synthetic.camera_features(6, 118, 80, (32, 88), 1, seed=0)]
pipe = frontend.HostPipeline(model, tables, dev, depth=2, batch=1, max_points=int(pts.shape[0]) + 4096,
                             example=([pts.to(dev)], depth.to(dev), ctx.to(dev)))
r = pipe.result(pipe.submit([pts], depth, ctx, output="rows"))
plan = pipe.plans[0]
nbytes = r.nbytes()
L, t = lib(), plan.tables
maxb = int(sys.argv[1]) if len(sys.argv) > 1 else 0


def run(dst_l, dst_c, reps=20):
    st = cur_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(2):
        e0.record()
        for _ in range(reps):
            check(L.bevf_pack_sparse_rows(ptr(plan.last_rows), ptr(plan.last_level.indices), plan.last_level.cap,
                                          ptr(plan.last_level.n_dev), r.c_lidar, ctypes.c_void_p(dst_l.data_ptr()),
                                          ctypes.c_size_t(dst_l.numel()), maxb, st))
            check(L.bevf_pack_cells(ptr(plan.cam_bev), ptr(t.interval_cell), t.n_intervals, int(dst_c.shape[0]), t.nz,
                                    t.nx * t.ny, ctypes.c_void_p(dst_c.data_ptr()), int(dst_c.shape[1]), maxb, st))
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_h = run(r.lidar_raw, r.cam)
ms_d = run(torch.empty_like(r.lidar_raw, device=dev), torch.empty_like(r.cam, device=dev))
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(2):
    e0.record()
    for _ in range(20):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
ms_c = e0.elapsed_time(e1) / 20
print("rows output: %.1f MB per frame; pack -> pinned host %.3f ms (%.1f GB/s), pack -> device %.3f ms, cudaMemcpy D2H of the "
      "same bytes %.3f ms (%.1f GB/s)" % (nbytes / 1e6, ms_h, nbytes / ms_h / 1e6, ms_d, ms_c, nbytes / ms_c / 1e6))
