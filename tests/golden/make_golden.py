"""Generate the committed golden fixtures by running the REAL reference (/root/reference) in this container.

    python tests/golden/make_golden.py

Sources of truth (nothing is copied from the reference; its modules are imported by path, see
oracle/ref_loader.py, and its C++ CPU ops are compiled into oracle/_ref by oracle/build_ref.py):
  voxel_numba_*.npz     mmdet3d/models/task_modules/voxel/voxel_generator.py points_to_voxel (numba)
  voxel_ref_cpp_*.npz   projects/BEVFusion/bevfusion/ops/voxel/src/voxelization_cpu.cpp hard_voxelize_cpu /
                        dynamic_voxelize_cpu ("asis" on a cubic grid, "fixed" = lookup-table shape fix)
  view_geometry.npz     projects/BEVFusion/bevfusion/depth_lss.py BaseViewTransform.get_geometry + bev_pool_aux
  quick_cumsum.npz      projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py QuickCumsum (pure torch)
  depth_prep.npz        depth_lss.py BaseDepthTransform.forward (LiDAR depth image, :372-420) and
                        DepthLSSTransform.get_cam_feats (counts_3d / gt_depth_distr, :632-661), run on CPU
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_loader  # noqa: E402
from bevfusion_3d_object_detection_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def voxel_numba():
    vg = ref_loader.voxel_generator()
    # (1) the reference's own known-answer test, tests/.../test_voxel_generator.py:7-20
    np.random.seed(0)
    points = np.random.uniform(0, 4, (20, 3))
    gen = vg.VoxelGenerator([5, 5, 1], [0, 0, 0, 20, 40, 4], 5)
    voxels, coors, npv = gen.generate(points)
    np.savez_compressed(os.path.join(OUT, "voxel_numba_reftest.npz"), points=points, voxels=voxels, coors=coors,
                        npv=npv, voxel_size=[5, 5, 1], coors_range=[0, 0, 0, 20, 40, 4], max_points=5,
                        max_voxels=20000)
    # (2) a small nuScenes-like sweep, xyz order (reverse_index=False), voxel cap NOT hit and hit
    pts = synthetic.lidar_sweeps(n_sweeps=2, beams=16, azimuth=180, seed=3)
    for tag, max_voxels in (("nocap", 20000), ("cap", 1500)):
        v, c, n = vg.points_to_voxel(pts, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10, False, max_voxels)
        np.savez_compressed(os.path.join(OUT, f"voxel_numba_sweep_{tag}.npz"), points=pts, voxels=v, coors=c, npv=n,
                            voxel_size=synthetic.NUSCENES_VOXEL, coors_range=synthetic.NUSCENES_RANGE,
                            max_points=10, max_voxels=max_voxels)
    # (3) coarse voxels so max_points overflows often
    v, c, n = vg.points_to_voxel(pts, [2.0, 2.0, 8.0], synthetic.NUSCENES_RANGE, 3, False, 20000)
    np.savez_compressed(os.path.join(OUT, "voxel_numba_sweep_coarse.npz"), points=pts, voxels=v, coors=c, npv=n,
                        voxel_size=[2.0, 2.0, 8.0], coors_range=synthetic.NUSCENES_RANGE, max_points=3,
                        max_voxels=20000)


def voxel_ref_cpp():
    build_ref.build()
    rng = np.random.default_rng(7)
    for name, vs, cr, mp, mv in (
            ("asis", [0.5, 0.5, 0.5], [0, 0, 0, 16, 16, 16], 4, 3000),        # cubic grid: safe as shipped
            ("fixed", [0.5, 0.25, 1.0], [-8, -4, -2, 8, 4, 2], 5, 900)):      # 32 x 32 x 4, cap hit
        mod = build_ref.load_ref("ref_voxel_" + name)
        lo, hi = np.array(cr[:3], np.float32), np.array(cr[3:], np.float32)
        pts = (rng.uniform(-0.1, 1.1, (6000, 4)) * np.append(hi - lo, 1.0) + np.append(lo, 0.0)).astype(np.float32)
        p = torch.from_numpy(pts)
        voxels = p.new_zeros((mv, mp, 4))
        coors = p.new_zeros((mv, 3), dtype=torch.int32)
        npv = p.new_zeros((mv,), dtype=torch.int32)
        m = mod.hard_voxelize(p, voxels, coors, npv, vs, cr, mp, mv, 3, True)
        dyn = p.new_zeros((pts.shape[0], 3), dtype=torch.int32)
        mod.dynamic_voxelize(p, dyn, vs, cr, 3)
        np.savez_compressed(os.path.join(OUT, f"voxel_ref_cpp_{name}.npz"), points=pts, voxels=voxels[:m].numpy(),
                            coors=coors[:m].numpy(), npv=npv[:m].numpy(), dyn_coors_cpu=dyn.numpy(), voxel_size=vs,
                            coors_range=cr, max_points=mp, max_voxels=mv)


def view_geometry():
    dl = ref_loader.depth_lss()
    vt = dl.BaseViewTransform(8, 6, (64, 96), (4, 6), [-12.0, 12.0, 0.75], [-12.0, 12.0, 0.75], [-10.0, 10.0, 20.0],
                              [1.0, 15.0, 1.0])
    rig = {k: torch.from_numpy(v) for k, v in synthetic.camera_rig(n_cams=3, image_size=(64, 96), batch=2,
                                                                  src_size=(200, 300), resize=0.4).items()}
    extra_rots = torch.eye(3).repeat(2, 1, 1)
    extra_rots[1] = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    extra_trans = torch.tensor([[0.0, 0.0, 0.0], [0.5, -0.25, 0.0]])
    geom = vt.get_geometry(**rig, extra_rots=extra_rots, extra_trans=extra_trans)
    geom_feats, kept, ranks, indices = vt.bev_pool_aux(geom)
    np.savez_compressed(os.path.join(OUT, "view_geometry.npz"), geom=geom.numpy(), geom_feats=geom_feats.numpy(),
                        kept=kept.numpy(), ranks=ranks.numpy(), extra_rots=extra_rots.numpy(),
                        extra_trans=extra_trans.numpy(), frustum=vt.frustum.detach().numpy(),
                        dx=vt.dx.detach().numpy(), bx=vt.bx.detach().numpy(), nx=vt.nx.detach().numpy(),
                        **{"rig_" + k: v.numpy() for k, v in rig.items()})


def quick_cumsum():
    mod, _ = ref_loader.bev_pool_py()
    g = torch.Generator().manual_seed(5)
    n, c = 500, 8
    cells = torch.randint(0, 40, (n,), generator=g)
    ranks, order = torch.sort(cells, stable=True)
    x = torch.randn(n, c, generator=g, dtype=torch.float64)
    geom = torch.stack([ranks // 8, ranks % 8, torch.zeros_like(ranks), torch.zeros_like(ranks)], 1)
    pooled, pooled_geom = mod.QuickCumsum.apply(x, geom, ranks)
    np.savez_compressed(os.path.join(OUT, "quick_cumsum.npz"), x=x.numpy(), geom=geom.numpy(), ranks=ranks.numpy(),
                        pooled=pooled.numpy(), pooled_geom=pooled_geom.numpy())


def depth_prep():
    dl = ref_loader.depth_lss()
    H, W, fH, fW = 64, 176, 8, 22
    dbound = [1.0, 30.0, 0.5]
    torch.manual_seed(0)
    vt = dl.DepthLSSTransform(8, 6, (H, W), (fH, fW), [-12.0, 12.0, 0.75], [-12.0, 12.0, 0.75], [-10.0, 10.0, 20.0],
                              dbound)
    vt.eval()
    B, N = 2, 3
    rig = synthetic.camera_rig(n_cams=N, image_size=(H, W), batch=B, src_size=(900, 1600), resize=0.12)
    l2i, iaug, laug = synthetic.camera_matrices(rig, lidar_yaw=0.1, lidar_scale=1.05, lidar_trans=(0.5, -0.2, 0.1))
    laug[1] = np.eye(4, dtype=np.float32)  # second sample: no LiDAR augmentation
    points = [synthetic.lidar_sweeps(n_sweeps=1, beams=32, azimuth=720, seed=1 + b)[:, :3].copy() for b in range(B)]
    # nothing on this path multiplies by lidar_aug's rotation, only by its inverse: augment the points so they
    # project sensibly
    for b in range(B):
        points[b][:, :3] = points[b][:, :3] @ laug[b, :3, :3].T + laug[b, :3, 3]

    class Captured(Exception):
        pass

    def grab(img, depth):
        raise Captured(depth)

    img = torch.randn(B, N, 8, fH, fW)
    eye = torch.eye(4).repeat(B, N, 1, 1)
    vt.get_cam_feats = grab
    try:
        with torch.no_grad():
            # the reference subtracts lidar_aug's translation from the caller's points in place (:379): pass copies
            vt.forward(img, [torch.from_numpy(p.copy()) for p in points], torch.from_numpy(l2i), eye.clone(),
                       eye.clone(), torch.from_numpy(iaug), torch.from_numpy(laug), None, None, None, None, None)
        raise AssertionError("get_cam_feats was not reached")
    except Captured as e:
        depth = e.args[0]
    del vt.get_cam_feats
    with torch.no_grad():
        _, _, gt_depth_distr, counts_3d = vt.get_cam_feats(img, depth.clone())
    np.savez_compressed(os.path.join(OUT, "depth_prep.npz"), points0=points[0], points1=points[1], lidar2image=l2i,
                        img_aug_matrix=iaug, lidar_aug_matrix=laug, depth=depth.numpy(),
                        counts_3d=counts_3d.numpy(), gt_depth_distr=gt_depth_distr.numpy(), image_size=[H, W],
                        feature_size=[fH, fW], dbound=dbound, D=vt.D)


if __name__ == "__main__":
    assert ref_loader.available(), "needs /root/reference"
    voxel_numba()
    voxel_ref_cpp()
    view_geometry()
    quick_cumsum()
    depth_prep()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
