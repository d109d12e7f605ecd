"""Seeded synthetic nuScenes-shaped inputs: LiDAR sweeps and a pinhole camera rig.

numpy only, deterministic for a given seed; used by tests, smoke() and bench.py.  There is no network for
real datasets, so every number reported by this repo is on these inputs (bench.py says "synthetic").
"""
import numpy as np

NUSCENES_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
NUSCENES_VOXEL = [0.075, 0.075, 0.2]


def lidar_sweeps(n_sweeps=10, beams=32, azimuth=1084, elev_deg=(-30.67, 10.67), seed=0, point_range=None,
                 dims=5, sensor_height=1.84, shuffle=True):
    """-> float32 [N, dims] (x, y, z, intensity, dt)[:dims].  10 sweeps of the default 32-beam scanner give
    ~320 k in-range points; sweeps are offset by the ego motion (-0.5 m per sweep in x, dt = 0.05 s)."""
    rng = np.random.default_rng(seed)
    point_range = NUSCENES_RANGE if point_range is None else point_range
    el = np.deg2rad(np.linspace(elev_deg[0], elev_deg[1], beams))[:, None]
    az = np.linspace(-np.pi, np.pi, azimuth, endpoint=False)[None, :]
    out = []
    for s in range(n_sweeps):
        obstacle = rng.lognormal(mean=3.0, sigma=0.6, size=(1, azimuth))
        with np.errstate(divide="ignore"):
            ground = np.where(el < 0, sensor_height / np.maximum(np.sin(-el), 1e-6), np.inf)
        r = np.minimum(ground, obstacle / np.maximum(np.cos(el), 1e-3))
        r = r * (1.0 + rng.normal(0.0, 0.002, size=r.shape))
        keep = (r > 1.0) & (r < 100.0)
        x = r * np.cos(el) * np.cos(az) - 0.5 * s
        y = r * np.cos(el) * np.sin(az)
        z = r * np.sin(el)
        inten = rng.uniform(0, 255, size=r.shape)
        dt = np.full(r.shape, 0.05 * s)
        pts = np.stack([x, y, z, inten, dt], -1)[keep]
        out.append(pts)
    pts = np.concatenate(out, 0).astype(np.float32)
    lo, hi = np.asarray(point_range[:3], np.float32), np.asarray(point_range[3:], np.float32)
    inside = np.all((pts[:, :3] >= lo) & (pts[:, :3] < hi), axis=1)  # PointsRangeFilter
    pts = pts[inside]
    if shuffle:
        pts = pts[rng.permutation(pts.shape[0])]
    return np.ascontiguousarray(pts[:, :dims])


def stress_sweep(seed=0, point_range=None, dims=5):
    """128-beam, 8192-azimuth single sweep (~0.9 M in-range points)."""
    return lidar_sweeps(n_sweeps=1, beams=128, azimuth=8192, elev_deg=(-25.0, 15.0), seed=seed,
                        point_range=point_range, dims=dims)


def _rot_z(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float64)


def camera_rig(n_cams=6, image_size=(256, 704), batch=1, src_size=(900, 1600), resize=0.48):
    """Pinhole rig -> dict of float32 arrays shaped like the arguments of BaseViewTransform.get_geometry
    (depth_lss.py:68-112): camera2lidar_rots [B,N,3,3], camera2lidar_trans [B,N,3], intrins_inverse [B,N,3,3],
    post_rots_inverse [B,N,3,3], post_trans [B,N,3].  Cameras at yaw {0,-55,55,180,-110,110} deg, mounted
    1.5 m ahead of / above the LiDAR origin, K = (1266, 1266, 816, 491) at 1600x900, image aug = resize 0.48 +
    centre crop to `image_size`, LiDAR aug = identity."""
    yaws = np.deg2rad([0.0, -55.0, 55.0, 180.0, -110.0, 110.0])[:n_cams]
    # camera frame: x right, y down, z forward  ->  lidar frame: x forward, y left, z up
    cam2ego = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], np.float64)
    K = np.array([[1266.0, 0, 816.0], [0, 1266.0, 491.0], [0, 0, 1]], np.float64)
    fh, fw = image_size
    new_h, new_w = src_size[0] * resize, src_size[1] * resize
    crop_h, crop_w = new_h - fh, (new_w - fw) / 2.0
    post_rot = np.eye(3) * resize
    post_rot[2, 2] = 1.0
    post_tran = np.array([-crop_w, -crop_h, 0.0])
    rots, trans = [], []
    for yaw in yaws:
        R = _rot_z(yaw) @ cam2ego
        rots.append(R)
        trans.append(_rot_z(yaw) @ np.array([1.5, 0.0, 1.5]))
    rots = np.stack(rots)
    trans = np.stack(trans)

    def tile(a):
        return np.ascontiguousarray(np.broadcast_to(a, (batch,) + a.shape)).astype(np.float32)

    n = len(yaws)
    return dict(
        camera2lidar_rots=tile(rots), camera2lidar_trans=tile(trans),
        intrins_inverse=tile(np.broadcast_to(np.linalg.inv(K), (n, 3, 3))),
        post_rots_inverse=tile(np.broadcast_to(np.linalg.inv(post_rot), (n, 3, 3))),
        post_trans=tile(np.broadcast_to(post_tran, (n, 3))))


def camera_features(n_cams=6, D=118, C=80, feature_size=(32, 88), batch=1, seed=0):
    """depth = softmax(N(0,1)) over D, context = N(0,1): ([B*N, D, fH, fW], [B*N, C, fH, fW]) float32."""
    rng = np.random.default_rng(seed + 1000)
    fh, fw = feature_size
    logits = rng.standard_normal((batch * n_cams, D, fh, fw)).astype(np.float32)
    logits -= logits.max(1, keepdims=True)
    e = np.exp(logits)
    depth = (e / e.sum(1, keepdims=True)).astype(np.float32)
    ctx = rng.standard_normal((batch * n_cams, C, fh, fw)).astype(np.float32)
    return depth, ctx


def camera_matrices(rig, lidar_yaw=0.0, lidar_scale=1.0, lidar_trans=(0.0, 0.0, 0.0)):
    """The 4x4 form of a `camera_rig` as BaseDepthTransform.forward takes it (depth_lss.py:333-362):
    lidar2image [B,N,4,4] = K @ inv(camera2lidar), img_aug_matrix [B,N,4,4], lidar_aug_matrix [B,4,4]
    (rotation about z by `lidar_yaw`, scale, translation) -- float32."""
    B, N = rig["camera2lidar_trans"].shape[:2]
    l2i = np.zeros((B, N, 4, 4), np.float64)
    aug = np.zeros((B, N, 4, 4), np.float64)
    for b in range(B):
        for n in range(N):
            c2l = np.eye(4)
            c2l[:3, :3] = rig["camera2lidar_rots"][b, n]
            c2l[:3, 3] = rig["camera2lidar_trans"][b, n]
            K = np.eye(4)
            K[:3, :3] = np.linalg.inv(rig["intrins_inverse"][b, n].astype(np.float64))
            l2i[b, n] = K @ np.linalg.inv(c2l)
            aug[b, n] = np.eye(4)
            aug[b, n, :3, :3] = np.linalg.inv(rig["post_rots_inverse"][b, n].astype(np.float64))
            aug[b, n, :3, 3] = rig["post_trans"][b, n]
    laug = np.tile(np.eye(4), (B, 1, 1))
    laug[:, :3, :3] = _rot_z(lidar_yaw) * lidar_scale
    laug[:, :3, 3] = np.asarray(lidar_trans, np.float64)
    return l2i.astype(np.float32), aug.astype(np.float32), laug.astype(np.float32)


def init_encoder_weights(enc, seed=0):
    """Seeded weights for a BEVFusionSparseEncoder-shaped module (there are no checkpoints offline): conv kernels
    ~ N(0, 2 / (kv * Cin)) so that activations keep O(1) magnitude through the 21 layers (torch's default
    kaiming_uniform(a=sqrt(5)) shrinks them by ~3x per layer and the output would measure nothing), BatchNorm1d
    affine parameters and running statistics away from the identity so that the folded epilogues are exercised."""
    import torch
    from torch import nn

    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in enc.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            elif hasattr(m, "indice_key") and hasattr(m, "weight"):
                fan = m.weight.shape[-1] * (m.weight.numel() // (m.weight.shape[0] * m.weight.shape[-1]))
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan) ** 0.5)
    return enc
