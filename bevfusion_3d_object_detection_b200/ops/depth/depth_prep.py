"""LiDAR depth image and its per-feature-cell histogram: the two index-arithmetic blocks upstream of bev_pool in the
reference's camera branch.

  lidar_depth_image   BaseDepthTransform.forward's per-sample loop, projects/BEVFusion/bevfusion/depth_lss.py:372-420
  depth_histogram     the counts_3d / gt_depth_distr block of DepthLSSTransform.get_cam_feats, depth_lss.py:632-661

CUDA only (csrc/depth_prep.cu through the C ABI); CPU tensors raise.
"""
import torch

from ... import _lib


def _dev_f32(t, device):
    return torch.as_tensor(t, dtype=torch.float32, device=device).contiguous()


@torch.no_grad()
def lidar_depth_image(points, lidar2image, img_aug_matrix, lidar_aug_matrix, image_size, lidar_aug_matrix_inverse=None,
                      out=None):
    """points: list of B CUDA tensors [N_b, >=3]; lidar2image / img_aug_matrix [B, N, 4, 4]; lidar_aug_matrix
    [B, 4, 4] -> depth [B, N, 1, H, W] float32.

    Differences from the reference, both deliberate: the points are NOT modified (the reference subtracts the
    augmentation translation from the caller's tensor in place, depth_lss.py:379), and a pixel hit by several points
    takes the one with the largest index (what a sequential scatter_ leaves; the reference's CUDA scatter_ keeps an
    arbitrary one, see its own note at :406-414)."""
    if not points[0].is_cuda:
        raise RuntimeError("lidar_depth_image: CUDA tensors only (there is no CPU path)")
    device = points[0].device
    B = len(points)
    H, W = image_size
    lidar2image = _dev_f32(lidar2image, device)
    img_aug_matrix = _dev_f32(img_aug_matrix, device)
    lidar_aug_matrix = _dev_f32(lidar_aug_matrix, device)
    if lidar_aug_matrix_inverse is None:
        lidar_aug_matrix_inverse = torch.inverse(lidar_aug_matrix)  # depth_lss.py:364-365
    inv_rot = _dev_f32(lidar_aug_matrix_inverse, device)[:, :3, :3].contiguous()
    trans = lidar_aug_matrix[:, :3, 3].contiguous()
    N = lidar2image.shape[1]
    if out is None:
        out = torch.empty(B, N, 1, H, W, dtype=torch.float32, device=device)
    owner = torch.empty(N * H * W, dtype=torch.int32, device=device)
    lib = _lib.lib()
    for b in range(B):
        p = points[b]
        if p.dtype != torch.float32 or not p.is_contiguous():
            p = p.float().contiguous()
        _lib.check(lib.bevf_lidar_depth_image(_lib.ptr(p), p.shape[0], p.shape[1] if p.dim() == 2 else 3,
                                              _lib.ptr(trans[b]), _lib.ptr(inv_rot[b]), _lib.ptr(lidar2image[b]),
                                              _lib.ptr(img_aug_matrix[b]), N, H, W, _lib.ptr(out[b]), _lib.ptr(owner),
                                              _lib.cur_stream(device)))
    return out


@torch.no_grad()
def depth_histogram(depth, feature_size, dbound, n_bins=None):
    """depth [B, N, 1, H, W] (or [BN, 1, H, W]) -> (counts_3d, gt_depth_distr), each [B, N, fH, fW, D]
    (or [BN, fH, fW, D]); D = len(arange(*dbound)) unless given."""
    if not depth.is_cuda:
        raise RuntimeError("depth_histogram: CUDA tensors only (there is no CPU path)")
    import ctypes

    lead = depth.shape[:-3]
    H, W = depth.shape[-2:]
    fH, fW = feature_size
    D = int(n_bins) if n_bins is not None else torch.arange(*dbound, dtype=torch.float).shape[0]
    d = depth.reshape(-1, H, W)
    if d.dtype != torch.float32 or not d.is_contiguous():
        d = d.float().contiguous()
    bn = d.shape[0]
    counts = torch.empty(bn, fH, fW, D, dtype=torch.float32, device=depth.device)
    distr = torch.empty_like(counts)
    f = ctypes.c_float
    _lib.check(_lib.lib().bevf_depth_histogram(_lib.ptr(d), bn, H, W, fH, fW, D, f(dbound[0]), f(dbound[1]),
                                               f(dbound[2]), _lib.ptr(counts), _lib.ptr(distr),
                                               _lib.cur_stream(depth.device)))
    return counts.view(*lead, fH, fW, D), distr.view(*lead, fH, fW, D)
