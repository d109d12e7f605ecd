"""View-transform geometry + pooling with the reference's method surface
(projects/BEVFusion/bevfusion/depth_lss.py:14-223: gen_dx_bx, BaseViewTransform.create_frustum /
get_geometry / bev_pool_aux / bev_pool / bev_pool_precomputed), plus the fused pooling entry points.

The dense conv stacks of DepthLSSTransform (dtransform / depthnet / downsample, cuDNN) are outside the hot
path and are not rebuilt here: a reference `DepthLSSTransform` keeps them and calls `bev_pool`
from `bevfusion_3d_object_detection_b200.ops` unchanged.
"""
from typing import Tuple

import torch
from torch import nn

from .ops import BevPoolTables, bev_pool, bev_pool_fused, depth_histogram, lidar_depth_image


def gen_dx_bx(xbound, ybound, zbound):
    """depth_lss.py:14-18: cell size, first cell centre, cell count per axis."""
    bounds = [xbound, ybound, zbound]
    dx = torch.Tensor([b[2] for b in bounds])
    bx = torch.Tensor([b[0] + b[2] / 2.0 for b in bounds])
    nx = torch.LongTensor([(b[1] - b[0]) / b[2] for b in bounds])
    return dx, bx, nx


class BaseViewTransform(nn.Module):

    def __init__(self, in_channels: int, out_channels: int, image_size: Tuple[int, int],
                 feature_size: Tuple[int, int], xbound, ybound, zbound, dbound) -> None:
        super().__init__()
        self.in_channels = in_channels
        self.image_size = image_size
        self.feature_size = feature_size
        self.xbound, self.ybound, self.zbound, self.dbound = xbound, ybound, zbound, dbound
        dx, bx, nx = gen_dx_bx(xbound, ybound, zbound)
        self.dx = nn.Parameter(dx, requires_grad=False)
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        self.C = out_channels
        self.frustum = self.create_frustum()
        self.D = self.frustum.shape[0]
        self.fp16_enabled = False
        self._tables = None

    def create_frustum(self):
        """depth_lss.py:53-66: (u, v, d) for every (depth bin, feature row, feature col)."""
        iH, iW = self.image_size
        fH, fW = self.feature_size
        ds = torch.arange(*self.dbound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
        D = ds.shape[0]
        us = torch.linspace(0, iW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
        vs = torch.linspace(0, iH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
        return nn.Parameter(torch.stack((us, vs, ds), -1), requires_grad=False)

    def get_geometry(self, camera2lidar_rots, camera2lidar_trans, intrins_inverse, post_rots_inverse, post_trans,
                     **kwargs):
        """depth_lss.py:68-112: frustum -> LiDAR-frame xyz, [B, N, D, fH, fW, 3]."""
        B, N, _ = camera2lidar_trans.shape
        pts = self.frustum - post_trans.view(B, N, 1, 1, 1, 3)
        pts = post_rots_inverse.view(B, N, 1, 1, 1, 3, 3).matmul(pts.unsqueeze(-1))
        pts = torch.cat((pts[..., :2, :] * pts[..., 2:3, :], pts[..., 2:3, :]), 5)
        combine = camera2lidar_rots.matmul(intrins_inverse)
        pts = combine.view(B, N, 1, 1, 1, 3, 3).matmul(pts).squeeze(-1)
        pts = pts + camera2lidar_trans.view(B, N, 1, 1, 1, 3)
        if "extra_rots" in kwargs:
            rot = kwargs["extra_rots"].view(B, 1, 1, 1, 1, 3, 3).repeat(1, N, 1, 1, 1, 1, 1)
            pts = rot.matmul(pts.unsqueeze(-1)).squeeze(-1)
        if "extra_trans" in kwargs:
            pts = pts + kwargs["extra_trans"].view(B, 1, 1, 1, 1, 3).repeat(1, N, 1, 1, 1, 1)
        return pts

    def get_cam_feats(self, x):
        """The conv stacks (dtransform / depthnet: dense cuDNN work) stay with the caller; the data path around them is
        BaseDepthTransform.get_cam_feats below."""
        raise NotImplementedError("BaseViewTransform has no depth network; see BaseDepthTransform.get_cam_feats")

    def bev_pool_aux(self, geom_feats):
        """depth_lss.py:118-176: quantise (truncation toward zero), drop out-of-grid points, rank, sort.
        -> (geom_feats[Nk,4] (x,y,z,b) sorted, kept[N'] bool, ranks[Nk], indices[Nk])"""
        B, N, D, H, W, C = geom_feats.shape
        assert C == 3
        nprime = B * N * D * H * W
        cells = ((geom_feats - (self.bx - self.dx / 2.0)) / self.dx).long().view(nprime, 3)
        batch_ix = torch.arange(B, device=cells.device, dtype=torch.long).repeat_interleave(nprime // B)
        cells = torch.cat((cells, batch_ix.view(-1, 1)), 1)
        kept = ((cells[:, 0] >= 0) & (cells[:, 0] < self.nx[0]) & (cells[:, 1] >= 0) & (cells[:, 1] < self.nx[1])
                & (cells[:, 2] >= 0) & (cells[:, 2] < self.nx[2]))
        cells = cells[kept]
        Dz, Wy = self.nx[2], self.nx[1]
        ranks = cells[:, 0] * (Wy * Dz * B) + cells[:, 1] * (Dz * B) + cells[:, 2] * B + cells[:, 3]
        indices = ranks.argsort(stable=True)  # the reference's argsort is unstable; ties fixed to frustum order
        return cells[indices], kept, ranks[indices], indices

    def bev_pool(self, x, geom_feats):
        """depth_lss.py:179-204 (boundary form: x is the materialised [B,N,D,fH,fW,C] frustum tensor)."""
        geom_feats, kept, ranks, indices = self.bev_pool_aux(geom_feats)
        return self.bev_pool_precomputed(x, geom_feats, kept, ranks, indices)

    def bev_pool_precomputed(self, x, geom_feats, kept, ranks, indices):
        """depth_lss.py:206-223."""
        B, N, D, H, W, C = x.shape
        x = x.reshape(B * N * D * H * W, C)[kept]
        assert x.shape[0] == geom_feats.shape[0]
        x = x[indices]
        x = bev_pool(x, geom_feats, ranks, B, self.nx[2], self.nx[0], self.nx[1], self.training)
        return torch.cat(x.unbind(dim=2), 1)  # collapse Z

    # ---- fused north-star path ----------------------------------------------------------------------
    def build_tables(self, geom, B=None, device_build=None):
        """Per-calibration tables for `pool_fused` from a geometry tensor [B,N,D,fH,fW,3].  CUDA geometry goes
        through the device-side builder (csrc/bev_tables.cu: no argsort, no host round trips but the final sizes);
        `device_build=False` (and CPU geometry) takes the reference's bev_pool_aux route."""
        B = geom.shape[0] if B is None else B
        if device_build is None:
            device_build = geom.is_cuda
        if device_build:
            self._tables = BevPoolTables.from_geometry(geom, B, self.bx.tolist(), self.dx.tolist(), self.nx.tolist())
            return self._tables
        geom_feats, kept, ranks, indices = self.bev_pool_aux(geom)
        Bg, N, D, fH, fW, _ = geom.shape
        self._tables = BevPoolTables(geom_feats, kept, ranks, indices, B, int(self.nx[2]), int(self.nx[0]),
                                     int(self.nx[1]), frustum_shape=(Bg * N, D, fH, fW))
        return self._tables

    def pool_fused(self, depth, ctx, tables=None):
        """depth [B*N, D, fH, fW] (softmax), ctx [B*N, C, fH, fW] -> [B, C*nz, nx, ny].  Same result as
        get_cam_feats' outer product (depth_lss.py:723-725) followed by bev_pool(), without the 638 MB
        frustum tensor or any of its copies."""
        return bev_pool_fused(depth, ctx, self._tables if tables is None else tables)


class BaseDepthTransform(BaseViewTransform):
    """The data-path half of depth_lss.py:333-520 (BaseDepthTransform.forward) and :632-661 (get_cam_feats'
    histogram): LiDAR points -> sparse depth image -> per-cell depth-bin counts / distribution.  The conv stacks
    that consume them stay with the caller."""

    def lidar_depth(self, points, lidar2image, img_aug_matrix, lidar_aug_matrix, lidar_aug_matrix_inverse=None):
        """-> depth [B, N, 1, H, W] (depth_lss.py:366-420)"""
        return lidar_depth_image(points, lidar2image, img_aug_matrix, lidar_aug_matrix, self.image_size,
                                 lidar_aug_matrix_inverse)

    def depth_distribution(self, depth):
        """depth [B, N, 1, H, W] -> (gt_depth_distr, counts_3d), each [B, N, fH, fW, D] (depth_lss.py:632-661)"""
        counts, distr = depth_histogram(depth, self.feature_size, self.dbound, self.D)
        return distr, counts

    def get_cam_feats(self, x, d, dtransform, depthnet):
        """DepthLSSTransform.get_cam_feats (depth_lss.py:617-725) with the frustum outer product left to the fused pooling:
        x [B, N, C_img, fH, fW] image features, d [B, N, 1, H, W] LiDAR depth image (`lidar_depth`); `dtransform` and
        `depthnet` are the caller's conv stacks (the reference's own modules: dense cuDNN work, out of scope here).
        -> (depth [B*N, D, fH, fW], ctx [B*N, C, fH, fW], est_depth_distr, gt_depth_distr, counts_3d): depth / ctx go to
        `pool_fused`; the other three are what the reference returns for its depth loss.  In training the softmax depth
        is lifted to the LiDAR distribution with the correction detached (:702-706)."""
        B, N, C_img, fH, fW = x.shape
        gt_depth_distr, counts_3d = self.depth_distribution(d)                       # :632-661 (depth_prep.cu)
        d = dtransform(d.view(B * N, *d.shape[2:]))
        y = depthnet(torch.cat([d, x.view(B * N, C_img, fH, fW)], dim=1))            # [B*N, D + C, fH, fW]
        depth = y[:, :self.D].softmax(dim=1)
        est_depth_distr = depth.permute(0, 2, 3, 1).reshape(B, N, fH, fW, self.D)
        if self.training:
            depth_aux = gt_depth_distr.view(B * N, fH, fW, self.D).permute(0, 3, 1, 2)
            depth = depth + (torch.maximum(depth_aux, depth) - depth).detach()
        return depth, y[:, self.D:self.D + self.C], est_depth_distr, gt_depth_distr, counts_3d
