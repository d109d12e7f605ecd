"""The BEV feature-construction front end as one object: LiDAR points -> voxelize -> sparse encoder -> BEV map,
and camera (depth, context) -> fused bev_pool -> BEV map.  It wires the three operators exactly where the
reference's detector does:

  BEVFusion.voxelize            projects/BEVFusion/bevfusion/bevfusion.py:227-255   (hard voxelize + mean + batch pad)
  BEVFusion.extract_pts_feat    projects/BEVFusion/bevfusion/bevfusion.py:196-225   (-> pts_middle_encoder)
  DepthLSSTransform.forward     projects/BEVFusion/bevfusion/depth_lss.py:699-725, 179-204 (outer product + bev_pool)

The dense image backbone / depthnet / fuser / head are outside the hot path: the camera branch starts
from the depthnet's softmax depth and context feature maps.
"""
import torch
from torch import nn

from . import synthetic
from .ops import Voxelization
from .sparse_encoder import NUSCENES_ENCODER_CFG, BEVFusionSparseEncoder
from .view_transform import BaseViewTransform

NUSCENES_VOXELIZE_CFG = dict(  # configs/nuscenes/bevfusion_lidar_voxel0075_second_secfpn_8xb4-cyclic-20e_nus-3d.py:49-54
    max_num_points=10, point_cloud_range=synthetic.NUSCENES_RANGE, voxel_size=synthetic.NUSCENES_VOXEL,
    max_voxels=(120000, 160000))
NUSCENES_VIEW_CFG = dict(  # configs/nuscenes/bevfusion_lidar-cam_...py:45-55
    in_channels=256, out_channels=80, image_size=(256, 704), feature_size=(32, 88), xbound=[-54.0, 54.0, 0.3],
    ybound=[-54.0, 54.0, 0.3], zbound=[-10.0, 10.0, 20.0], dbound=[1.0, 60.0, 0.5])


class BEVFrontEnd(nn.Module):

    def __init__(self, voxelize_cfg=None, encoder_cfg=None, view_cfg=None, precision="fp32"):
        super().__init__()
        self.pts_voxel_layer = Voxelization(**(voxelize_cfg or NUSCENES_VOXELIZE_CFG))
        self.pts_middle_encoder = BEVFusionSparseEncoder(**(encoder_cfg or NUSCENES_ENCODER_CFG))
        self.view_transform = BaseViewTransform(**(view_cfg or NUSCENES_VIEW_CFG))
        self.precision = precision
        for m in self.pts_middle_encoder.modules():
            if hasattr(m, "precision") and hasattr(m, "indice_key"):
                m.precision = precision

    @torch.no_grad()
    def set_calibration(self, rig):
        """rig: dict of [B, N, ...] tensors (camera2lidar_rots / _trans, intrins_inverse, post_rots_inverse,
        post_trans) -> builds the per-calibration pooling tables (deploy/voxel_detection.py:98-106 does this once
        per rig as well)."""
        geom = self.view_transform.get_geometry(**rig)
        return self.view_transform.build_tables(geom)

    @torch.no_grad()
    def voxelize(self, points):
        """bevfusion.py:227-255 with voxelize_reduce=True: list of [N_k, C] -> feats[M,C], coords[M,4], sizes[M]."""
        return self.pts_voxel_layer.forward_mean(points)

    def extract_pts_feat(self, points):
        feats, coords, _ = self.voxelize(points)
        return self.pts_middle_encoder(feats, coords, len(points))

    def extract_img_bev(self, depth, ctx, tables=None):
        return self.view_transform.pool_fused(depth, ctx, tables)

    def forward(self, points, depth, ctx, tables=None):
        """-> (lidar_bev [B, 256, 180, 180], camera_bev [B, 80, 360, 360])"""
        return self.extract_pts_feat(points), self.extract_img_bev(depth, ctx, tables)


class HostPipeline:
    """Host-buffer entry point of the front end (what `bench.py` times as `e2e`).

    `depth` sync-free plans (StaticFrontEnd, one CUDA graph each) are used round robin.  Pinned host inputs go to a
    plan's static input buffers on a copy-in stream, the plan's graph replays on the compute stream, and both BEV maps
    return to pinned host buffers on a copy-out stream: the copies of frame i overlap the compute of frame i+1, so
    steady-state throughput is max(compute, copy-in, copy-out), not their sum, and the host issues ~10 calls per frame.
    """

    def __init__(self, model, tables, device, depth=2, batch=1, max_points=400000, example=None, level_growth=None):
        from .static_frontend import StaticFrontEnd

        self.device = torch.device(device)
        # one compute stream per plan: consecutive frames overlap on the GPU (the head of frame i+1 -- voxelizer,
        # index build -- and its small kernels fill the SMs the tail of frame i leaves idle)
        self.computes = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.depth = depth
        self.plans = [StaticFrontEnd(model, tables, device, batch=batch, max_points=max_points, level_growth=level_growth)
                      for _ in range(depth)]
        if example is not None:  # (points list, depth, ctx): representative frame for warm-up + capture
            for p, c in zip(self.plans, self.computes):
                c.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(c):
                    p.load_inputs(*example)
                    p.capture()
                torch.cuda.current_stream(self.device).wait_stream(c)
        self.host = [None] * depth           # (lidar_host, cam_host) pinned
        self.e_comp = [None] * depth         # plan's graph finished (its inputs may be overwritten)
        self.e_done = [None] * depth         # plan's outputs copied out (its outputs may be overwritten)
        self.n = 0

    @torch.no_grad()
    def submit(self, points, depth, ctx, compact=False):
        """points: list of pinned [N_k, C] tensors (one per sample); depth / ctx pinned.  Returns the slot id.
        compact=True (opt-in): the BEV maps are cast to bf16 on the device and leave as bf16 -- half the device->host
        bytes (the dense fp32 maps are 79 % of the bytes a frame moves over the host link, which is what bounds the
        end-to-end rate when 8 GPUs share the host's links); the default keeps the reference's fp32 maps."""
        slot = self.n % self.depth
        self.n += 1
        plan = self.plans[slot]
        compute = self.computes[slot]
        with torch.cuda.stream(self.s_in):
            if self.e_comp[slot] is not None:
                self.s_in.wait_event(self.e_comp[slot])
            plan.load_inputs(points, depth, ctx)
            e_in = torch.cuda.Event()
            e_in.record(self.s_in)
        with torch.cuda.stream(compute):
            compute.wait_event(e_in)
            if self.e_done[slot] is not None:
                compute.wait_event(self.e_done[slot])
            lidar, cam = plan.replay() if plan.graph is not None else plan.run()
            if compact:
                if getattr(plan, "_compact", None) is None:
                    plan._compact = (torch.empty(lidar.shape, dtype=torch.bfloat16, device=self.device),
                                     torch.empty(cam.shape, dtype=torch.bfloat16, device=self.device))
                plan._compact[0].copy_(lidar)
                plan._compact[1].copy_(cam)
                lidar, cam = plan._compact
            self.e_comp[slot] = torch.cuda.Event()
            self.e_comp[slot].record(compute)
        if self.host[slot] is None or self.host[slot][0].dtype != lidar.dtype:
            self.host[slot] = (torch.empty(lidar.shape, dtype=lidar.dtype).pin_memory(),
                               torch.empty(cam.shape, dtype=cam.dtype).pin_memory())
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.e_comp[slot])
            self.host[slot][0].copy_(lidar, non_blocking=True)
            self.host[slot][1].copy_(cam, non_blocking=True)
            self.e_done[slot] = torch.cuda.Event()
            self.e_done[slot].record(self.s_out)
        return slot

    def result(self, slot):
        """Blocks until the slot's copies have landed -> (lidar_bev_host, camera_bev_host)."""
        self.e_done[slot].synchronize()
        return self.host[slot]

    def submit_device(self, points, depth, ctx, ways=None):
        """Same rotation for inputs that already live on the device (no host copies): -> (slot, lidar, cam); the
        returned maps are the plan's own buffers, valid until the slot comes round again.  `ways` (<= depth) limits the
        rotation to the first plans: without host copies to hide, two frames in flight keep the GPU busiest."""
        slot = self.n % (min(ways, self.depth) if ways else self.depth)
        self.n += 1
        plan, compute = self.plans[slot], self.computes[slot]
        caller = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(compute):
            compute.wait_stream(caller)          # the caller's writes to the inputs are ordered before the copy
            plan.load_inputs(points, depth, ctx)
            lidar, cam = plan.replay() if plan.graph is not None else plan.run()
            self.e_comp[slot] = torch.cuda.Event()
            self.e_comp[slot].record(compute)
        return slot, lidar, cam

    def join(self, stream=None):
        """Make `stream` (default: the current one) wait for everything submitted so far."""
        stream = stream or torch.cuda.current_stream(self.device)
        for c in self.computes:
            stream.wait_stream(c)
        stream.wait_stream(self.s_out)

    def drain(self):
        self.s_out.synchronize()
        for c in self.computes:
            c.synchronize()
