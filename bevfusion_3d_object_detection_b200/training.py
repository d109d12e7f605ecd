"""Training split of the front end (BASELINE configs[2]): forward + backward of the sparse encoder and of the fused
bev_pool over a batch of frames per GPU, frame-parallel across GPUs, with the parameter-gradient all-reduce overlapped
with the backward pass (parallel.GradBucketReducer; the reference: mmengine's DDP wrapper,
configs/_base_/default_runtime.py:14, tools/dist_train.sh:10-19).

What is on the path, where the reference's detector has it:
  BEVFusion.voxelize            bevfusion.py:227-255   hard voxelize (max_voxels[0], the training cap) + mean, no grad
  BEVFusionSparseEncoder        sparse_encoder.py:112-156   train mode: batch-statistics BatchNorm1d, no fused epilogues;
                                                            conv forward / data gradient / weight gradient on tcgen05
  DepthLSSTransform.get_cam_feats depth_lss.py:699-725   depth = depth + (max(depth_gt, depth) - depth).detach(), outer
                                                            product + bev_pool fused, gradients to depth and context
The heads / losses are outside the hot path: the step back-propagates a fixed random linear functional of both BEV maps
(every output element gets a gradient, as a dense head would give it).
"""
import torch

from . import parallel


def calibrated_depth(depth, depth_gt=None):
    """depth_lss.py:702-706: in training the softmax depth is lifted to at least the LiDAR ground-truth distribution,
    with the correction detached (gradients flow through the prediction only)."""
    if depth_gt is None:
        return depth
    return depth + (torch.maximum(depth_gt, depth) - depth).detach()


class TrainStep:
    """One data-parallel training step of the BEV front end on this rank's frames.

        step = TrainStep(model, tables, lr=1e-4)
        loss = step(points_list, depth, ctx, depth_gt=None)      # tensors already on the device

    points_list: B frames [N_k, C]; depth [B*N_cam, D, fH, fW] (softmax), ctx [B*N_cam, C, fH, fW]: both get gradients
    (they come from the depthnet).  Returns the scalar loss tensor (on the device; the caller decides when to read it)."""

    def __init__(self, model, tables, lr=1e-4, bucket_bytes=2 << 20, seed=0, process_group=None):
        self.model, self.tables, self.lr = model, tables, float(lr)
        self.enc = model.pts_middle_encoder
        self.enc.train()
        model.pts_voxel_layer.train()          # selects max_voxels[0]
        self.params = [p for p in self.enc.parameters() if p.requires_grad]
        self.reducer = parallel.GradBucketReducer(self.params, bucket_bytes=bucket_bytes, process_group=process_group)
        self.seed = seed
        self._proj = {}

    def _projection(self, t, key):
        p = self._proj.get(key)
        if p is None or p.shape != t.shape:
            g = torch.Generator(device=t.device).manual_seed(self.seed + len(self._proj))
            p = torch.randn(t.shape, device=t.device, generator=g) * (1.0 / t[0].numel())
            self._proj[key] = p
        return p

    def __call__(self, points_list, depth, ctx, depth_gt=None):
        model, enc = self.model, self.enc
        self.reducer.prepare()
        with torch.no_grad():
            feats, coords, _ = model.voxelize(points_list)
        lidar = enc(feats, coords, len(points_list))
        d = depth.detach().requires_grad_(True)     # fresh leaves over the caller's storage: no gradient build-up
        c = ctx.detach().requires_grad_(True)
        cam = model.view_transform.pool_fused(calibrated_depth(d, depth_gt), c, self.tables)
        loss = (lidar * self._projection(lidar, "lidar")).sum() + (cam * self._projection(cam, "cam")).sum()
        loss.backward()                        # bucket all-reduces start as the gradients of a bucket complete
        self.reducer.finish()
        self.cam_grads = (d.grad, c.grad)
        with torch.no_grad():                  # SGD on the (averaged) bucket views: one fused op per bucket
            for b in self.reducer.buckets:
                torch._foreach_add_([p for p in b["params"]], [p.grad for p in b["params"]], alpha=-self.lr)
        return loss.detach()

    def grads(self):
        return {n: p.grad for n, p in self.enc.named_parameters()}
