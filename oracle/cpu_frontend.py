"""The reference's CPU formulation of the BEV front end, one frame at a time -- TEST / BASELINE INFRASTRUCTURE ONLY
(imported by bench.py's `cpu_baseline` leg and `--impl reference` arm, never by the product package).

  voxelize      the reference's own C++ hard_voxelize_cpu (oracle/_ref/ref_voxel_fixed, compiled from
                projects/BEVFusion/bevfusion/ops/voxel/src/*.cpp; serial code) when that build is present, else the
                C restatement in bevfront_oracle.c; then the mean of bevfusion.py:251-253 in torch.
  sparse conv   the reference has no CPU sparse conv of its own (spconv is CUDA-only); its documented fallback
                mmcv.ops runs indice_conv on CPU as gather -> at::mm -> scatter-add per kernel tap, which is what
                `cpu_sparse_conv` does with torch (all host threads), on the rulebook from the C oracle (OpenMP).
                Eval-mode BatchNorm1d / ReLU / residual follow sparse_block.py:137-154 in torch.
  bev_pool      the data path of depth_lss.py:723-725, 184-202 in torch: outer product -> reshape -> x[kept] ->
                x[indices] -> index_add_ into [B*D*H*W, C] (the formulation BASELINE.json names) -> permute ->
                collapse Z.
"""
import numpy as np
import torch

import oracle
from oracle import build_ref


_REF_VOXEL = None


def ref_voxel_module():
    global _REF_VOXEL
    if _REF_VOXEL is None:
        try:
            build_ref.build()
        except Exception:
            pass
        try:
            _REF_VOXEL = build_ref.load_ref("ref_voxel_fixed") or False
        except Exception:
            _REF_VOXEL = False
    return _REF_VOXEL or None


def cpu_voxelize_mean(points, voxel_size, coors_range, max_points, max_voxels):
    """-> (feats[M,C] f32, coords[M,4] i32 (b=0,x,y,z), kind)"""
    mod = ref_voxel_module()
    pts = torch.from_numpy(np.ascontiguousarray(points))
    if mod is not None:
        voxels = pts.new_zeros((max_voxels, max_points, pts.shape[1]))
        coors = pts.new_zeros((max_voxels, 3), dtype=torch.int32)
        npv = pts.new_zeros((max_voxels,), dtype=torch.int32)
        m = mod.hard_voxelize(pts, voxels, coors, npv, [float(v) for v in voxel_size],
                              [float(v) for v in coors_range], int(max_points), int(max_voxels), 3, True)
        voxels, coors, npv = voxels[:m], coors[:m], npv[:m]
        kind = "reference"
    else:
        v, c, n = oracle.hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels)
        voxels, coors, npv = torch.from_numpy(v), torch.from_numpy(c), torch.from_numpy(n)
        kind = "port"
    feats = voxels.sum(dim=1) / npv.type_as(voxels).view(-1, 1)
    coords = torch.nn.functional.pad(coors, (1, 0), mode="constant", value=0)
    return feats.contiguous(), coords.contiguous(), kind


def cpu_sparse_conv(feats, idx, shape, weight, ksize, stride, padding, dilation, subm):
    """feats [n,Cin] torch f32, idx [n,4] numpy i32, weight [Cout,kD,kH,kW,Cin] torch -> (out, out_idx, out_shape)"""
    out_idx, pair, out_shape = oracle.spconv_rulebook(idx, shape, ksize, stride, padding, dilation, subm)
    cout, cin = weight.shape[0], weight.shape[-1]
    w = weight.reshape(cout, -1, cin)
    out = torch.zeros((out_idx.shape[0], cout), dtype=torch.float32)
    pair_t = torch.from_numpy(pair)
    for k in range(pair.shape[0]):
        sel = torch.nonzero(pair_t[k] >= 0, as_tuple=False).squeeze(1)
        if sel.numel() == 0:
            continue
        out.index_add_(0, sel, feats[pair_t[k][sel].long()] @ w[:, k, :].t())
    return out, out_idx, [int(v) for v in out_shape]


def _bn(bn, x, train=False):
    if train:   # batch statistics (train-mode BatchNorm1d); running statistics are not updated by the checker
        return torch.nn.functional.batch_norm(x, None, None, bn["weight"], bn["bias"], True, 0.0, bn["eps"])
    return torch.nn.functional.batch_norm(x, bn["mean"], bn["var"], bn["weight"], bn["bias"], False, 0.0, bn["eps"])


def cpu_sparse_encoder(layers, feats, idx, shape, batch, train=False):
    """layers: the plan produced by `encoder_plan` (weights on CPU) -> dense BEV [B, C*Z, X, Y].  train=True: batch-
    statistics BatchNorm; every op is a differentiable torch op, so autograd gives the reference formulation's backward
    (gather -> mm -> scatter-add per tap and their transposes: what mmcv's CPU indice_conv_backward does)."""
    f, i, shp = feats, idx, list(shape)
    for L in layers:
        if L["type"] == "convmodule":
            f, i, shp = cpu_sparse_conv(f, i, shp, L["weight"], L["ksize"], L["stride"], L["padding"], L["dilation"],
                                        L["subm"])
            f = torch.relu(_bn(L["bn"], f, train))
        else:  # basic block
            h, _, _ = cpu_sparse_conv(f, i, shp, L["w1"], (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1), True)
            h = torch.relu(_bn(L["bn1"], h, train))
            h, _, _ = cpu_sparse_conv(h, i, shp, L["w2"], (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1), True)
            f = torch.relu(_bn(L["bn2"], h, train) + f)
    c = f.shape[1]
    X, Y, Z = shp
    it = torch.from_numpy(i).long()
    dense = torch.zeros((batch, X, Y, Z, c), dtype=torch.float32).index_put((it[:, 0], it[:, 1], it[:, 2], it[:, 3]), f)
    return dense.permute(0, 4, 3, 1, 2).contiguous().view(batch, c * Z, X, Y)


def encoder_plan(enc):
    """Flatten a BEVFusionSparseEncoder-shaped module tree (ours or the reference's: same attribute names) into
    the list `cpu_sparse_encoder` walks.  Weights are copied to CPU fp32."""
    def bn(m):
        return dict(weight=m.weight.detach().float().cpu(), bias=m.bias.detach().float().cpu(),
                    mean=m.running_mean.detach().float().cpu(), var=m.running_var.detach().float().cpu(), eps=m.eps)

    def convmodule(seq):
        c = seq[0]
        return dict(type="convmodule", weight=c.weight.detach().float().cpu(), ksize=tuple(c.kernel_size),
                    stride=tuple(c.stride), padding=tuple(c.padding), dilation=tuple(c.dilation), subm=bool(c.subm),
                    bn=bn(seq[1]))

    plan = [convmodule(enc.conv_input)]
    for stage in enc.encoder_layers:
        for blk in stage:
            if hasattr(blk, "conv1"):
                plan.append(dict(type="block", w1=blk.conv1.weight.detach().float().cpu(), bn1=bn(blk.bn1),
                                 w2=blk.conv2.weight.detach().float().cpu(), bn2=bn(blk.bn2)))
            else:
                plan.append(convmodule(blk))
    plan.append(convmodule(enc.conv_out))
    return plan


def cpu_bev_pool(depth, ctx, kept, indices, geom_feats, B, N, nz, nx, ny):
    """depth [B*N, D, fH, fW], ctx [B*N, C, fH, fW] torch f32; kept / indices / geom_feats from bev_pool_aux
    (torch, CPU) -> [B, C*nz, nx, ny]"""
    BN, D, fH, fW = depth.shape
    C = ctx.shape[1]
    x = depth.unsqueeze(1) * ctx.unsqueeze(2)                    # depth_lss.py:723
    x = x.view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)     # :724-725
    x = x.reshape(B * N * D * fH * fW, C)                        # :184
    x = x[kept]                                                  # :189
    x = x[indices]                                               # :192
    g = geom_feats.long()
    cell = ((g[:, 3] * nz + g[:, 2]) * nx + g[:, 0]) * ny + g[:, 1]
    out = torch.zeros((B * nz * nx * ny, C), dtype=torch.float32)
    out.index_add_(0, cell, x)                                   # K1 (bev_pool_cuda.cu:20-42) as index_add_
    out = out.view(B, nz, nx, ny, C).permute(0, 4, 1, 2, 3).contiguous()   # bev_pool.py:170
    return torch.cat(out.unbind(dim=2), 1)                       # depth_lss.py:202
