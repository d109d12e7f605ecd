"""bev_pool() and its autograd functions with the reference's surface
(projects/BEVFusion/bevfusion/ops/bev_pool/bev_pool.py:7-172), plus the fused north-star form
(`BevPoolTables`, `bev_pool_fused`) in which depth x context, the kept / sort gathers, the interval sum,
the permute and the collapse-Z of depth_lss.py:184-202, 723-725 are one kernel.
"""
import os

import torch

from ..._lib import check, cur_stream, lib, ptr
from . import bev_pool_ext


def _interval_table(ranks, n):
    """interval starts / lengths of a rank-sorted sequence (bev_pool.py:46-55, 154-159)."""
    first = torch.ones(n, device=ranks.device, dtype=torch.bool)
    first[1:] = ranks[1:] != ranks[:-1]
    starts = torch.where(first)[0].int()
    lengths = torch.empty_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = n - starts[-1]
    return starts, lengths


class QuickCumsum(torch.autograd.Function):
    """Pure-torch cumsum formulation (bev_pool.py:7-34); kept for API parity, not used by bev_pool()."""

    @staticmethod
    def forward(ctx, x, geom_feats, ranks):
        csum = x.cumsum(0)
        last = torch.ones(x.shape[0], device=x.device, dtype=torch.bool)
        last[:-1] = ranks[1:] != ranks[:-1]
        csum, geom_feats = csum[last], geom_feats[last]
        pooled = torch.cat((csum[:1], csum[1:] - csum[:-1]))
        ctx.save_for_backward(last)
        ctx.mark_non_differentiable(geom_feats)
        return pooled, geom_feats

    @staticmethod
    def backward(ctx, gradx, gradgeom):
        (last,) = ctx.saved_tensors
        owner = torch.cumsum(last, 0)
        owner[last] -= 1
        return gradx[owner], None, None


class QuickCumsumTrainingCuda(torch.autograd.Function):
    """bev_pool.py:43-90: forward builds the interval table from `ranks`; backward broadcasts out_grad."""

    @staticmethod
    def forward(ctx, x, geom_feats, ranks, B, D, H, W):
        starts, lengths = _interval_table(ranks, x.shape[0])
        geom_feats = geom_feats.int()
        out = bev_pool_ext.bev_pool_forward(x.contiguous(), geom_feats.contiguous(), lengths, starts, int(B), int(D),
                                            int(H), int(W))
        ctx.save_for_backward(starts, lengths, geom_feats)
        ctx.saved_shapes = int(B), int(D), int(H), int(W)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        starts, lengths, geom_feats = ctx.saved_tensors
        B, D, H, W = ctx.saved_shapes
        x_grad = bev_pool_ext.bev_pool_backward(out_grad.contiguous(), geom_feats.contiguous(), lengths, starts, B, D,
                                                H, W)
        return x_grad, None, None, None, None, None, None


class QuickCumsumCuda(torch.autograd.Function):
    """bev_pool.py:93-143: inference-only form taking the interval table; ONNX symbolic
    `autoware::QuickCumsumCuda` with the same attribute names."""

    @staticmethod
    def symbolic(g, x, geom_feats, interval_lengths, interval_starts, B, D, H, W):
        from torch.onnx.symbolic_helper import _get_tensor_dim_size, _get_tensor_sizes

        output = g.op("autoware::QuickCumsumCuda", x, geom_feats, interval_lengths, interval_starts,
                      batch_size_i=B, dimension_i=D, height_i=H, width_i=W, outputs=1)
        if _get_tensor_sizes(x) is not None and hasattr(x.type(), "with_sizes"):
            output.setType(x.type().with_sizes([B, D, H, W, _get_tensor_dim_size(x, -1)]))
        return output

    @staticmethod
    def forward(ctx, x, geom_feats, interval_lengths, interval_starts, B, D, H, W):
        return bev_pool_ext.bev_pool_forward(x.contiguous(), geom_feats.contiguous(), interval_lengths,
                                             interval_starts, B, D, H, W)

    @staticmethod
    def backward(ctx, out_grad):
        raise NotImplementedError


def _as_int(v):
    return int(v.item()) if isinstance(v, torch.Tensor) else int(v)


def bev_pool(feats, coords, ranks, B, D, H, W, is_training):
    """bev_pool.py:146-172.  feats[Nk,C] / coords[Nk,4] (x,y,z,b) / ranks[Nk] sorted by rank ->
    [B, C, D, H, W] contiguous.  D, H, W may be python ints, 0-dim tensors or nn.Parameters."""
    assert feats.shape[0] == coords.shape[0]
    if is_training:
        x = QuickCumsumTrainingCuda.apply(feats, coords, ranks, _as_int(B), _as_int(D), _as_int(H), _as_int(W))
    else:
        starts, lengths = _interval_table(ranks, feats.shape[0])
        if coords.dtype != torch.int32:
            coords = coords.int()
        x = QuickCumsumCuda.apply(feats, coords, lengths, starts, _as_int(B), _as_int(D), _as_int(H), _as_int(W))
    return x.permute(0, 4, 1, 2, 3).contiguous()


# ------------------------------------------------------------------------------------------------------
# Fused form
# ------------------------------------------------------------------------------------------------------
class BevPoolTables:
    """Per-calibration index tables of the fused kernel, derived from the reference's own
    `bev_pool_aux` outputs (depth_lss.py:118-176; the deploy path precomputes exactly these once per rig,
    deploy/voxel_detection.py:98-106):

      src            [nk]      int32  frustum index (into B*N*D*fH*fW) of each kept point, rank-sorted
      interval_starts[n_int+1] int32  CSR offsets into src
      interval_cell  [n_int]   int32  output cell (b*nz+z)*nx*ny + x*ny + y, ascending
      cell_of_point  [N']      int32  output cell of every frustum point or -1 (backward)

    With `frustum_shape=(B*N, D, fH, fW)` the ray-major run tables of the two-phase forward are built as well
    (run_p0, run_len, cell_run_starts, cell_run_ids; see csrc/bev_pool.cu) and used when runs average >= 4 points.
    """

    def __init__(self, geom_feats, kept, ranks, indices, B, nz, nx, ny, frustum_shape=None):
        dev = geom_feats.device
        nk = geom_feats.shape[0]
        kept_idx = torch.nonzero(kept.reshape(-1), as_tuple=False).squeeze(1)
        src = kept_idx[indices]
        g = geom_feats.long()
        cell = (g[:, 3] * nz + g[:, 2]) * (nx * ny) + g[:, 0] * ny + g[:, 1]
        # the fused kernel walks cells in memory order: re-sort by cell (stable), which for B == 1, nz == 1 is
        # already the rank order
        if nk > 1 and bool((cell[1:] < cell[:-1]).any()):
            cell, perm = torch.sort(cell, stable=True)
            src = src[perm]
        first = torch.ones(nk, device=dev, dtype=torch.bool)
        if nk > 0:
            first[1:] = cell[1:] != cell[:-1]
        starts = torch.nonzero(first, as_tuple=False).squeeze(1)
        self.n_intervals = int(starts.numel())
        self.nk = int(nk)
        self.src = src.int().contiguous()
        self.interval_cell = cell[starts].int().contiguous()
        self.interval_starts = torch.cat([starts, starts.new_tensor([nk])]).int().contiguous()
        nprime = kept.numel()
        cop = torch.full((nprime,), -1, dtype=torch.int32, device=dev)
        cop[src] = cell.int()
        self.cell_of_point = cop
        self.B, self.nz, self.nx, self.ny = int(B), int(nz), int(nx), int(ny)
        self.use_runs = False
        self.n_runs = 0
        if frustum_shape is not None and nk > 0:
            self._build_runs(src.long(), cell.long(), frustum_shape)

    _workspaces = {}

    @classmethod
    def from_geometry(cls, geom, B, bx, dx, nx):
        """Device-side build (csrc/bev_tables.cu): geom [B, N, D, fH, fW, 3] CUDA fp32 -> the same tables as
        `BevPoolTables(*bev_pool_aux(geom), ...)`, bit for bit, in ~10 kernels and one 12-byte read-back.
        bx / dx / nx: the view transform's grid (first-cell centre, cell size, (nx, ny, nz) counts)."""
        if not geom.is_cuda:
            raise RuntimeError("BevPoolTables.from_geometry: CUDA tensors only (use the bev_pool_aux form on CPU)")
        import ctypes

        from ..._lib import f32_array

        Bg, N, D, fH, fW, three = geom.shape
        assert three == 3 and Bg == B
        geom = geom.contiguous().float()
        dev = geom.device
        nxi, nyi, nzi = (int(v) for v in nx)
        nprime = Bg * N * D * fH * fW
        cells = B * nzi * nxi * nyi
        cap_int = min(nprime, cells)
        L = lib()
        L.bevf_bev_pool_tables_workspace_bytes.restype = ctypes.c_size_t
        ws_bytes = int(L.bevf_bev_pool_tables_workspace_bytes(ctypes.c_longlong(nprime), B, nzi, nxi, nyi))
        key = (dev, ws_bytes)
        ws = cls._workspaces.get(key)
        if ws is None:
            ws = cls._workspaces[key] = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        n_tiles = int(L.bevf_bev_pool_num_tiles(B, nzi, nxi, nyi))

        def i32(n):
            return torch.empty(n, dtype=torch.int32, device=dev)

        t = cls.__new__(cls)
        cop, src, run_p0, run_len, run_ids = i32(nprime), i32(nprime), i32(nprime), i32(nprime), i32(nprime)
        istarts, icell, crs = i32(cap_int + 1), i32(cap_int), i32(cap_int + 1)
        tile_starts, col_starts, counts = i32(n_tiles + 1), i32(Bg * N * fW + 1), i32(4)
        with torch.cuda.device(dev):
            check(L.bevf_bev_pool_build_tables(
                ptr(geom), Bg * N, D, fH, fW, B, f32_array([float(v) for v in bx]), f32_array([float(v) for v in dx]),
                nxi, nyi, nzi, ptr(cop), ptr(src), ptr(istarts), ptr(icell), ptr(tile_starts), ptr(run_p0),
                ptr(run_len), ptr(col_starts), ptr(run_ids), ptr(crs), ptr(counts), ptr(ws),
                ctypes.c_size_t(ws_bytes), cur_stream(dev)))
        nk, n_int, n_runs = (int(v) for v in counts[:3].tolist())   # the one host round trip
        t.nk, t.n_intervals, t.n_runs = nk, n_int, n_runs
        t.src = src[:nk]
        t.interval_cell = icell[:n_int]
        t.interval_starts = istarts[:n_int + 1]
        t.cell_of_point = cop
        t.B, t.nz, t.nx, t.ny = int(B), nzi, nxi, nyi
        t.run_p0, t.run_len, t.cell_run_ids = run_p0[:n_runs], run_len[:n_runs], run_ids[:n_runs]
        t.run_pos = _inverse_permutation(t.cell_run_ids)
        t.col_run_starts, t.cell_run_starts, t.tile_starts = col_starts, crs[:n_int + 1], tile_starts
        t.use_runs = nk > 0 and n_runs * 4 <= nk
        return t

    def _build_runs(self, p, cell, frustum_shape):
        _, D, fH, fW = [int(v) for v in frustum_shape]
        plane = fH * fW
        bn = p // (D * plane)
        d = (p // plane) % D
        h = (p // fW) % fH
        w = p % fW
        q = ((bn * fW + w) * D + d) * fH + h           # ray-major key: rows h of one (camera, column, depth bin) adjacent
        q, order = torch.sort(q)
        cs, ps = cell[order], p[order]
        first = torch.ones_like(q, dtype=torch.bool)
        first[1:] = (q[1:] != q[:-1] + 1) | ((q[1:] // fH) != (q[:-1] // fH)) | (cs[1:] != cs[:-1])
        starts = torch.nonzero(first, as_tuple=False).squeeze(1)
        n_runs = int(starts.numel())
        lens = torch.diff(torch.cat([starts, starts.new_tensor([q.numel()])]))
        run_cell = cs[starts]
        by_cell = torch.argsort(run_cell, stable=True)
        cells_u, counts = torch.unique_consecutive(run_cell[by_cell], return_counts=True)
        assert torch.equal(cells_u.int(), self.interval_cell), "run tables and interval tables disagree on the cell set"
        self.n_runs = n_runs
        self.run_p0 = ps[starts].int().contiguous()
        self.run_len = lens.int().contiguous()
        col = (self.run_p0.long() // (D * plane)) * fW + self.run_p0.long() % fW
        n_cols = int(frustum_shape[0]) * fW
        self.col_run_starts = torch.cat([col.new_zeros(1), torch.cumsum(torch.bincount(col, minlength=n_cols), 0)]
                                        ).int().contiguous()
        self.cell_run_ids = by_cell.int().contiguous()
        self.run_pos = _inverse_permutation(self.cell_run_ids)
        self.cell_run_starts = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)]).int().contiguous()
        self.use_runs = n_runs * 4 <= self.nk
        L = lib()
        n_tiles = int(L.bevf_bev_pool_num_tiles(self.B, self.nz, self.nx, self.ny))
        self.tile_starts = torch.empty(n_tiles + 1, dtype=torch.int32, device=p.device)
        if p.is_cuda:
            with torch.cuda.device(p.device):
                check(L.bevf_bev_pool_tile_starts(ptr(self.interval_cell), self.n_intervals, self.B, self.nz, self.nx,
                                                  self.ny, ptr(self.tile_starts), cur_stream(p.device)))
        else:  # CPU table build (tests): same table with searchsorted; tiles are 32 cells along y
            tiles_y = n_tiles // (self.B * self.nz * self.nx)
            t = torch.arange(n_tiles + 1)
            cell0 = (t // tiles_y) * self.ny + (t % tiles_y) * 32
            self.tile_starts = torch.searchsorted(self.interval_cell.long(), cell0).int()


def _inverse_permutation(perm):
    """run_pos[cell_run_ids[j]] = j: where a run's partial row goes so that every cell's rows are contiguous."""
    inv = torch.empty_like(perm)
    inv[perm.long()] = torch.arange(perm.numel(), dtype=perm.dtype, device=perm.device)
    return inv.contiguous()


def fused_v2_supported(c, fw):
    return fw % 4 == 0 and c % 16 == 0 and c <= 256


def nchw_to_nhwc(x):
    """[N, C, H, W] contiguous -> [N, H, W, C] contiguous (library transpose kernel)."""
    n, c, h, w = x.shape
    out = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().bevf_nchw_to_nhwc(ptr(x), ptr(out), int(n), int(c), int(h * w), cur_stream(x.device)))
    return out


def nhwc_to_nchw(x):
    n, h, w, c = x.shape
    out = torch.empty((n, c, h, w), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().bevf_nhwc_to_nchw(ptr(x), ptr(out), int(n), int(c), int(h * w), cur_stream(x.device)))
    return out


FUSED_V2 = os.environ.get("BEVFRONT_POOL_V2", "1") == "1"   # 0: the first-generation three-launch forward


class _BevPoolFused(torch.autograd.Function):

    @staticmethod
    def forward(ctx, depth, ctx_feats, tables):
        # depth [BN, D, fH, fW], ctx_feats [BN, C, fH, fW] (NCHW, as the depthnet emits them)
        depth = depth.contiguous()
        ctx_feats = ctx_feats.contiguous()
        bn, d, fh, fw = depth.shape
        c = ctx_feats.shape[1]
        t = tables
        out = torch.empty((t.B, c * t.nz, t.nx, t.ny), dtype=torch.float32, device=depth.device)
        v2 = t.use_runs and fused_v2_supported(c, fw) and FUSED_V2
        ctx_nhwc = None if v2 else nchw_to_nhwc(ctx_feats)
        with torch.cuda.device(depth.device):
            if v2:   # two launches straight from the NCHW tensors (csrc/bev_pool_v2.cu)
                partial = torch.empty((t.n_runs, c), dtype=torch.float32, device=depth.device)
                check(lib().bevf_bev_pool_fused_forward_v2(
                    ptr(depth), ptr(ctx_feats), ptr(t.run_p0), ptr(t.run_len), ptr(t.run_pos), t.n_runs,
                    ptr(t.col_run_starts), ptr(t.cell_run_starts), ptr(t.interval_cell), ptr(t.tile_starts),
                    t.n_intervals, int(bn), int(d), int(fh), int(fw), int(c), t.B, t.nz, t.nx, t.ny, ptr(partial),
                    ptr(out), cur_stream(depth.device)))
            elif t.use_runs:
                partial = torch.empty((t.n_runs, c), dtype=torch.float32, device=depth.device)
                check(lib().bevf_bev_pool_fused_forward_runs(
                    ptr(depth), ptr(ctx_nhwc), ptr(t.run_p0), ptr(t.run_len), t.n_runs, ptr(t.col_run_starts),
                    ptr(t.cell_run_starts),
                    ptr(t.cell_run_ids), ptr(t.interval_cell), ptr(t.tile_starts), t.n_intervals, int(bn), int(d), int(fh), int(fw),
                    int(c), t.B, t.nz, t.nx, t.ny, ptr(partial), ptr(out), cur_stream(depth.device)))
            else:
                check(lib().bevf_bev_pool_fused_forward(ptr(depth), ptr(ctx_nhwc), ptr(t.src), ptr(t.interval_starts),
                                                        ptr(t.interval_cell), t.n_intervals, t.nk, int(bn), int(d),
                                                        int(fh), int(fw), int(c), t.B, t.nz, t.nx, t.ny, ptr(out),
                                                        cur_stream(depth.device)))
        ctx.save_for_backward(depth, ctx_feats if ctx_nhwc is None else ctx_nhwc)
        ctx.ctx_is_nchw = ctx_nhwc is None
        ctx.tables = tables
        return out

    @staticmethod
    def backward(ctx, out_grad):
        depth, ctx_nhwc = ctx.saved_tensors
        if ctx.ctx_is_nchw:   # the v2 forward read NCHW directly; the backward kernel wants channels-last
            ctx_nhwc = nchw_to_nhwc(ctx_nhwc)
        t = ctx.tables
        bn, d, fh, fw = depth.shape
        c = ctx_nhwc.shape[3]
        dev = depth.device
        out_grad = out_grad.contiguous()
        g_nhwc = torch.empty((t.B * t.nz * t.nx * t.ny, c), dtype=torch.float32, device=dev)
        d_depth = torch.empty_like(depth)
        d_ctx_nhwc = torch.empty_like(ctx_nhwc)
        with torch.cuda.device(dev):
            L = lib()
            # out[b, z*C + ch, x, y] (z-major channels, depth_lss.py:202) viewed as [B*nz, C, nx*ny] -> cell-major rows
            check(L.bevf_nchw_to_nhwc(ptr(out_grad), ptr(g_nhwc), t.B * t.nz, int(c), t.nx * t.ny, cur_stream(dev)))
            check(L.bevf_bev_pool_fused_backward(ptr(g_nhwc), ptr(depth), ptr(ctx_nhwc), ptr(t.cell_of_point), int(bn),
                                                 int(d), int(fh), int(fw), int(c), ptr(d_depth), ptr(d_ctx_nhwc),
                                                 cur_stream(dev)))
        return d_depth, nhwc_to_nchw(d_ctx_nhwc), None


def bev_pool_fused(depth, ctx_feats, tables):
    """depth [B*N, D, fH, fW] (softmax), ctx_feats [B*N, C, fH, fW] -> [B, C*nz, nx, ny]; differentiable in
    both inputs.  Equals depth_lss.py get_cam_feats' outer product followed by BaseViewTransform.bev_pool."""
    return _BevPoolFused.apply(depth, ctx_feats, tables)
