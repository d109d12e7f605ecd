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


class RowsResult:
    """Host side of HostPipeline's "rows" output: pinned buffers the pack kernels wrote + what is needed to read them.

      lidar()  -> (indices [n, 4] int32 (b, x, y, z), rows [n, C] fp32): the sparse encoder's last level
      camera() -> (cells [n_cells] int32 (cell = (b*nz + z)*nx*ny + x*ny + y), columns [C, n_cells] fp32)
      dense()  -> the two dense fp32 maps, bit-identical to the "dense" output (host-side scatter)"""

    def __init__(self, lidar_raw, cam, cap, c_lidar, cells, n_cells, lidar_shape, cam_shape, nz):
        self.lidar_raw, self.cam, self.cap, self.c_lidar = lidar_raw, cam, cap, c_lidar
        self.cells, self.n_cells, self.lidar_shape, self.cam_shape, self.nz = cells, n_cells, lidar_shape, cam_shape, nz
        self.cam_dev = None    # device staging of the camera columns (copy-engine path)

    def views(self):
        """(indices [cap, 4], rows [cap, C]) views of the pinned buffer: header {n, ...} | indices | rows"""
        idx = self.lidar_raw[16:16 + self.cap * 16].view(torch.int32).view(self.cap, 4)
        rows = self.lidar_raw[16 + self.cap * 16:].view(torch.float32).view(self.cap, self.c_lidar)
        return idx, rows

    def lidar(self):
        n = min(int(self.lidar_raw[:4].view(torch.int32)[0]), self.cap)
        idx, rows = self.views()
        return idx[:n], rows[:n]

    def camera(self):
        return self.cells, self.cam[:, :self.n_cells]

    def nbytes(self):
        """bytes the pack kernels wrote for this frame"""
        n = min(int(self.lidar_raw[:4].view(torch.int32)[0]), self.cap)
        return 16 + n * 16 + n * self.c_lidar * 4 + self.cam.shape[0] * ((self.n_cells + 3) // 4 * 4) * 4

    def dense(self):
        idx, rows = self.lidar()
        B, CZ, X, Y = self.lidar_shape
        Z = CZ // self.c_lidar
        lidar = torch.zeros((B, self.c_lidar, Z, X, Y), dtype=torch.float32)      # bev[b, ch*Z + z, x, y]
        i = idx.long()
        lidar[i[:, 0], :, i[:, 3], i[:, 1], i[:, 2]] = rows
        cells, cols = self.camera()
        Bc, CZc, nx, ny = self.cam_shape
        C = CZc // self.nz
        cam = torch.zeros((Bc * self.nz, C, nx * ny), dtype=torch.float32)        # [b*nz + z, ch, xy]
        cl = cells.long()
        cam[cl // (nx * ny), :, cl % (nx * ny)] = cols.t()
        return lidar.view(B, CZ, X, Y), cam.view(Bc, self.nz * C, nx, ny)


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
        self._pending = []                   # slots whose "rows" copies wait for their row count

    @torch.no_grad()
    def submit(self, points, depth, ctx, compact=False, output=None):
        """points: list of pinned [N_k, C] tensors (one per sample); depth / ctx pinned.  Returns the slot id.
        output (the dense fp32 maps are 79 % of the bytes a frame moves over the host link, which is what bounds the
        end-to-end rate, above all when 8 GPUs share the host's links):
          "dense" (default)  the reference's fp32 dense maps [B,256,180,180], [B,80,360,360]: 74.6 MB per frame
          "rows"             LOSSLESS: the LiDAR map as the sparse encoder's active rows + coordinates, the camera map as
                             the columns of the cells the frustum reaches (fp32; `RowsResult.dense()` rebuilds both maps
                             bit for bit on the host): ~30 MB per frame, leaving through the
                             copy engine with exact sizes once the frame's row count has arrived ("rows_zero_copy":
                             the pack kernels of csrc/pack_out.cu store straight into the pinned buffers instead)
          "bf16" (= compact=True)  both dense maps cast to bf16: 37 MB, lossy"""
        output = output or ("bf16" if compact else "dense")
        assert output in ("dense", "rows", "rows_zero_copy", "bf16")
        compact = output == "bf16"
        if output.startswith("rows"):
            return self._submit_rows(points, depth, ctx, zero_copy=output == "rows_zero_copy")
        slot = self.n % self.depth
        self.n += 1
        self._before_reuse(slot)
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
        if not isinstance(self.host[slot], tuple) or self.host[slot][0].dtype != lidar.dtype:
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
        """Blocks until the slot's copies have landed -> (lidar_bev_host, camera_bev_host), or a RowsResult."""
        self._before_reuse(slot)
        self.e_done[slot].synchronize()
        return self.host[slot]

    # ---- lossless "rows" output ---------------------------------------------------------------------------------
    def _rows_buffers(self, slot, plan):
        from ._lib import lib

        bufs = getattr(self, "_rows", None)
        if bufs is None:
            bufs = self._rows = [None] * self.depth
        if bufs[slot] is None:
            lv, rows, t = plan.last_level, plan.last_rows, plan.tables
            c_l, c_c = int(rows.shape[1]), int(plan.cam_bev.shape[1]) // t.nz
            nbytes = int(lib().bevf_pack_sparse_rows_bytes(lv.cap, c_l))
            pitch = (t.n_intervals + 3) // 4 * 4
            bufs[slot] = RowsResult(torch.empty(nbytes, dtype=torch.uint8).pin_memory(),
                                    torch.empty((c_c, pitch), dtype=torch.float32).pin_memory(), lv.cap, c_l,
                                    t.interval_cell.cpu(), t.n_intervals, tuple(plan.lidar_bev.shape),
                                    tuple(plan.cam_bev.shape), t.nz)
        return bufs[slot]

    @torch.no_grad()
    def _submit_rows(self, points, depth, ctx, zero_copy=False, max_blocks=0):
        """zero_copy=False (default): the camera columns are gathered into a device buffer and both maps leave through the
        copy engine with EXACT sizes; the LiDAR row count (known on the device only) follows the frame as a 4-byte copy,
        and a frame's copies are enqueued once it has arrived -- i.e. depth - 1 submissions later, so the GPU always has
        the next frames queued and the host never waits for the frame it just submitted.
        zero_copy=True: the pack kernels store straight into the pinned host buffers (no row count on the host at all;
        50 GB/s alone on the GPU, but their CTAs compete with the frames in flight for SM slots: measured slower)."""
        import ctypes

        from ._lib import check, cur_stream, lib, ptr

        slot = self.n % self.depth
        self.n += 1
        self._before_reuse(slot)
        plan, compute = self.plans[slot], self.computes[slot]
        with torch.cuda.stream(self.s_in):
            if self.e_comp[slot] is not None:
                self.s_in.wait_event(self.e_comp[slot])
            plan.load_inputs(points, depth, ctx)
            e_in = torch.cuda.Event()
            e_in.record(self.s_in)
        L, t = lib(), plan.tables
        with torch.cuda.stream(compute), torch.cuda.device(self.device):
            compute.wait_event(e_in)
            if self.e_done[slot] is not None:
                compute.wait_event(self.e_done[slot])
            plan.replay() if plan.graph is not None else plan.run()
            res = self._rows_buffers(slot, plan)
            if not zero_copy:
                if res.cam_dev is None:
                    res.cam_dev = torch.empty(res.cam.shape, dtype=torch.float32, device=self.device)
                check(L.bevf_pack_cells(ptr(plan.cam_bev), ptr(t.interval_cell), t.n_intervals, int(res.cam.shape[0]), t.nz,
                                        t.nx * t.ny, ptr(res.cam_dev), int(res.cam.shape[1]), 0, cur_stream(self.device)))
                res.lidar_raw[:4].copy_(plan.last_level.n_dev.view(torch.uint8), non_blocking=True)   # the row count
            self.e_comp[slot] = torch.cuda.Event()
            self.e_comp[slot].record(compute)
        self.host[slot] = res
        if zero_copy:
            with torch.cuda.stream(self.s_out), torch.cuda.device(self.device):
                self.s_out.wait_event(self.e_comp[slot])
                st = cur_stream(self.device)
                check(L.bevf_pack_sparse_rows(ptr(plan.last_rows), ptr(plan.last_level.indices), plan.last_level.cap,
                                              ptr(plan.last_level.n_dev), res.c_lidar,
                                              ctypes.c_void_p(res.lidar_raw.data_ptr()),
                                              ctypes.c_size_t(res.lidar_raw.numel()), int(max_blocks), st))
                check(L.bevf_pack_cells(ptr(plan.cam_bev), ptr(t.interval_cell), t.n_intervals, int(res.cam.shape[0]),
                                        t.nz, t.nx * t.ny, ctypes.c_void_p(res.cam.data_ptr()), int(res.cam.shape[1]),
                                        int(max_blocks), st))
                self.e_done[slot] = torch.cuda.Event()
                self.e_done[slot].record(self.s_out)
            return slot
        self.e_done[slot] = None
        self._pending.append(slot)
        while len(self._pending) > self.depth - 1:
            self._flush_slot(self._pending[0])
        return slot

    def _flush_slot(self, slot):
        """Enqueue the exact-size device->host copies of a "rows" frame whose row count has arrived."""
        assert self._pending and self._pending[0] == slot
        self._pending.pop(0)
        plan, res = self.plans[slot], self.host[slot]
        self.e_comp[slot].synchronize()                       # the 4-byte count copy is ordered before this event
        n = min(int(res.lidar_raw[:4].view(torch.int32)[0]), res.cap)
        idx_h, rows_h = res.views()
        with torch.cuda.stream(self.s_out):
            idx_h[:n].copy_(plan.last_level.indices[:n], non_blocking=True)
            rows_h[:n].copy_(plan.last_rows[:n], non_blocking=True)
            res.cam.copy_(res.cam_dev, non_blocking=True)
            self.e_done[slot] = torch.cuda.Event()
            self.e_done[slot].record(self.s_out)

    def _before_reuse(self, slot):
        """A slot that still has un-flushed "rows" copies must enqueue them before its plan runs again."""
        while slot in self._pending:
            self._flush_slot(self._pending[0])

    def submit_device(self, points, depth, ctx, ways=None):
        """Same rotation for inputs that already live on the device (no host copies): -> (slot, lidar, cam); the
        returned maps are the plan's own buffers, valid until the slot comes round again.  `ways` (<= depth) limits the
        rotation to the first plans: without host copies to hide, two frames in flight keep the GPU busiest."""
        slot = self.n % (min(ways, self.depth) if ways else self.depth)
        self.n += 1
        self._before_reuse(slot)
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
        while self._pending:
            self._flush_slot(self._pending[0])
        for c in self.computes:
            stream.wait_stream(c)
        stream.wait_stream(self.s_out)

    def drain(self):
        while self._pending:
            self._flush_slot(self._pending[0])
        self.s_out.synchronize()
        for c in self.computes:
            c.synchronize()
