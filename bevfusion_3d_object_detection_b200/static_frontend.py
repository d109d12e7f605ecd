"""Sync-free execution plan of the BEV front end (eval mode), optionally replayed as ONE CUDA graph per frame.

`BEVFrontEnd.forward` keeps the reference's call structure, and with it the reference's host round trips: the
voxel count comes back to the host to slice the padded tensors (voxelize.py:57-60, bevfusion.py:230-246) and every
strided sparse conv returns its number of active outputs to the host (spconv's num_act_out).  Here every buffer is
sized for its worst case once, every row count stays in device memory (the `n_dev` arguments of the C ABI), and the
whole frame -- voxelize + mean -> cell sort -> 5 coordinate indices / rulebooks -> 21 gather-GEMMs -> dense BEV, and
the camera branch's two-phase bev_pool -- is a fixed launch sequence that CUDA-graph capture turns into one replay.
Same kernels, same order of arithmetic as the module path: the outputs are bit-identical (tests/test_static_gpu.py).
"""
import ctypes

import torch
from torch import nn

from ._lib import check, cur_stream, f32_array, i32_array, lib, ptr
from .ops.bev_pool.bev_pool import BevPoolTables
from .sparse_encoder import SparseBasicBlock
from .spconv import functional as Fsp
from .spconv.conv import SparseConvolution
from .spconv.modules import _fold_bn, bn_is_foldable

SENTINEL = 1.0e30  # coordinate of padding points: outside every point-cloud range, so the voxelizer drops them


def _pad4(n):
    return max(4, (int(n) + 3) // 4 * 4)


class _Level:
    """One sparse resolution level: capacity-sized coordinate buffers + device row count."""

    def __init__(self, cap, shape, batch, dev):
        L = lib()
        self.cap, self.shape, self.ld = int(cap), [int(s) for s in shape], _pad4(cap)
        self.shape_c = i32_array(self.shape)
        self.indices = torch.zeros((self.cap, 4), dtype=torch.int32, device=dev)
        self.n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.index_bytes = int(L.bevf_spconv_index_bytes(batch, self.shape_c))
        if self.index_bytes == 0:
            raise RuntimeError("sparse grid not supported: " + L.bevf_last_error().decode())
        self.index_mem = torch.empty(self.index_bytes, dtype=torch.uint8, device=dev)
        self.subm_pairs = {}   # (ksize, dilation) -> pair_fwd [kv, ld]
        self.subm_fresh = set()
        self.hint = 0          # expected active rows (tile-shape hint for the GEMM launches; 0 = unknown)


class StaticFrontEnd:

    def __init__(self, model, tables, device, batch=1, max_points=400000, level_growth=None):
        """level_growth: None sizes every strided level for its worst case (each input site reaching prod(ceil(k/s))
        outputs: 8x per level); a number caps the growth per level at that factor (real LiDAR frames grow < 2x) -- the
        kernels clip at the capacity, so `check_capacity()` must be called on a representative frame."""
        assert not model.training, "the static plan folds BatchNorm: eval mode only"
        self.level_growth = level_growth
        self.model, self.tables, self.dev = model, tables, torch.device(device)
        assert isinstance(tables, BevPoolTables) and tables.use_runs, "camera branch needs the run tables"
        self.batch, self.max_points = int(batch), int(max_points)
        self.precision = model.precision
        vox = model.pts_voxel_layer
        enc = model.pts_middle_encoder
        self.voxel_size, self.pc_range = list(vox.voxel_size), list(vox.point_cloud_range)
        self.max_num_points, self.max_voxels = int(vox.max_num_points), int(vox.max_voxels[1])
        self.c_in = int(enc.in_channels)
        dev, L = self.dev, lib()
        # ---- inputs (static addresses: what a captured graph reads) ------------------------------------------
        self.points = [torch.full((self.max_points, self.c_in), SENTINEL, dtype=torch.float32, device=dev)
                       for _ in range(self.batch)]
        self._n_valid = [0] * self.batch
        # ---- voxelizer ----------------------------------------------------------------------------------------
        cap0 = self.batch * min(self.max_voxels, self.max_points)
        self.vox_feats = torch.empty((cap0, self.c_in), dtype=torch.float32, device=dev)
        self.vox_coords = torch.empty((cap0, 4), dtype=torch.int32, device=dev)
        self.vox_sizes = torch.empty((cap0,), dtype=torch.int32, device=dev)
        self.vox_num = torch.zeros(1, dtype=torch.int32, device=dev)
        self.vox_ws_bytes = int(L.bevf_hard_voxelize_workspace_bytes(self.max_points, self.max_num_points,
                                                                     self.max_voxels))
        self.vox_ws = torch.empty(self.vox_ws_bytes, dtype=torch.uint8, device=dev)
        # ---- encoder plan -------------------------------------------------------------------------------------
        self.ops = self._flatten_encoder(enc)
        self.levels = [_Level(cap0, enc.sparse_shape, self.batch, dev)]
        self.perm0 = torch.empty(cap0, dtype=torch.int32, device=dev)
        self._plan_levels()
        self._plan_chain()
        self._alloc_features()
        # ---- camera branch ------------------------------------------------------------------------------------
        self.depth = None
        self.ctx = None
        self.graph = None
        self.cam_bev = None
        # side streams: rulebooks and the camera branch run beside the GEMM chain (also inside the captured graph)
        self.overlap = True
        self.s_rule = torch.cuda.Stream(dev)
        self.s_cam = torch.cuda.Stream(dev)

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _flatten_encoder(enc):
        """[(conv, bn_scale, bn_shift, relu, residual)] with residual in {None, 'block_in'} in execution order."""
        ops = []

        def convmodule(seq):
            mods = list(seq)
            conv, bn = mods[0], mods[1]
            assert isinstance(conv, SparseConvolution) and bn_is_foldable(bn) and isinstance(mods[2], nn.ReLU)
            s, b = _fold_bn(bn)
            ops.append(dict(conv=conv, scale=s, shift=b, relu=True, residual=None, block_start=False))

        convmodule(enc.conv_input)
        for stage in enc.encoder_layers:
            for blk in stage:
                if isinstance(blk, SparseBasicBlock):
                    assert blk.downsample is None
                    s1, b1 = _fold_bn(blk.norm1)
                    s2, b2 = _fold_bn(blk.norm2)
                    ops.append(dict(conv=blk.conv1, scale=s1, shift=b1, relu=True, residual=None, block_start=True))
                    ops.append(dict(conv=blk.conv2, scale=s2, shift=b2, relu=True, residual="block_in",
                                    block_start=False))
                else:
                    convmodule(blk)
        convmodule(enc.conv_out)
        return ops

    def _plan_levels(self):
        """Walk the ops, create a level for every strided conv, size everything for the worst case."""
        lvl = 0
        for op in self.ops:
            conv = op["conv"]
            op["level_in"] = lvl
            if not conv.subm:
                cur = self.levels[lvl]
                out_shape = Fsp.conv_out_shape(cur.shape, conv.kernel_size, conv.stride, conv.padding, conv.dilation)
                reach = 1
                for k, s, d in zip(conv.kernel_size, conv.stride, conv.dilation):
                    reach *= min(k, -(-k // s)) if d == 1 else k   # gcd(dilation, stride) > 1: up to k outputs per axis
                cells = self.batch * out_shape[0] * out_shape[1] * out_shape[2]
                if self.level_growth is not None:
                    reach = min(float(reach), float(self.level_growth))
                self.levels.append(_Level(min(int(cur.cap * reach), cells), out_shape, self.batch, self.dev))
                lvl += 1
                op["pair"] = torch.empty((conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2],
                                          self.levels[lvl].ld), dtype=torch.int32, device=self.dev)
            else:
                key = (tuple(conv.kernel_size), tuple(conv.dilation))
                level = self.levels[lvl]
                if key not in level.subm_pairs:
                    kv = key[0][0] * key[0][1] * key[0][2]
                    level.subm_pairs[key] = torch.empty((kv, level.ld), dtype=torch.int32, device=self.dev)
                op["pair"] = level.subm_pairs[key]
                op["subm_key"] = key
            op["level_out"] = lvl

    def _plan_chain(self):
        """Arguments of bevf_spconv_strided_sites_chain: the output sites of ALL strided levels in three launches from the
        level-0 coordinates (instead of mark + scan + emit per level, each waiting for the level below).  Used when every
        strided conv has dilation 1 (the encoder's do); BEVFRONT_SITES_CHAIN=0 keeps the level-by-level calls."""
        import os

        strided = [op for op in self.ops if not op["conv"].subm]
        ok = (os.environ.get("BEVFRONT_SITES_CHAIN", "1") == "1" and 1 <= len(strided) <= 6
              and all(tuple(op["conv"].dilation) == (1, 1, 1) for op in strided)
              and [op["level_out"] for op in strided] == list(range(1, len(strided) + 1)))
        self.chain = None
        if not ok:
            return
        n = len(strided)
        flat = lambda key: (ctypes.c_int * (3 * n))(*[int(v) for op in strided for v in getattr(op["conv"], key)])
        lv = [self.levels[op["level_out"]] for op in strided]
        self.chain = dict(
            n=n, ks=flat("kernel_size"), st=flat("stride"), pd=flat("padding"),
            mems=(ctypes.c_void_p * n)(*[l.index_mem.data_ptr() for l in lv]),
            bytes=(ctypes.c_size_t * n)(*[l.index_bytes for l in lv]),
            outs=(ctypes.c_void_p * n)(*[l.indices.data_ptr() for l in lv]),
            caps=(ctypes.c_int * n)(*[l.cap for l in lv]),
            ndevs=(ctypes.c_void_p * n)(*[l.n_dev.data_ptr() for l in lv]))

    def _alloc_features(self):
        """Three rotating feature slots per level (block input / hidden / output), fp32 and (bf16 path) bf16."""
        dev, L = self.dev, lib()
        bf16 = self.precision == "bf16"
        chans = {}
        for op in self.ops:
            chans[op["level_out"]] = max(chans.get(op["level_out"], 0), op["conv"].out_channels)
        self.slots = {}
        last_lvl = self.ops[-1]["level_out"]
        for lvl, c in chans.items():
            cap = self.levels[lvl].ld   # rows = the leading dimension the GEMM's TMA maps are encoded with (pad4(cap))
            need32 = (not bf16) or lvl == last_lvl or any(
                op["conv"].need_f32 for op in self.ops if op["level_out"] == lvl)
            self.slots[lvl] = [dict(f32=torch.empty((cap, c), dtype=torch.float32, device=dev) if need32 else None,
                                    bf16=torch.empty((cap, c), dtype=torch.bfloat16, device=dev) if bf16 else None)
                               for _ in range(3)]
        cap0 = self.levels[0].cap
        if bf16:
            self.cin_pad = int(L.bevf_spconv_tc_cin_pad(self.c_in))
            self.in_bf16 = torch.empty((cap0, self.cin_pad), dtype=torch.bfloat16, device=dev)
            self.in_f32 = None
        else:
            self.cin_pad = 0
            self.in_bf16 = None
            self.in_f32 = torch.empty((cap0, self.c_in), dtype=torch.float32, device=dev)
        last = self.ops[-1]["level_out"]
        X, Y, Z = self.levels[last].shape
        self.lidar_bev = torch.empty((self.batch, chans[last] * Z, X, Y), dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------------------------------------------------
    def load_inputs(self, points, depth, ctx):
        """Copy one frame into the static input buffers (async on the current stream).  points: list of [N_k, C]
        tensors (host pinned or device); the tail of each buffer keeps the out-of-range sentinel."""
        assert len(points) == self.batch
        for k, p in enumerate(points):
            n = int(p.shape[0])
            assert n <= self.max_points and p.shape[1] == self.c_in
            self.points[k][:n].copy_(p, non_blocking=True)
            if n < self._n_valid[k]:
                self.points[k][n:self._n_valid[k]].fill_(SENTINEL)
            self._n_valid[k] = n
        if self.depth is None:
            self.depth = torch.empty(depth.shape, dtype=torch.float32, device=self.dev)
            self.ctx = torch.empty(ctx.shape, dtype=torch.float32, device=self.dev)
            bn, c, fh, fw = ctx.shape
            self.ctx_nhwc = torch.empty((bn, fh, fw, c), dtype=torch.float32, device=self.dev)
            t = self.tables
            self.pool_partial = torch.empty((t.n_runs, c), dtype=torch.float32, device=self.dev)
            self.cam_bev = torch.empty((t.B, c * t.nz, t.nx, t.ny), dtype=torch.float32, device=self.dev)
        self.depth.copy_(depth, non_blocking=True)
        self.ctx.copy_(ctx, non_blocking=True)

    def _enqueue(self, gemm_events=None):
        """The whole frame as a fixed sequence of launches on the current stream (capturable).  `gemm_events`: list
        that receives a (start, end) CUDA event pair around every gather-GEMM launch (profiling, eager only)."""
        L, dev = lib(), self.dev
        st = cur_stream(dev)
        vs, rg = f32_array(self.voxel_size), f32_array(self.pc_range)
        lv0 = self.levels[0]
        # 0. the camera branch shares nothing with the LiDAR branch: its own stream, joined at the end
        if self.overlap:
            fork0 = torch.cuda.Event()
            fork0.record(torch.cuda.current_stream(dev))
            self.s_cam.wait_event(fork0)
            with torch.cuda.stream(self.s_cam):
                self._enqueue_camera(cur_stream(dev))
                cam_done = torch.cuda.Event()
                cam_done.record(self.s_cam)
        # 1. voxelize + mean + batch pad, appended sample after sample at the device row offset
        lv0.n_dev.zero_()
        for k in range(self.batch):
            check(L.bevf_voxelize_mean(ptr(self.points[k]), self.max_points, self.c_in, ptr(self.vox_feats),
                                       ptr(self.vox_coords), ptr(self.vox_sizes), vs, rg, self.max_num_points,
                                       self.max_voxels, k, ptr(self.vox_ws), ctypes.c_size_t(self.vox_ws_bytes),
                                       ptr(self.vox_num), ptr(lv0.n_dev), st))
        # 2. level-0 coordinate index + rows into ascending cell order (features straight into the conv operand)
        check(L.bevf_spconv_index_build(ptr(self.vox_coords), lv0.cap, ptr(lv0.n_dev), self.batch, lv0.shape_c,
                                        ptr(lv0.index_mem), ctypes.c_size_t(lv0.index_bytes), ptr(self.perm0), st))
        check(L.bevf_spconv_permute_rows(ptr(self.vox_feats), ptr(self.vox_coords), ptr(self.perm0), lv0.cap,
                                         ptr(lv0.n_dev), self.c_in, self.cin_pad, ptr(self.in_f32), ptr(self.in_bf16),
                                         ptr(lv0.indices), st))
        # 3. every rulebook on a side stream: they depend on coordinates only, so the index / site / rulebook kernels of
        #    the later levels overlap the gather-GEMMs of the earlier ones (which leave most SM threads idle).  The
        #    camera branch is independent of the LiDAR branch altogether and gets its own stream.
        main = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(main)
        if self.overlap:
            self.s_rule.wait_event(fork)
        rule_stream = self.s_rule if self.overlap else main
        with torch.cuda.stream(rule_stream):
            st_r = cur_stream(dev)
            chain_done = self.chain is not None
            if chain_done:
                # the sites of EVERY strided level follow from the level-0 coordinates alone: three launches instead of
                # mark + scan + emit (+ three memory operations) per level.  The sorted rows are the input (not the
                # voxelizer's first-appearance order): neighbouring threads then mark neighbouring cells (23 vs 41 us)
                ch = self.chain
                check(L.bevf_spconv_strided_sites_chain(ptr(lv0.indices), lv0.cap, ptr(lv0.n_dev), self.batch, lv0.shape_c,
                                                        ch["n"], ch["ks"], ch["st"], ch["pd"], ch["mems"], ch["bytes"],
                                                        ch["outs"], ch["caps"], ch["ndevs"], st_r))
            for lv in self.levels:
                lv.subm_fresh.clear()
            for op in self.ops:
                conv = op["conv"]
                lin, lout = self.levels[op["level_in"]], self.levels[op["level_out"]]
                ks, st_, pd, dl = (i32_array(conv.kernel_size), i32_array(conv.stride), i32_array(conv.padding),
                                   i32_array(conv.dilation))
                pair = op["pair"]
                op["ready"] = None
                if conv.subm:
                    if op["subm_key"] not in lin.subm_fresh:
                        check(L.bevf_spconv_subm_rulebook(ptr(lin.indices), lin.cap, ptr(lin.n_dev), self.batch,
                                                          lin.shape_c, ks, dl, ptr(lin.index_mem),
                                                          ctypes.c_size_t(lin.index_bytes), None, ptr(pair), lin.ld,
                                                          st_r))
                        lin.subm_fresh.add(op["subm_key"])
                        if self.overlap:
                            op["ready"] = torch.cuda.Event()
                            op["ready"].record(rule_stream)
                else:
                    if not chain_done:
                        check(L.bevf_spconv_strided_sites(ptr(lin.indices), lin.cap, ptr(lin.n_dev), self.batch,
                                                          lin.shape_c, ks, st_, pd, dl, ptr(lout.index_mem),
                                                          ctypes.c_size_t(lout.index_bytes), ptr(lout.indices), lout.cap,
                                                          ptr(lout.n_dev), st_r))
                    check(L.bevf_spconv_strided_rulebook(ptr(lout.indices), lout.cap, ptr(lout.n_dev), self.batch,
                                                         lin.shape_c, ks, st_, pd, dl, ptr(lin.index_mem),
                                                         ctypes.c_size_t(lin.index_bytes), None, ptr(pair), lout.ld,
                                                         st_r))
                    if self.overlap:
                        op["ready"] = torch.cuda.Event()
                        op["ready"].record(rule_stream)
        # 4. the 21 gather-GEMMs on the main stream, each after its rulebook
        cur = dict(f32=self.in_f32, bf16=self.in_bf16, f32_valid=self.in_f32 is not None)
        block_in = None
        for op in self.ops:
            conv = op["conv"]
            lin, lout = self.levels[op["level_in"]], self.levels[op["level_out"]]
            kv = conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2]
            pair = op["pair"]
            if op["ready"] is not None:
                main.wait_event(op["ready"])
            if op["block_start"]:
                block_in = cur
            # output slot: any of the level's three that is neither the input nor the pending block input
            slots = self.slots[op["level_out"]]
            out = next(s for s in slots if s is not cur and s is not block_in)
            residual = None
            if op["residual"] == "block_in" and block_in.get("f32_valid", False):
                residual = block_in["f32"]
            w = conv._packed_weight(self.precision)
            if gemm_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if self.precision == "bf16":
                cin_pad = int(L.bevf_spconv_tc_cin_pad(conv.in_channels))
                # same choices as the module path: fp32 copy only where a layer asks for it, skip connection from the
                # fp32 copy if the block input has one, else from its bf16 operand copy
                residual_bf16 = None
                if op["residual"] == "block_in" and residual is None:
                    residual_bf16 = block_in["bf16"]
                out_f32 = out["f32"] if conv.need_f32 else None
                out["f32_valid"] = out_f32 is not None
                check(L.bevf_spconv_gemm_bf16(ptr(cur["bf16"]), lin.cap, ptr(w), ptr(pair), lout.ld, lout.hint,
                                              ptr(lout.n_dev), kv, cin_pad, conv.out_channels, ptr(conv.bias),
                                              ptr(op["scale"]), ptr(op["shift"]), ptr(residual), ptr(residual_bf16),
                                              int(op["relu"]), ptr(out_f32), ptr(out["bf16"]), st))
            else:
                out["f32_valid"] = True
                check(L.bevf_spconv_gemm_f32(ptr(cur["f32"]), ptr(w), ptr(pair), lout.ld, lout.hint, ptr(lout.n_dev), kv,
                                             conv.in_channels, conv.out_channels, ptr(conv.bias), ptr(op["scale"]),
                                             ptr(op["shift"]), ptr(residual), int(op["relu"]), ptr(out["f32"]), st))
            if gemm_events is not None:
                ev[1].record()
                gemm_events.append(ev)
            if op["residual"] == "block_in":
                block_in = None
            cur = out
        # 5. dense()+permute+view of the last level
        last = self.levels[self.ops[-1]["level_out"]]
        c_last = self.ops[-1]["conv"].out_channels
        import os as _os
        if c_last % 4 == 0 and _os.environ.get("BEVFRONT_BEV_INDEXED", "1") == "1":
            # output-driven: every (channel, z) line of the map written once, zeros included (no memset, no scattered stores)
            check(L.bevf_sparse_to_bev_indexed(ptr(cur["f32"]), last.cap, ptr(last.n_dev), c_last, self.batch, last.shape_c,
                                               ptr(last.index_mem), ctypes.c_size_t(last.index_bytes), ptr(self.lidar_bev),
                                               st))
        else:
            check(L.bevf_sparse_to_dense(ptr(cur["f32"]), ptr(last.indices), last.cap, ptr(last.n_dev), c_last, self.batch,
                                         last.shape_c, ptr(self.lidar_bev), 1, st))
        # the same information as active rows (HostPipeline's lossless "rows" output reads these)
        self.last_rows, self.last_level = cur["f32"], last
        # 6. camera branch (already in flight on its own stream when overlapping)
        if self.overlap:
            main.wait_event(cam_done)
        else:
            self._enqueue_camera(st)

    def _enqueue_camera(self, st):
        """context to channels-last, ray-major partial sums, cell-major gather (csrc/bev_pool.cu)."""
        L = lib()
        t = self.tables
        bn, c, fh, fw = self.ctx.shape
        d = self.depth.shape[1]
        import importlib

        _bp = importlib.import_module(__package__ + ".ops.bev_pool.bev_pool")   # the module (the package re-exports a function of that name)

        if _bp.FUSED_V2 and _bp.fused_v2_supported(int(c), int(fw)):
            check(L.bevf_bev_pool_fused_forward_v2(ptr(self.depth), ptr(self.ctx), ptr(t.run_p0), ptr(t.run_len),
                                                   ptr(t.run_pos), t.n_runs, ptr(t.col_run_starts), ptr(t.cell_run_starts),
                                                   ptr(t.interval_cell), ptr(t.tile_starts), t.n_intervals, int(bn), int(d),
                                                   int(fh), int(fw), int(c), t.B, t.nz, t.nx, t.ny, ptr(self.pool_partial),
                                                   ptr(self.cam_bev), st))
            return
        check(L.bevf_nchw_to_nhwc(ptr(self.ctx), ptr(self.ctx_nhwc), int(bn), int(c), int(fh * fw), st))
        check(L.bevf_bev_pool_fused_forward_runs(ptr(self.depth), ptr(self.ctx_nhwc), ptr(t.run_p0), ptr(t.run_len),
                                                 t.n_runs, ptr(t.col_run_starts), ptr(t.cell_run_starts),
                                                 ptr(t.cell_run_ids), ptr(t.interval_cell), ptr(t.tile_starts),
                                                 t.n_intervals, int(bn), int(d), int(fh), int(fw), int(c), t.B, t.nz,
                                                 t.nx, t.ny, ptr(self.pool_partial), ptr(self.cam_bev), st))

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self):
        """Enqueue the frame eagerly (inputs must have been loaded).  -> (lidar_bev, cam_bev), plan-owned tensors."""
        with torch.cuda.device(self.dev):
            self._enqueue()
        return self.lidar_bev, self.cam_bev

    @torch.no_grad()
    def capture(self, calibrate=True):
        """Capture the frame into a CUDA graph.  One eager warm-up first (kernel attributes get set outside the
        capture); with `calibrate` its per-level site counts (one host sync, once) become the tile-shape hints of
        the captured GEMM launches.  Load a representative frame before calling."""
        with torch.cuda.device(self.dev):
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self._enqueue()
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            if calibrate:
                for lv, n in zip(self.levels, self.counts()):
                    lv.hint = min(int(n), lv.cap)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self.graph = g
        return self

    def replay(self):
        self.graph.replay()
        return self.lidar_bev, self.cam_bev

    @torch.no_grad()
    def profile(self, reps=5):
        """Device time of every gather-GEMM launch (best of `reps` eager runs) and its useful flops
        (2 * valid rulebook pairs * Cin * Cout).  -> list of dict(cin, cout, subm, ms, flops, rows)."""
        best = None
        with torch.cuda.device(self.dev):
            for _ in range(reps):
                evs = []
                self._enqueue(gemm_events=evs)
                torch.cuda.synchronize(self.dev)
                ms = [a.elapsed_time(b) for a, b in evs]
                best = ms if best is None else [min(x, y) for x, y in zip(best, ms)]
        counts = self.counts()
        out = []
        for op, ms in zip(self.ops, best):
            conv = op["conv"]
            n = counts[op["level_out"]]
            pairs = int((op["pair"][:, :n] >= 0).sum().item())
            out.append(dict(cin=conv.in_channels, cout=conv.out_channels, subm=bool(conv.subm), ms=ms, rows=n,
                            flops=2.0 * pairs * conv.in_channels * conv.out_channels))
        return out

    def check_capacity(self):
        """Raise if the last frame overflowed a level's capacity (only possible with level_growth set): the site
        kernels count every site and clip the rows they write, so n_dev > cap means truncated output."""
        for i, (lv, n) in enumerate(zip(self.levels, self.counts())):
            if n > lv.cap:
                raise RuntimeError(f"static plan: level {i} has {n} active sites but capacity {lv.cap}; raise "
                                   "level_growth (or pass None for worst-case sizing)")

    def counts(self):
        """Active sites per level (one host sync; diagnostics only)."""
        return [int(lv.n_dev.item()) for lv in self.levels]

    def error_codes(self):
        """Coordinate-index error flag of every level (0 ok / 1 out of grid / 2 duplicate); one host sync each."""
        out = []
        for lv in self.levels:
            addr = lib().bevf_spconv_index_error_flag(ptr(lv.index_mem), ctypes.c_size_t(lv.index_bytes), self.batch,
                                                      lv.shape_c)
            off = addr - lv.index_mem.data_ptr()
            out.append(int(lv.index_mem[off:off + 4].view(torch.int32).item()))
        return out
