"""SparseConvTensor with the spconv-2.x surface the reference uses
(`spconv.pytorch.SparseConvTensor`: sparse_encoder.py:131-147 of projects/BEVFusion, sparse_block.py:17-24 of
mmdet3d/models/layers, projects/SparseConvolution/sparse_conv.py:70-103,155-162): features / indices /
spatial_shape / batch_size / indice_dict / find_indice_pair / replace_feature / dense / shadow_copy.
"""
import ctypes
import os

import torch

from .._lib import check, cur_stream, i32_array, lib, ptr


_CHECK_INDEX = os.environ.get("BEVFRONT_CHECK_INDEX", "0") == "1"


class CoordIndex:
    """Coordinate -> row lookup structure of one sparse level (occupancy bitmap + popcount prefix in one
    device buffer, plus the rank -> row permutation when the rows are not in ascending linear order)."""

    def __init__(self, indices, batch_size, spatial_shape, sorted_rows=False, mem=None):
        self.batch_size = int(batch_size)
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.shape_c = i32_array(self.spatial_shape)
        dev = indices.device
        L = lib()
        self.nbytes = int(L.bevf_spconv_index_bytes(self.batch_size, self.shape_c))
        if self.nbytes == 0:
            raise RuntimeError("sparse grid not supported: " + L.bevf_last_error().decode())
        if mem is not None:  # built by bevf_spconv_strided_sites
            self.mem = mem
            self.perm = None
            return
        self.mem = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        n = indices.shape[0]
        self.perm = None if sorted_rows else torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(L.bevf_spconv_index_build(ptr(indices), int(n), None, self.batch_size, self.shape_c, ptr(self.mem),
                                            ctypes.c_size_t(self.nbytes), ptr(self.perm), cur_stream(dev)))

    def error_code(self):
        """0 ok, 1 a coordinate lies outside the grid, 2 duplicate coordinates (one host sync)."""
        addr = lib().bevf_spconv_index_error_flag(ptr(self.mem), ctypes.c_size_t(self.nbytes), self.batch_size,
                                                  self.shape_c)
        off = addr - self.mem.data_ptr()
        return int(self.mem[off:off + 4].view(torch.int32).item())


class IndicePair:
    """The rulebook of one convolution geometry (what spconv stores in `indice_dict[indice_key]`)."""

    def __init__(self, out_indices, pair_fwd, n_out, out_spatial_shape, in_index, out_index, ksize, stride, padding,
                 dilation, subm):
        self.out_indices = out_indices
        self.pair_fwd = pair_fwd            # [kv, n_out] int32, -1 = no input
        self.indice_pairs = pair_fwd        # spconv attribute name
        self.n_out = n_out
        self.out_spatial_shape = out_spatial_shape
        self.in_index = in_index
        self.out_index = out_index
        self.ksize, self.stride, self.padding, self.dilation, self.subm = ksize, stride, padding, dilation, subm

    def matches(self, ksize, stride, padding, dilation, subm):
        return (self.ksize, self.stride, self.padding, self.dilation, self.subm) == (ksize, stride, padding,
                                                                                      dilation, subm)


class _ToDense(torch.autograd.Function):
    """dense() / dense_bev() through bevf_sparse_to_dense; backward gathers the dense gradient at the active sites
    (bevf_dense_to_sparse), so the encoder trains through its BEV tail."""

    @staticmethod
    def forward(ctx, features, indices, batch_size, spatial_shape, bev_layout):
        f = features.contiguous().float()
        n, c = f.shape
        X, Y, Z = spatial_shape
        shape = (batch_size, c * Z, X, Y) if bev_layout else (batch_size, c, X, Y, Z)
        out = torch.empty(shape, dtype=torch.float32, device=f.device)
        with torch.cuda.device(f.device):
            check(lib().bevf_sparse_to_dense(ptr(f), ptr(indices), int(n), None, int(c), batch_size,
                                             i32_array(spatial_shape), ptr(out), int(bev_layout),
                                             cur_stream(f.device)))
        ctx.save_for_backward(indices)
        ctx.meta = (n, c, batch_size, spatial_shape, bev_layout)
        return out

    @staticmethod
    def backward(ctx, grad):
        (indices,) = ctx.saved_tensors
        n, c, batch_size, spatial_shape, bev_layout = ctx.meta
        g = grad.contiguous().float()
        out = torch.empty((n, c), dtype=torch.float32, device=g.device)
        if n:
            with torch.cuda.device(g.device):
                check(lib().bevf_dense_to_sparse(ptr(g), ptr(indices), int(n), None, int(c), batch_size,
                                                 i32_array(spatial_shape), int(bev_layout), ptr(out),
                                                 cur_stream(g.device)))
        return out, None, None, None, None


class SparseConvTensor:

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None, indice_dict=None,
                 benchmark=False, permanent_thrust_allocator=False, enable_timer=False, force_algo=None):
        assert features.dim() == 2, "features must be [N, C]"
        assert indices.dim() == 2 and indices.shape[1] == 4, "indices must be [N, 4] (batch, x, y, z)"
        assert indices.dtype == torch.int32, "indices must be int32"
        assert features.shape[0] == indices.shape[0]
        self._features = features
        self.indices = indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = {} if indice_dict is None else indice_dict
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self.benchmark_record = {}
        self.thrust_allocator = None
        self.force_algo = force_algo
        self.int8_scale = None
        self._index = None          # CoordIndex of this level (shared through shadow copies)
        self._sorted_rows = False   # rows known to be in ascending linear order (output of a strided conv)
        self._bf16 = None           # bf16 copy of the features written by a tensor-core epilogue

    # ---- spconv surface ---------------------------------------------------------------------------------
    @property
    def features(self):
        # a tensor-core layer that was told not to write fp32 (need_f32 = False) leaves only the bf16 copy; the
        # fp32 view is materialised on first access
        if self._features is None and self._bf16 is not None:
            self._features = self._bf16.float()
        return self._features

    @features.setter
    def features(self, val):  # spconv 2.x forbids in-place assignment; mmcv-style code paths use it
        self._features = val
        self._bf16 = None

    def replace_feature(self, feature):
        """New tensor sharing indices / rulebooks with `self` (spconv 2.x idiom, sparse_block.py:17-24)."""
        new = self.shadow_copy()
        new._features = feature
        new._bf16 = None
        return new

    def shadow_copy(self):
        new = SparseConvTensor.__new__(SparseConvTensor)
        new.__dict__.update(self.__dict__)
        return new

    @property
    def spatial_size(self):
        size = 1
        for s in self.spatial_shape:
            size *= s
        return size

    @property
    def sparity(self):
        return self.indices.shape[0] / self.spatial_size / self.batch_size

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def dense(self, channels_first=True):
        """[B, C, X, Y, Z] (channels_first) or [B, X, Y, Z, C]."""
        out = self._to_dense(bev_layout=False)
        return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()

    def dense_bev(self):
        """Extension: dense().permute(0,1,4,2,3).view(B, C*Z, X, Y) in one kernel (the tail of
        BEVFusionSparseEncoder.forward, sparse_encoder.py:147-151)."""
        return self._to_dense(bev_layout=True)

    # ---- internals --------------------------------------------------------------------------------------
    def _to_dense(self, bev_layout):
        return _ToDense.apply(self.features, self.indices, self.batch_size, tuple(self.spatial_shape), bool(bev_layout))

    def coord_index(self):
        if self._index is None:
            self._index = CoordIndex(self.indices, self.batch_size, self.spatial_shape, sorted_rows=self._sorted_rows)
            if _CHECK_INDEX:   # debug mode (one host sync per tensor): out-of-grid / duplicate coordinates raise here
                code = self._index.error_code()
                if code:
                    raise ValueError("SparseConvTensor indices: " + ("a coordinate lies outside the grid" if code == 1
                                                                    else "duplicate coordinates"))
        return self._index

    def __repr__(self):
        src = self._features if self._features is not None else self._bf16
        return f"SparseConvTensor[shape={tuple(src.shape)}]"
