"""`voxel_layer` -- same entry points as the reference pybind module
projects/BEVFusion/bevfusion/ops/voxel/src/voxelization.cpp:6-11 (hard_voxelize, dynamic_voxelize,
dynamic_point_to_voxel_forward, dynamic_point_to_voxel_backward), implemented over the C ABI of
libbevfront_b200.so.  CUDA tensors only: the reference's CPU branch is not shipped (calls raise).
"""
import ctypes

import torch

from ..._lib import check, cur_stream, f32_array, i32_array, lib, ptr

_REDUCE = {"sum": 0, "mean": 1, "max": 2}  # voxelization.h:4, :97-106


def _check_input(t, name, dtype=None):
    # CHECK_INPUT of voxelization_cuda.cu:8-14
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (the B200 build ships no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")


def _workspace(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points, max_voxels,
                  NDim=3, deterministic=True):
    """voxelization.h:58-81.  Fills the caller-allocated (zero-filled) voxels / coors /
    num_points_per_voxel and returns voxel_num as a host int (one D2H sync, part of the contract)."""
    _check_input(points, "points", torch.float32)
    _check_input(voxels, "voxels", torch.float32)
    _check_input(coors, "coors", torch.int32)
    _check_input(num_points_per_voxel, "num_points_per_voxel", torch.int32)
    n, c = points.shape
    with torch.cuda.device(points.device):
        L = lib()
        nbytes = L.bevf_hard_voxelize_workspace_bytes(int(n), int(max_points), int(max_voxels))
        ws = _workspace(nbytes, points.device)
        vnum = torch.empty(1, dtype=torch.int32, device=points.device)
        host = ctypes.c_int(0)
        check(L.bevf_hard_voxelize_sync(ptr(points), int(n), int(c), ptr(voxels), ptr(coors),
                                        ptr(num_points_per_voxel), f32_array(voxel_size), f32_array(coors_range),
                                        int(max_points), int(max_voxels), int(NDim), int(bool(deterministic)), 0,
                                        ptr(ws), ctypes.c_size_t(nbytes), ptr(vnum), cur_stream(points.device),
                                        ctypes.byref(host)))
    return int(host.value)


def hard_voxelize_async(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points,
                        max_voxels, zero_fill=True):
    """Extension: no host sync; returns voxel_num as a 1-element device int32 tensor.  With zero_fill the
    output buffers may be torch.empty()."""
    _check_input(points, "points", torch.float32)
    _check_input(voxels, "voxels", torch.float32)
    _check_input(coors, "coors", torch.int32)
    _check_input(num_points_per_voxel, "num_points_per_voxel", torch.int32)
    n, c = points.shape
    with torch.cuda.device(points.device):
        L = lib()
        nbytes = L.bevf_hard_voxelize_workspace_bytes(int(n), int(max_points), int(max_voxels))
        ws = _workspace(nbytes, points.device)
        vnum = torch.empty(1, dtype=torch.int32, device=points.device)
        check(L.bevf_hard_voxelize(ptr(points), int(n), int(c), ptr(voxels), ptr(coors), ptr(num_points_per_voxel),
                                   f32_array(voxel_size), f32_array(coors_range), int(max_points), int(max_voxels),
                                   3, 1, int(bool(zero_fill)), ptr(ws), ctypes.c_size_t(nbytes), ptr(vnum),
                                   cur_stream(points.device)))
    return vnum


def voxelize_mean(points, feats, coords4, sizes, voxel_size, coors_range, max_points, max_voxels, batch_idx=0,
                  row_offset=None):
    """Extension: hard voxelize + mean reduce + batch pad in one pass (bevfusion.py:227-255).
    Returns voxel_num as a 1-element device int32 tensor; rows are appended at *row_offset when given."""
    _check_input(points, "points", torch.float32)
    _check_input(feats, "feats", torch.float32)
    _check_input(coords4, "coords4", torch.int32)
    _check_input(sizes, "sizes", torch.int32)
    n, c = points.shape
    with torch.cuda.device(points.device):
        L = lib()
        nbytes = L.bevf_hard_voxelize_workspace_bytes(int(n), int(max_points), int(max_voxels))
        ws = _workspace(nbytes, points.device)
        vnum = torch.empty(1, dtype=torch.int32, device=points.device)
        check(L.bevf_voxelize_mean(ptr(points), int(n), int(c), ptr(feats), ptr(coords4), ptr(sizes),
                                   f32_array(voxel_size), f32_array(coors_range), int(max_points), int(max_voxels),
                                   int(batch_idx), ptr(ws), ctypes.c_size_t(nbytes), ptr(vnum), ptr(row_offset),
                                   cur_stream(points.device)))
    return vnum


def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
    """voxelization.h:83-95.  coors[N,3] int32 written in place (xyz; partial -1 rows as the GPU reference)."""
    _check_input(points, "points", torch.float32)
    _check_input(coors, "coors", torch.int32)
    n, c = points.shape
    with torch.cuda.device(points.device):
        check(lib().bevf_dynamic_voxelize(ptr(points), int(n), int(c), ptr(coors), f32_array(voxel_size),
                                          f32_array(coors_range), int(NDim), cur_stream(points.device)))


def _reduce_code(reduce_type):
    if reduce_type not in _REDUCE:
        raise RuntimeError("do not support reduce type " + str(reduce_type))  # voxelization.h:104
    return _REDUCE[reduce_type]


def dynamic_point_to_voxel_forward(feats, coors, reduce_type):
    """voxelization.h:107-121 -> [reduced_feats[M,C], out_coors[M,ndim], coors_map[N] i32, reduce_count[M] i32]"""
    code = _reduce_code(reduce_type)
    _check_input(feats, "feats", torch.float32)
    _check_input(coors, "coors", torch.int32)
    n, c = feats.shape
    ndim = coors.shape[1]
    dev = feats.device
    if n == 0:  # scatter_points_cuda.cu:192-196
        return [feats.clone().detach(), coors.clone().detach(), coors.new_empty((0,), dtype=torch.int32),
                coors.new_empty((0,), dtype=torch.int32)]
    with torch.cuda.device(dev):
        L = lib()
        ext_dev = torch.empty(4, dtype=torch.int32, device=dev)
        check(L.bevf_dynamic_scatter_extents(ptr(coors), int(n), int(ndim), ptr(ext_dev), cur_stream(dev)))
        ext = i32_array(ext_dev.tolist())
        nbytes = L.bevf_dynamic_scatter_workspace_bytes(int(n), int(ndim), ext)
        if nbytes == 0:
            raise RuntimeError("dynamic_point_to_voxel_forward: " + L.bevf_last_error().decode())
        ws = _workspace(nbytes, dev)
        reduced = torch.empty((n, c), dtype=torch.float32, device=dev)
        out_coors = torch.empty((n, ndim), dtype=torch.int32, device=dev)
        coors_map = torch.empty((n,), dtype=torch.int32, device=dev)
        reduce_count = torch.empty((n,), dtype=torch.int32, device=dev)
        m_dev = torch.empty(1, dtype=torch.int32, device=dev)
        check(L.bevf_dynamic_scatter_forward(ptr(feats), ptr(coors), int(n), int(c), int(ndim), ext, code,
                                             ptr(reduced), ptr(out_coors), ptr(coors_map), ptr(reduce_count),
                                             ptr(m_dev), ptr(ws), ctypes.c_size_t(nbytes), cur_stream(dev)))
        m = int(m_dev.item())
    return [reduced[:m], out_coors[:m], coors_map, reduce_count[:m]]


def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_idx, reduce_count,
                                    reduce_type):
    """voxelization.h:123-138.  grad_feats[N,C] is overwritten."""
    code = _reduce_code(reduce_type)
    for t, nme in ((grad_feats, "grad_feats"), (grad_reduced_feats, "grad_reduced_feats"), (feats, "feats"),
                   (reduced_feats, "reduced_feats")):
        _check_input(t, nme, torch.float32)
    _check_input(coors_idx, "coors_idx", torch.int32)
    _check_input(reduce_count, "reduce_count", torch.int32)
    n, c = feats.shape
    m = reduced_feats.shape[0]
    dev = feats.device
    with torch.cuda.device(dev):
        nbytes = max(256, m * c * 4 + 256)
        ws = _workspace(nbytes, dev)
        check(lib().bevf_dynamic_scatter_backward(ptr(grad_feats), ptr(grad_reduced_feats), ptr(feats),
                                                  ptr(reduced_feats), ptr(coors_idx), ptr(reduce_count), int(n),
                                                  int(m), int(c), code, ptr(ws), ctypes.c_size_t(nbytes),
                                                  cur_stream(dev)))
