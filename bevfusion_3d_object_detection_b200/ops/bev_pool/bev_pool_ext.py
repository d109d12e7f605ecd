"""`bev_pool_ext` -- same entry points as the reference pybind module
projects/BEVFusion/bevfusion/ops/bev_pool/src/bev_pool.cpp:89-94 (bev_pool_forward, bev_pool_backward;
argument order `interval_lengths` before `interval_starts` as in the reference), over libbevfront_b200.
Unlike the reference (raw data_ptr casts, bev_pool.cpp:32-35) the inputs are validated, and kernels run on
the CURRENT stream rather than the legacy default stream (bev_pool_cuda.cu:88).
"""
import ctypes

import torch

from ..._lib import check, cur_stream, lib, ptr


def _chk(t, name, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (the B200 build ships no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def bev_pool_forward(x, geom_feats, interval_lengths, interval_starts, b, d, h, w):
    """x[n,c] f32 (rows of one interval contiguous), geom_feats[n,4] i32 (x,y,z,batch) -> out[b,d,h,w,c]"""
    _chk(x, "x", torch.float32)
    _chk(geom_feats, "geom_feats", torch.int32)
    _chk(interval_lengths, "interval_lengths", torch.int32)
    _chk(interval_starts, "interval_starts", torch.int32)
    n, c = x.shape
    if geom_feats.shape[0] != n or geom_feats.shape[1] != 4:
        raise RuntimeError(f"geom_feats must be [{n}, 4], got {tuple(geom_feats.shape)}")
    n_int = interval_lengths.shape[0]
    b, d, h, w = int(b), int(d), int(h), int(w)
    with torch.cuda.device(x.device):
        L = lib()
        out = torch.empty((b, d, h, w, c), dtype=x.dtype, device=x.device)
        nbytes = L.bevf_bev_pool_workspace_bytes(int(n), int(c))
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=x.device)
        check(L.bevf_bev_pool_forward(ptr(x), ptr(geom_feats), ptr(interval_lengths), ptr(interval_starts), int(n),
                                      int(c), int(n_int), b, d, h, w, ptr(out), ptr(ws), ctypes.c_size_t(nbytes),
                                      cur_stream(x.device)))
    return out


def bev_pool_backward(out_grad, geom_feats, interval_lengths, interval_starts, b, d, h, w):
    """out_grad[b,d,h,w,c] -> x_grad[n,c]"""
    _chk(out_grad, "out_grad", torch.float32)
    _chk(geom_feats, "geom_feats", torch.int32)
    _chk(interval_lengths, "interval_lengths", torch.int32)
    _chk(interval_starts, "interval_starts", torch.int32)
    n = geom_feats.shape[0]
    c = out_grad.shape[4]
    n_int = interval_lengths.shape[0]
    b, d, h, w = int(b), int(d), int(h), int(w)
    with torch.cuda.device(out_grad.device):
        L = lib()
        x_grad = torch.empty((n, c), dtype=out_grad.dtype, device=out_grad.device)
        nbytes = L.bevf_bev_pool_workspace_bytes(int(n), int(c))
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=out_grad.device)
        check(L.bevf_bev_pool_backward(ptr(out_grad), ptr(geom_feats), ptr(interval_lengths), ptr(interval_starts),
                                       int(n), int(c), int(n_int), b, d, h, w, ptr(x_grad), ptr(ws),
                                       ctypes.c_size_t(nbytes), cur_stream(out_grad.device)))
    return x_grad
