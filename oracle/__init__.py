"""CPU oracle for the BEV front-end hot path -- TEST INFRASTRUCTURE ONLY.

numpy-facing wrappers over ``oracle/libbevfront_oracle.so`` (plain C, see ``bevfront_oracle.c`` for the
reference file:line each function restates).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as the
checker / the timed CPU arm.  The product package never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbevfront_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "bevfront_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_bev_pool_aux.restype = ctypes.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return int(lib().oracle_num_threads())


def grid_size(voxel_size, coors_range):
    g = np.zeros(3, np.int32)
    lib().oracle_grid_size(_p(_f32(voxel_size), _f32p), _p(_f32(coors_range), _f32p), _p(g, _i32p))
    return g


def dynamic_voxelize(points, voxel_size, coors_range, gpu_partial=True, coors_init=None):
    points = _f32(points)
    n, c = points.shape
    coors = np.zeros((n, 3), np.int32) if coors_init is None else _i32(coors_init).copy()
    lib().oracle_dynamic_voxelize(_p(points, _f32p), n, c, _p(_f32(voxel_size), _f32p),
                                  _p(_f32(coors_range), _f32p), _p(coors, _i32p), int(gpu_partial))
    return coors


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels):
    """-> (voxels[M,max_points,C], coors[M,3] xyz, num_points_per_voxel[M])"""
    points = _f32(points)
    n, c = points.shape
    voxels = np.zeros((max_voxels, max_points, c), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    npv = np.zeros((max_voxels,), np.int32)
    m = lib().oracle_hard_voxelize(_p(points, _f32p), n, c, _p(_f32(voxel_size), _f32p),
                                   _p(_f32(coors_range), _f32p), int(max_points), int(max_voxels),
                                   _p(voxels, _f32p), _p(coors, _i32p), _p(npv, _i32p))
    assert m >= 0
    return voxels[:m], coors[:m], npv[:m]


def voxel_mean(voxels, npv):
    voxels = _f32(voxels)
    m, mp, c = voxels.shape
    out = np.zeros((m, c), np.float32)
    lib().oracle_voxel_mean(_p(voxels, _f32p), _p(_i32(npv), _i32p), m, mp, c, _p(out, _f32p))
    return out


_REDUCE = {"sum": 0, "mean": 1, "max": 2}


def dynamic_scatter(feats, coors, reduce_type="max"):
    """-> (reduced[M,C], out_coors[M,ndim], coors_map[N], reduce_count[M])"""
    feats = _f32(feats)
    coors = _i32(coors)
    n, c = feats.shape
    ndim = coors.shape[1]
    reduced = np.zeros((max(n, 1), c), np.float32)
    out_coors = np.zeros((max(n, 1), ndim), np.int32)
    cmap = np.zeros((max(n, 1),), np.int32)
    cnt = np.zeros((max(n, 1),), np.int32)
    m = lib().oracle_dynamic_scatter(_p(feats, _f32p), _p(coors, _i32p), n, c, ndim, _REDUCE[reduce_type],
                                     _p(reduced, _f32p), _p(out_coors, _i32p), _p(cmap, _i32p), _p(cnt, _i32p))
    return reduced[:m], out_coors[:m], cmap[:n], cnt[:m]


def dynamic_scatter_backward(grad_reduced, feats, reduced, coors_map, reduce_count, reduce_type):
    feats = _f32(feats)
    n, c = feats.shape
    m = reduced.shape[0]
    g = np.zeros((n, c), np.float32)
    lib().oracle_dynamic_scatter_backward(_p(g, _f32p), _p(_f32(grad_reduced), _f32p), _p(feats, _f32p),
                                          _p(_f32(reduced), _f32p), _p(_i32(coors_map), _i32p),
                                          _p(_i32(reduce_count), _i32p), n, m, c, _REDUCE[reduce_type])
    return g


def intervals_from_ranks(ranks):
    """bev_pool.py:46-55: interval starts / lengths from sorted ranks."""
    ranks = np.asarray(ranks)
    n = ranks.shape[0]
    if n == 0:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    kept = np.ones(n, bool)
    kept[1:] = ranks[1:] != ranks[:-1]
    starts = np.nonzero(kept)[0].astype(np.int32)
    lengths = np.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = n - starts[-1]
    return starts, lengths


def bev_pool_forward(x, geom, lengths, starts, b, d, h, w):
    """bev_pool_ext.bev_pool_forward restated -> out[b,d,h,w,c]"""
    x = _f32(x)
    c = x.shape[1]
    out = np.zeros((b, d, h, w, c), np.float32)
    lib().oracle_bev_pool_forward(_p(x, _f32p), _p(_i32(geom), _i32p), _p(_i32(lengths), _i32p),
                                  _p(_i32(starts), _i32p), len(starts), c, b, d, h, w, _p(out, _f32p))
    return out


def bev_pool_backward(out_grad, geom, lengths, starts, b, d, h, w):
    out_grad = _f32(out_grad)
    c = out_grad.shape[-1]
    geom = _i32(geom)
    x_grad = np.zeros((geom.shape[0], c), np.float32)
    lib().oracle_bev_pool_backward(_p(out_grad, _f32p), _p(geom, _i32p), _p(_i32(lengths), _i32p),
                                   _p(_i32(starts), _i32p), len(starts), c, b, d, h, w, _p(x_grad, _f32p))
    return x_grad


def bev_pool(feats, coords, ranks, B, D, H, W):
    """ops/bev_pool/bev_pool.py:146-172 restated -> [B,C,D,H,W]"""
    starts, lengths = intervals_from_ranks(ranks)
    out = bev_pool_forward(feats, coords, lengths, starts, B, D, H, W)
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3))


def bev_pool_aux(geom, B, bx, dx, nx):
    """depth_lss.py:118-176 -> (geom_feats[nk,4] i64, kept[N'] bool, ranks[nk] i64, indices[nk] i64)"""
    geom = _f32(geom).reshape(-1, 3)
    nprime = geom.shape[0]
    g = np.zeros((nprime, 4), np.int64)
    kept = np.zeros(nprime, np.uint8)
    ranks = np.zeros(nprime, np.int64)
    indices = np.zeros(nprime, np.int64)
    nxa = np.ascontiguousarray(nx, dtype=np.int64)
    nk = lib().oracle_bev_pool_aux(_p(geom, _f32p), ctypes.c_int64(nprime), int(B), _p(_f32(bx), _f32p),
                                   _p(_f32(dx), _f32p), _p(nxa, _i64p), _p(g, _i64p), _p(kept, _u8p),
                                   _p(ranks, _i64p), _p(indices, _i64p))
    return g[:nk], kept.astype(bool), ranks[:nk], indices[:nk]


def bev_pool_fused(depth, ctx, src, geom4, starts, lengths, B, nz, nx, ny):
    """Reference view-transform data path (outer product + gathers + K1 + permute + collapse-Z).
    depth [BN,D,fH,fW], ctx [BN,C,fH,fW] -> [B, C*nz, nx, ny]"""
    depth = _f32(depth)
    ctx = _f32(ctx)
    BN, D, fH, fW = depth.shape
    C = ctx.shape[1]
    out = np.zeros((B, C, nz, nx, ny), np.float32)
    src = np.ascontiguousarray(src, dtype=np.int64)
    lib().oracle_bev_pool_fused(_p(depth, _f32p), _p(ctx, _f32p), _p(src, _i64p), _p(_i32(geom4), _i32p),
                                _p(_i32(starts), _i32p), _p(_i32(lengths), _i32p), len(starts), BN, C, D, fH, fW,
                                B, nz, nx, ny, _p(out, _f32p))
    # depth_lss.py:202 torch.cat(x.unbind(dim=2), 1): the collapsed channel index is z*C + ch (z-major)
    return np.ascontiguousarray(out.transpose(0, 2, 1, 3, 4)).reshape(B, nz * C, nx, ny)


def lidar_depth_image(points, laug_trans, laug_inv_rot, lidar2image, img_aug, H, W):
    """depth_lss.py:372-420 for one sample -> depth [n_cams, H, W]"""
    points = _f32(points)
    l2i, ia = _f32(lidar2image), _f32(img_aug)
    n_cams = l2i.shape[0]
    depth = np.zeros((n_cams, H, W), np.float32)
    lib().oracle_lidar_depth_image(_p(points, _f32p), points.shape[0], points.shape[1], _p(_f32(laug_trans), _f32p),
                                   _p(_f32(laug_inv_rot), _f32p), _p(l2i, _f32p), _p(ia, _f32p), n_cams, H, W,
                                   _p(depth, _f32p))
    return depth


def depth_histogram(depth, fH, fW, D, dbound):
    """depth_lss.py:632-661: depth [bn, H, W] -> (counts, distr) [bn, fH, fW, D]"""
    depth = _f32(depth)
    bn, H, W = depth.shape
    counts = np.zeros((bn, fH, fW, D), np.float32)
    distr = np.zeros_like(counts)
    f = ctypes.c_float
    lib().oracle_depth_histogram(_p(depth, _f32p), bn, H, W, fH, fW, D, f(dbound[0]), f(dbound[1]), f(dbound[2]),
                                 _p(counts, _f32p), _p(distr, _f32p))
    return counts, distr


def _i3(v):
    if isinstance(v, int):
        v = (v, v, v)
    return np.ascontiguousarray(v, dtype=np.int32)


def spconv_out_shape(shape, ksize, stride, padding, dilation):
    o = np.zeros(3, np.int32)
    lib().oracle_spconv_out_shape(_p(_i3(shape), _i32p), _p(_i3(ksize), _i32p), _p(_i3(stride), _i32p),
                                  _p(_i3(padding), _i32p), _p(_i3(dilation), _i32p), _p(o, _i32p))
    return o


def spconv_rulebook(indices, spatial_shape, ksize, stride=1, padding=0, dilation=1, subm=False):
    """-> (out_indices[n_out,4], pair_fwd[kv,n_out], out_shape[3])"""
    indices = _i32(indices)
    n_in = indices.shape[0]
    k, s, p, d = _i3(ksize), _i3(stride), _i3(padding), _i3(dilation)
    shape = _i3(spatial_shape)
    kv = int(np.prod(k))
    if subm:
        out_idx = indices.copy()
        out_shape = shape.copy()
    else:
        out_shape = spconv_out_shape(shape, k, s, p, d)
        buf = np.zeros((max(n_in * kv, 1), 4), np.int32)
        n_out = lib().oracle_spconv_out_sites(_p(indices, _i32p), n_in, _p(shape, _i32p), _p(k, _i32p),
                                              _p(s, _i32p), _p(p, _i32p), _p(d, _i32p), _p(buf, _i32p))
        out_idx = buf[:n_out].copy()
    n_out = out_idx.shape[0]
    pair = np.full((kv, max(n_out, 1)), -1, np.int32)
    if n_out:
        lib().oracle_spconv_rulebook(_p(indices, _i32p), n_in, _p(out_idx, _i32p), n_out, _p(shape, _i32p),
                                     _p(k, _i32p), _p(s, _i32p), _p(p, _i32p), _p(d, _i32p), int(subm),
                                     _p(pair, _i32p))
    return out_idx, pair[:, :n_out], out_shape


def strided_chain_boxes(indices, spatial_shape, convs):
    """Numpy / pure-Python restatement of the library's one-pass site construction for a CHAIN of strided convolutions
    (csrc/spconv_rulebook.cu: mark_levels_kernel): the box a level-0 site reaches on level l follows per axis from its box on
    level l-1 -- ceil((i0 + p - (k-1)) / s) .. floor((i1 + p) / s), clipped to the grid -- and level l's site set is the
    union of the boxes.  convs = [(ksize, stride, padding), ...]; -> per level the sorted [n_l, 4] sites.  The test checks it
    against spconv_rulebook applied level after level (the definition).  Small inputs only (Python loops)."""
    indices = np.asarray(indices, np.int32).reshape(-1, 4)
    shapes, shape = [], [int(v) for v in spatial_shape]
    for ks, st, pd in convs:
        shape = [int(v) for v in spconv_out_shape(shape, ks, st, pd, (1, 1, 1))]
        shapes.append(shape)
    sets = [set() for _ in convs]
    for b, x, y, z in indices.tolist():
        lo, hi = [x, y, z], [x, y, z]
        for l, (ks, st, pd) in enumerate(convs):
            empty = False
            for a in range(3):
                k, s, p = int(ks[a]), int(st[a]), int(pd[a])
                nlo = max(-((-(lo[a] + p - (k - 1))) // s), 0)          # ceil division
                nhi = min((hi[a] + p) // s, shapes[l][a] - 1)            # floor division
                lo[a], hi[a] = nlo, nhi
                empty |= nlo > nhi
            if empty:
                break
            for xx in range(lo[0], hi[0] + 1):
                for yy in range(lo[1], hi[1] + 1):
                    for zz in range(lo[2], hi[2] + 1):
                        sets[l].add((b, xx, yy, zz))
    return [(np.array(sorted(s_), np.int32).reshape(-1, 4), shp) for s_, shp in zip(sets, shapes)]


def spconv_gemm(feats, weight, pair_fwd, bias=None, acc_double=True):
    """weight [Cout, kD, kH, kW, Cin] (spconv-2.x layout) -> out[n_out, Cout]"""
    feats = _f32(feats)
    weight = _f32(weight)
    cout, cin = weight.shape[0], weight.shape[-1]
    kv = int(np.prod(weight.shape[1:-1]))
    pair_fwd = _i32(pair_fwd)
    n_out = pair_fwd.shape[1]
    out = np.zeros((n_out, cout), np.float32)
    b = None if bias is None else _p(_f32(bias), _f32p)
    if n_out:
        lib().oracle_spconv_gemm(_p(feats, _f32p), _p(weight, _f32p), b, _p(pair_fwd, _i32p), n_out, kv, cin,
                                 cout, int(acc_double), _p(out, _f32p))
    return out


def spconv_backward(feats, weight, pair_fwd, grad_out):
    """Gradients of spconv_gemm (numpy, float64): the transpose of the forward restatement above -- spconv's backward
    lives in the third-party package (not in /root/reference), so this is the definition
    d_feats[pair[k, j]] += grad_out[j] @ W[:, k, :],  d_W[:, k, :] = grad_out[valid]^T @ feats[pair[k, valid]].
    weight [Cout, kD, kH, kW, Cin] -> (d_feats [n_in, Cin], d_weight like weight)."""
    feats = np.asarray(feats, np.float64)
    grad_out = np.asarray(grad_out, np.float64)
    w = np.asarray(weight, np.float64)
    cout, cin = w.shape[0], w.shape[-1]
    kv = int(np.prod(w.shape[1:-1]))
    w3 = w.reshape(cout, kv, cin)
    d_feats = np.zeros_like(feats)
    d_w = np.zeros_like(w3)
    pair_fwd = np.asarray(pair_fwd)
    for k in range(kv):
        valid = np.nonzero(pair_fwd[k] >= 0)[0]
        if valid.size == 0:
            continue
        rows = pair_fwd[k][valid]
        np.add.at(d_feats, rows, grad_out[valid] @ w3[:, k, :])
        d_w[:, k, :] = grad_out[valid].T @ feats[rows]
    return d_feats, d_w.reshape(w.shape)


def sparse_to_dense(feats, indices, batch_size, spatial_shape):
    feats = _f32(feats)
    n, c = feats.shape
    shape = _i3(spatial_shape)
    dense = np.zeros((batch_size, c, int(shape[0]), int(shape[1]), int(shape[2])), np.float32)
    lib().oracle_sparse_to_dense(_p(feats, _f32p), _p(_i32(indices), _i32p), n, c, batch_size, _p(shape, _i32p),
                                 _p(dense, _f32p))
    return dense


def bn_relu(x, gamma, beta, mean, var, eps, residual=None, relu=True):
    x = _f32(x).copy()
    n, c = x.shape
    r = None if residual is None else _p(_f32(residual), _f32p)
    lib().oracle_bn_relu(_p(x, _f32p), r, n, c, _p(_f32(gamma), _f32p), _p(_f32(beta), _f32p),
                         _p(_f32(mean), _f32p), _p(_f32(var), _f32p), ctypes.c_float(eps), int(relu))
    return x
