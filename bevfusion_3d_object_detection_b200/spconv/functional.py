"""Rulebook generation and gather-GEMM-scatter: the functional boundary spconv exposes as
SpconvOps.get_indice_pairs_implicit_gemm / ConvGemmOps.implicit_gemm (reference call sites
projects/SparseConvolution/sparse_functional.py:118-137, 287-314), on libbevfront_b200.
"""
import ctypes

import torch

from .._lib import check, cur_stream, i32_array, lib, ptr
from .core import CoordIndex, IndicePair


# bench.py sets this to a list to collect (start_event, end_event, useful_flops) per GEMM launch; the flop count
# costs a reduction + host sync per layer, so it is None on the product path
GEMM_TIMING = None


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return tuple(int(x) for x in v)
    return (int(v),) * 3


def _pad4(n):
    return max(4, (int(n) + 3) // 4 * 4)


def conv_out_shape(spatial_shape, ksize, stride, padding, dilation):
    out = (ctypes.c_int * 3)()
    lib().bevf_spconv_out_shape(i32_array(spatial_shape), i32_array(ksize), i32_array(stride), i32_array(padding),
                                i32_array(dilation), out)
    return [int(v) for v in out]


def get_indice_pairs(x, ksize, stride, padding, dilation, subm):
    """x: SparseConvTensor -> IndicePair.  SubM: out sites == in sites (same order).  Strided: out sites in
    ascending linear order; ONE host sync to learn n_out (spconv returns num_act_out to the host as well)."""
    ksize, stride, padding, dilation = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    dev = x.indices.device
    L = lib()
    kv = ksize[0] * ksize[1] * ksize[2]
    n_in = x.indices.shape[0]
    in_index = x.coord_index()
    with torch.cuda.device(dev):
        st = cur_stream(dev)
        if subm:
            ld = _pad4(n_in)  # 16-byte aligned rulebook rows: the GEMM producer reads four entries per load
            pair = torch.empty((kv, ld), dtype=torch.int32, device=dev)
            check(L.bevf_spconv_subm_rulebook(ptr(x.indices), int(n_in), None, x.batch_size, in_index.shape_c,
                                              i32_array(ksize), i32_array(dilation), ptr(in_index.mem),
                                              ctypes.c_size_t(in_index.nbytes), ptr(in_index.perm), ptr(pair),
                                              int(ld), st))
            return IndicePair(x.indices, pair[:, :n_in], n_in, list(x.spatial_shape), in_index, in_index, ksize,
                              (1, 1, 1), padding, dilation, True)
        out_shape = conv_out_shape(x.spatial_shape, ksize, stride, padding, dilation)
        reach = 1
        for k, s, d in zip(ksize, stride, dilation):
            reach *= min(k, -(-k // s)) if d == 1 else k   # gcd(dilation, stride) > 1: up to k outputs per axis
        cells = x.batch_size * out_shape[0] * out_shape[1] * out_shape[2]
        cap = max(1, min(n_in * reach, cells))
        out_bytes = int(L.bevf_spconv_index_bytes(x.batch_size, i32_array(out_shape)))
        if out_bytes == 0:
            raise RuntimeError("sparse grid not supported: " + L.bevf_last_error().decode())
        out_mem = torch.empty(out_bytes, dtype=torch.uint8, device=dev)
        out_indices = torch.empty((cap, 4), dtype=torch.int32, device=dev)
        n_out_dev = torch.empty(1, dtype=torch.int32, device=dev)
        check(L.bevf_spconv_strided_sites(ptr(x.indices), int(n_in), None, x.batch_size, in_index.shape_c, i32_array(ksize),
                                          i32_array(stride), i32_array(padding), i32_array(dilation), ptr(out_mem),
                                          ctypes.c_size_t(out_bytes), ptr(out_indices), int(cap), ptr(n_out_dev), st))
        n_out = int(n_out_dev.item())
        assert n_out <= cap
        out_indices = out_indices[:n_out]
        ld = _pad4(n_out)
        pair = torch.empty((kv, ld), dtype=torch.int32, device=dev)
        check(L.bevf_spconv_strided_rulebook(ptr(out_indices), int(n_out), None, x.batch_size, in_index.shape_c,
                                             i32_array(ksize), i32_array(stride), i32_array(padding),
                                             i32_array(dilation), ptr(in_index.mem), ctypes.c_size_t(in_index.nbytes),
                                             ptr(in_index.perm), ptr(pair), int(ld), st))
        out_index = CoordIndex(out_indices, x.batch_size, out_shape, mem=out_mem)
        return IndicePair(out_indices, pair[:, :n_out] if n_out else pair[:, :0], n_out, out_shape, in_index,
                          out_index, ksize, stride, padding, dilation, False)


def strided_sites_chain(indices, batch_size, spatial_shape, convs):
    """Output sites of a chain of strided convolutions (dilation 1), all levels in three launches
    (bevf_spconv_strided_sites_chain).  indices [n, 4] int32 (b, x, y, z) in any order; convs = [(ksize, stride, padding),
    ...].  -> [(out_indices [n_l, 4] in ascending cell order, out_shape)] per level: what get_indice_pairs(...).out_indices
    gives when the convolutions are applied one after the other."""
    dev = indices.device
    L = lib()
    n = len(convs)
    shapes, shape = [], list(spatial_shape)
    for ks, st_, pd in convs:
        shape = conv_out_shape(shape, ks, st_, pd, (1, 1, 1))
        shapes.append(shape)
    n_in = int(indices.shape[0])
    caps, cap = [], max(1, n_in)
    for (ks, st_, pd), shp in zip(convs, shapes):
        reach = 1
        for k, s_ in zip(ks, st_):
            reach *= min(k, -(-k // s_))
        cap = max(1, min(cap * reach, batch_size * shp[0] * shp[1] * shp[2]))
        caps.append(cap)
    mems, nbytes = [], []
    for shp in shapes:
        b = int(L.bevf_spconv_index_bytes(batch_size, i32_array(shp)))
        if b == 0:
            raise RuntimeError("sparse grid not supported: " + L.bevf_last_error().decode())
        nbytes.append(b)
        mems.append(torch.empty(b, dtype=torch.uint8, device=dev))
    outs = [torch.empty((c, 4), dtype=torch.int32, device=dev) for c in caps]
    ndev = torch.zeros(n, dtype=torch.int32, device=dev)
    flat = lambda j: (ctypes.c_int * (3 * n))(*[int(v) for c in convs for v in c[j]])
    with torch.cuda.device(dev):
        check(L.bevf_spconv_strided_sites_chain(
            ptr(indices), n_in, None, int(batch_size), i32_array(list(spatial_shape)), n, flat(0), flat(1), flat(2),
            (ctypes.c_void_p * n)(*[m.data_ptr() for m in mems]), (ctypes.c_size_t * n)(*nbytes),
            (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs]), (ctypes.c_int * n)(*caps),
            (ctypes.c_void_p * n)(*[ndev[l:].data_ptr() for l in range(n)]), cur_stream(dev)))
    counts = ndev.tolist()
    assert all(c <= cap for c, cap in zip(counts, caps))
    return [(o[:c], shp) for o, c, shp in zip(outs, counts, shapes)]


def pack_weight_f32(weight):
    """[Cout, kD, kH, kW, Cin] fp32 -> [kv, Cin, Cout] fp32."""
    cout, cin = weight.shape[0], weight.shape[-1]
    kv = weight.numel() // (cout * cin)
    w = weight.detach().contiguous().float()
    out = torch.empty((kv, cin, cout), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        check(lib().bevf_spconv_pack_weight_f32(ptr(w), ptr(out), int(kv), int(cin), int(cout), cur_stream(w.device)))
    return out


def pack_weight_bf16(weight):
    """[Cout, kD, kH, kW, Cin] fp32 -> per-tap UMMA core-matrix image, bf16 [kv, Cout*cin_pad]."""
    cout, cin = weight.shape[0], weight.shape[-1]
    kv = weight.numel() // (cout * cin)
    L = lib()
    cin_pad = L.bevf_spconv_tc_cin_pad(int(cin))
    w = weight.detach().contiguous().float()
    L.bevf_spconv_packed_weight_bytes.restype = ctypes.c_longlong
    nbytes = int(L.bevf_spconv_packed_weight_bytes(int(kv), int(cin), int(cout)))
    out = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)   # UMMA image [+ fragment-order image]
    with torch.cuda.device(w.device):
        check(L.bevf_spconv_pack_weight_bf16(ptr(w), ptr(out), int(kv), int(cin), int(cout), cur_stream(w.device)))
    return out


def cast_features_bf16(features, cin_pad):
    n, cin = features.shape
    f = features.contiguous().float()
    out = torch.empty((n, cin_pad), dtype=torch.bfloat16, device=f.device)
    with torch.cuda.device(f.device):
        check(lib().bevf_spconv_cast_bf16(ptr(f), ptr(out), int(n), int(cin), int(cin_pad), None, cur_stream(f.device)))
    return out


def pair_bwd(pair_fwd, n_out, n_in):
    """Inverse rulebook [kv, n_in]: pair_bwd[k, i] = j where pair_fwd[k, j] = i, -1 elsewhere."""
    kv = pair_fwd.shape[0]
    out = torch.empty((kv, max(n_in, 1)), dtype=torch.int32, device=pair_fwd.device)
    ld = pair_fwd.stride(0) if pair_fwd.shape[1] > 0 else max(n_out, 1)
    with torch.cuda.device(pair_fwd.device):
        check(lib().bevf_spconv_pair_bwd(ptr(pair_fwd), int(ld), int(n_out), None, int(kv), ptr(out), int(out.stride(0)),
                                         int(n_in), cur_stream(pair_fwd.device)))
    return out[:, :n_in]


def wgrad_f32(features, grad_out, pair_fwd, n_out, kv, cin, cout):
    """d_weight [Cout, kv, Cin] fp32 = sum over valid pairs of grad_out[j]^T (x) features[pair_fwd[k, j]]."""
    f = features.detach().contiguous().float()
    g = grad_out.detach().contiguous().float()
    out = torch.empty((cout, kv, cin), dtype=torch.float32, device=f.device)
    ld = pair_fwd.stride(0) if pair_fwd.shape[1] > 0 else max(n_out, 1)
    with torch.cuda.device(f.device):
        check(lib().bevf_spconv_wgrad_f32(ptr(f), ptr(g), ptr(pair_fwd), int(ld), int(n_out), None, int(kv), int(cin),
                                          int(cout), ptr(out), cur_stream(f.device)))
    return out


def wgrad_bf16(features_bf16, grad_out_bf16, pair_fwd, n_out, kv, cin, cout):
    """d_weight [Cout, kv, Cin] fp32 on the tensor cores: bf16 operands (features [n_in, cin_pad], grad_out
    [n_out, Cout]), fp32 accumulation."""
    assert features_bf16.dtype == torch.bfloat16 and grad_out_bf16.dtype == torch.bfloat16
    dev = features_bf16.device
    out = torch.empty((cout, kv, cin), dtype=torch.float32, device=dev)
    ld = pair_fwd.stride(0) if pair_fwd.shape[1] > 0 else max(n_out, 1)
    with torch.cuda.device(dev):
        check(lib().bevf_spconv_wgrad_bf16(ptr(features_bf16), ptr(grad_out_bf16), ptr(pair_fwd), int(ld), int(n_out),
                                           int(kv), int(cin), int(features_bf16.shape[1]), int(cout), ptr(out),
                                           cur_stream(dev)))
    return out


def tc_supported(cin, cout):
    return bool(lib().bevf_spconv_tc_supported(int(cin), int(cout)))


def implicit_gemm(features, pair_fwd, n_out, weight_packed, kv, cin, cout, precision="fp32", bias=None, bn_scale=None,
                  bn_shift=None, residual=None, relu=False, features_bf16=None, want_bf16=False, want_f32=True,
                  timing_tag="forward"):
    """out[n_out, Cout] fp32 (and optionally its bf16 copy) = epilogue(sum_k feats[pair_fwd[k]] @ W[k]).
    `residual` may be fp32 or (tensor-core path) bf16.  want_f32=False (tensor-core path only) skips the fp32 output."""
    dev = pair_fwd.device
    L = lib()
    want_f32 = want_f32 or precision != "bf16" or not want_bf16
    out = torch.empty((n_out, cout), dtype=torch.float32, device=dev) if want_f32 else None
    out_bf16 = None
    ld = pair_fwd.stride(0) if pair_fwd.shape[1] > 0 else max(n_out, 1)
    residual_bf16 = None
    if residual is not None:
        if residual.dtype == torch.bfloat16 and precision == "bf16":
            residual_bf16, residual = residual.contiguous(), None
            assert residual_bf16.shape[1] == cout
        else:
            residual = residual.contiguous().float()
    if n_out == 0:  # empty level: nothing to launch (zero-sized tensors have no device pointer to pass)
        return out, (torch.empty((0, cout), dtype=torch.bfloat16, device=dev)
                     if (precision == "bf16" and want_bf16) else None)
    timing = GEMM_TIMING
    if timing is not None:
        pairs = int((pair_fwd[:, :n_out] >= 0).sum().item())
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), 2.0 * pairs * cin * cout,
              timing_tag)
    with torch.cuda.device(dev):
        st = cur_stream(dev)
        if timing is not None:
            ev[0].record()
        if precision == "fp32":
            f = features.contiguous().float()
            check(L.bevf_spconv_gemm_f32(ptr(f), ptr(weight_packed), ptr(pair_fwd), int(ld), int(n_out), None, int(kv),
                                         int(cin), int(cout), ptr(bias), ptr(bn_scale), ptr(bn_shift), ptr(residual),
                                         int(bool(relu)), ptr(out), st))
        elif precision == "bf16":
            cin_pad = L.bevf_spconv_tc_cin_pad(int(cin))
            fb = features_bf16 if features_bf16 is not None else cast_features_bf16(features, cin_pad)
            assert fb.shape[1] == cin_pad and fb.dtype == torch.bfloat16
            if want_bf16:
                out_bf16 = torch.empty((n_out, cout), dtype=torch.bfloat16, device=dev)
            check(L.bevf_spconv_gemm_bf16(ptr(fb), int(fb.shape[0]), ptr(weight_packed), ptr(pair_fwd), int(ld), int(n_out), None,
                                          int(kv), int(cin_pad), int(cout), ptr(bias), ptr(bn_scale), ptr(bn_shift),
                                          ptr(residual), ptr(residual_bf16), int(bool(relu)), ptr(out), ptr(out_bf16),
                                          st))
        else:
            raise ValueError(f"unknown precision {precision!r}")
        if timing is not None:
            ev[1].record()
            timing.append(ev)
    return out, out_bf16
