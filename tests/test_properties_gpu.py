"""Randomised differential tests (hypothesis): the CUDA path against the CPU oracle on many small, adversarial
configurations the fixed cases do not enumerate -- tiny grids, caps that bite, points on cell boundaries, odd kernel /
stride / padding / dilation combinations, duplicate-heavy clouds.  Bar: bit-exact for coordinates, counts, voxels and
rulebooks; fp32 sparse-conv features 1e-5 of the output scale."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from bevfusion_3d_object_detection_b200 import ops, spconv
from bevfusion_3d_object_detection_b200.spconv import functional as Fsp

pytestmark = pytest.mark.gpu

COMMON = dict(deadline=None, max_examples=100, suppress_health_check=list(HealthCheck), derandomize=True)


@st.composite
def voxel_cases(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    n = draw(st.integers(0, 3000))
    c = draw(st.sampled_from([3, 4, 5, 7]))
    grid = [draw(st.integers(1, 12)) for _ in range(3)]
    vs = [float(draw(st.sampled_from([0.25, 0.5, 1.0, 0.075, 0.3]))) for _ in range(3)]
    lo = [float(draw(st.sampled_from([0.0, -1.5, -54.0, 2.0]))) for _ in range(3)]
    cr = lo + [lo[j] + grid[j] * vs[j] for j in range(3)]
    mode = draw(st.sampled_from(["uniform", "clustered", "lattice"]))
    span = np.array([grid[j] * vs[j] for j in range(3)], np.float64)
    if mode == "uniform":        # some points fall outside the range on purpose
        xyz = rng.uniform(-0.1, 1.1, (n, 3)) * span + np.array(lo)
    elif mode == "clustered":    # many points per voxel: max_points bites
        centres = rng.uniform(0, 1, (max(1, n // 50 + 1), 3)) * span + np.array(lo)
        xyz = centres[rng.integers(0, centres.shape[0], n)] + rng.normal(0, 0.3, (n, 3)) * np.array(vs)
    else:                        # exactly on cell boundaries (floor / fp32 division edge cases), incl. the upper range edge
        k = rng.integers(-1, np.array(grid) + 2, (n, 3))
        xyz = np.array(lo) + k * np.array(vs)
    pts = np.concatenate([xyz, rng.standard_normal((n, c - 3))], 1).astype(np.float32)
    max_points = draw(st.sampled_from([1, 2, 5, 10, 35]))          # 35 > the register-list width of any fast path
    max_voxels = draw(st.sampled_from([1, 3, 50, 100000]))
    return pts, vs, cr, max_points, max_voxels


@settings(**COMMON)
@given(voxel_cases())
def test_hard_voxelize_random_configurations(oracle_mod, case):
    pts, vs, cr, mp, mv = case
    ov, oc, on = oracle_mod.hard_voxelize(pts, vs, cr, mp, mv)
    v, c, n = ops.voxelization(torch.from_numpy(pts).cuda(), list(vs), list(cr), mp, mv, True)
    np.testing.assert_array_equal(c.cpu().numpy(), oc)
    np.testing.assert_array_equal(n.cpu().numpy(), on)
    np.testing.assert_array_equal(v.cpu().numpy(), ov)


@st.composite
def conv_cases(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    subm = draw(st.booleans())
    ks = tuple(draw(st.sampled_from([1, 2, 3])) for _ in range(3))
    if subm:
        ks = tuple(k if k % 2 == 1 else 3 for k in ks)           # SubM kernels are centred: odd sizes
        stride, pad = (1, 1, 1), tuple(k // 2 for k in ks)
    else:
        stride = tuple(draw(st.sampled_from([1, 2, 3])) for _ in range(3))
        pad = tuple(draw(st.integers(0, 1)) for _ in range(3))
    dil = tuple(draw(st.sampled_from([1, 1, 2])) for _ in range(3))
    shape = tuple(draw(st.integers(max(3, d * (k - 1) + 1), 14)) for k, d in zip(ks, dil))
    batch = draw(st.integers(1, 3))
    cells = batch * shape[0] * shape[1] * shape[2]
    n = draw(st.integers(1, min(cells, 600)))
    lin = rng.choice(cells, size=n, replace=False)
    if draw(st.booleans()):
        lin = np.sort(lin)
    idx = np.stack([lin // (shape[0] * shape[1] * shape[2]), (lin // (shape[1] * shape[2])) % shape[0],
                    (lin // shape[2]) % shape[1], lin % shape[2]], 1).astype(np.int32)
    cin, cout = draw(st.sampled_from([(3, 5), (4, 16), (16, 8), (5, 16)]))
    return seed, subm, ks, stride, pad, dil, shape, batch, idx, cin, cout


@settings(**COMMON)
@given(conv_cases())
def test_rulebook_and_fp32_conv_random_geometries(oracle_mod, case):
    seed, subm, ks, stride, pad, dil, shape, batch, idx, cin, cout = case
    o_idx, o_pair, o_shape = oracle_mod.spconv_rulebook(idx, shape, ks, stride, pad, dil, subm)
    rng = np.random.default_rng(seed ^ 0x5EED)
    feats = rng.standard_normal((idx.shape[0], cin)).astype(np.float32)
    x = spconv.SparseConvTensor(torch.from_numpy(feats).cuda(), torch.from_numpy(idx).cuda(), list(shape), batch)
    if any(o <= 0 for o in o_shape):
        return
    datas = Fsp.get_indice_pairs(x, ks, stride, pad, dil, subm)
    assert x.coord_index().error_code() == 0
    assert datas.n_out == o_idx.shape[0] and list(datas.out_spatial_shape) == [int(v) for v in o_shape]
    np.testing.assert_array_equal(datas.out_indices.cpu().numpy()[:datas.n_out], o_idx)
    np.testing.assert_array_equal(datas.pair_fwd.cpu().numpy(), o_pair)
    cls = spconv.SubMConv3d if subm else spconv.SparseConv3d
    conv = cls(cin, cout, ks, stride=stride, padding=pad, dilation=dil, bias=True, precision="fp32").cuda()
    with torch.no_grad():
        got = conv(x).features.cpu().numpy()
    ref = oracle_mod.spconv_gemm(feats, conv.weight.detach().cpu().numpy(), o_pair) + conv.bias.detach().cpu().numpy()
    assert got.shape == ref.shape
    if ref.size:
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


@st.composite
def chain_cases(draw):
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    nlev = draw(st.integers(1, 4))
    convs = []
    for _ in range(nlev):
        ks = tuple(draw(st.sampled_from([1, 2, 3])) for _ in range(3))
        stride = tuple(draw(st.sampled_from([1, 2, 3])) for _ in range(3))
        pad = tuple(draw(st.integers(0, 1)) for _ in range(3))
        convs.append((ks, stride, pad))
    shape = tuple(draw(st.integers(6, 40)) for _ in range(3))
    batch = draw(st.integers(1, 3))
    cells = batch * shape[0] * shape[1] * shape[2]
    n = draw(st.integers(0, min(cells, 1500)))
    lin = rng.choice(cells, size=n, replace=False)            # unsorted: the chain takes rows in any order
    idx = np.stack([lin // (shape[0] * shape[1] * shape[2]), (lin // (shape[1] * shape[2])) % shape[0],
                    (lin // shape[2]) % shape[1], lin % shape[2]], 1).astype(np.int32).reshape(-1, 4)
    return convs, shape, batch, idx


@settings(**COMMON)
@given(chain_cases())
def test_strided_sites_chain_equals_level_by_level(oracle_mod, case):
    """bevf_spconv_strided_sites_chain (all levels from the level-0 coordinates: per-axis interval composition) against the
    oracle's rulebook applied level after level, on random chains of strided convolutions (kernel 1-3, stride 1-3, padding
    0-1 per axis: includes kernels smaller than the stride, i.e. gaps) and unsorted input rows."""
    convs, shape, batch, idx = case
    cur_shape, ok = list(shape), True
    for ks, stride, pad in convs:                         # skip chains whose grid collapses
        cur_shape = [(cur_shape[j] + 2 * pad[j] - (ks[j] - 1) - 1) // stride[j] + 1 for j in range(3)]
        ok &= all(v >= 1 for v in cur_shape)
    if not ok:
        return
    got = Fsp.strided_sites_chain(torch.from_numpy(idx).cuda(), batch, list(shape), convs)
    cur_idx, cur_shape = idx, list(shape)
    for (ks, stride, pad), (g_idx, g_shape) in zip(convs, got):
        if cur_idx.shape[0]:
            o_idx, _, o_shape = oracle_mod.spconv_rulebook(cur_idx, cur_shape, ks, stride, pad, (1, 1, 1), False)
        else:
            o_idx = np.zeros((0, 4), np.int32)
            o_shape = [(cur_shape[j] + 2 * pad[j] - (ks[j] - 1) - 1) // stride[j] + 1 for j in range(3)]
        assert [int(v) for v in g_shape] == [int(v) for v in o_shape]
        np.testing.assert_array_equal(g_idx.cpu().numpy(), o_idx)
        cur_idx, cur_shape = o_idx, [int(v) for v in o_shape]
