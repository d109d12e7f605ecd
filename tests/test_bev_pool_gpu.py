"""CUDA bev_pool (boundary + fused forms, through the C ABI) against the CPU oracle.
Tolerance: 1e-5 relative (fp32; summation order differs from the reference, whose own argsort is unstable)."""
import numpy as np
import pytest
import torch

from conftest import golden
from bevfusion_3d_object_detection_b200 import ops, synthetic
from bevfusion_3d_object_detection_b200.ops.bev_pool import bev_pool_ext
from bevfusion_3d_object_detection_b200.ops.bev_pool.bev_pool import (QuickCumsum, QuickCumsumCuda,
                                                                      QuickCumsumTrainingCuda)
from bevfusion_3d_object_detection_b200.view_transform import BaseViewTransform

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-5


def _random_sorted_case(rng, n, c, B, D, H, W, max_run=60):
    """rank-sorted rows with skewed interval lengths, reference rank formula (depth_lss.py:169)."""
    cells = []
    while sum(len(x) for x in cells) < n:
        run = int(rng.integers(1, max_run)) if rng.random() < 0.9 else int(rng.integers(max_run, 12 * max_run))
        cells.append(np.full(run, rng.integers(0, B * D * H * W)))
    cell = np.concatenate(cells)[:n]
    b, rem = cell // (D * H * W), cell % (D * H * W)
    z, rem = rem // (H * W), rem % (H * W)
    x, y = rem // W, rem % W
    ranks = x * (W * D * B) + y * (D * B) + z * B + b
    order = np.argsort(ranks, kind="stable")
    geom = np.stack([x, y, z, b], 1)[order].astype(np.int32)
    feats = rng.standard_normal((n, c)).astype(np.float32)
    return feats, geom, ranks[order]


@pytest.mark.parametrize("n,c,B,D,H,W", [(20000, 80, 1, 1, 60, 60), (50000, 80, 2, 2, 30, 40), (5000, 64, 1, 1, 16, 16),
                                         (3000, 6, 1, 1, 8, 8), (7, 80, 1, 1, 4, 4), (40000, 256, 1, 1, 20, 20)])
def test_boundary_forward_backward(oracle_mod, n, c, B, D, H, W):
    rng = np.random.default_rng(n + c)
    feats, geom, ranks = _random_sorted_case(rng, n, c, B, D, H, W)
    x = torch.from_numpy(feats).cuda().requires_grad_(True)
    out = ops.bev_pool(x, torch.from_numpy(geom).cuda().long(), torch.from_numpy(ranks).cuda(), B, D, H, W, True)
    ref = oracle_mod.bev_pool(feats, geom, ranks, B, D, H, W)
    assert out.shape == (B, c, D, H, W) and out.is_contiguous()
    # fp32 sums in a different order than the reference: the error is bounded relative to sum |x_i| of each
    # cell (1e-6 of it), plus the stated 1e-5 relative to the result
    l1 = oracle_mod.bev_pool(np.abs(feats), geom, ranks, B, D, H, W)
    err = np.abs(out.detach().cpu().numpy() - ref)
    assert (err <= RTOL * np.abs(ref) + 1e-6 * l1).all(), float((err - RTOL * np.abs(ref) - 1e-6 * l1).max())
    og = rng.standard_normal(ref.shape).astype(np.float32)
    out.backward(torch.from_numpy(og).cuda())
    starts, lengths = oracle_mod.intervals_from_ranks(ranks)
    want = oracle_mod.bev_pool_backward(np.ascontiguousarray(og.transpose(0, 2, 3, 4, 1)), geom, lengths, starts, B, D,
                                        H, W)
    np.testing.assert_array_equal(x.grad.cpu().numpy(), want)  # pure broadcast: exact
    # inference branch: 0-dim tensor dims like self.nx[i] (depth_lss.py:199)
    out2 = ops.bev_pool(x.detach(), torch.from_numpy(geom).cuda().long(), torch.from_numpy(ranks).cuda(), B,
                        torch.tensor(D), torch.tensor(H), torch.nn.Parameter(torch.tensor(W), requires_grad=False), False)
    np.testing.assert_array_equal(out2.cpu().numpy(), out.detach().cpu().numpy())  # deterministic kernel


def test_ext_generic_path_for_non_tiling_intervals(oracle_mod):
    """bev_pool_forward with (starts, lengths) that do NOT tile [0, n): the reference kernel only sums the rows
    each interval names; rows outside every interval contribute nothing (and get zero gradient)."""
    rng = np.random.default_rng(0)
    n, c = 4000, 80
    feats = rng.standard_normal((n, c)).astype(np.float32)
    starts = np.arange(0, n, 100, dtype=np.int32)
    lengths = rng.integers(1, 90, starts.shape[0]).astype(np.int32)  # gaps after every interval
    geom = np.zeros((n, 4), np.int32)
    geom[:, 0] = np.repeat(np.arange(starts.shape[0]), 100)[:n] // 8
    geom[:, 1] = np.repeat(np.arange(starts.shape[0]), 100)[:n] % 8
    args = [torch.from_numpy(a).cuda() for a in (geom, lengths, starts)]
    out = bev_pool_ext.bev_pool_forward(torch.from_numpy(feats).cuda(), *args, 1, 1, 8, 8)
    ref = oracle_mod.bev_pool_forward(feats, geom, lengths, starts, 1, 1, 8, 8)
    l1 = oracle_mod.bev_pool_forward(np.abs(feats), geom, lengths, starts, 1, 1, 8, 8)
    assert (np.abs(out.cpu().numpy() - ref) <= RTOL * np.abs(ref) + 1e-6 * l1).all()
    og = rng.standard_normal(ref.shape).astype(np.float32)
    xg = bev_pool_ext.bev_pool_backward(torch.from_numpy(og).cuda(), *args, 1, 1, 8, 8)
    np.testing.assert_array_equal(xg.cpu().numpy(), oracle_mod.bev_pool_backward(og, geom, lengths, starts, 1, 1, 8, 8))


def test_ext_validates_inputs():
    x = torch.zeros(8, 80, device="cuda")
    g = torch.zeros(8, 4, dtype=torch.int32, device="cuda")
    i = torch.zeros(1, dtype=torch.int32, device="cuda")
    with pytest.raises(RuntimeError, match="int32"):
        bev_pool_ext.bev_pool_forward(x, g.long(), i, i, 1, 1, 2, 2)
    with pytest.raises(RuntimeError, match="float32"):
        bev_pool_ext.bev_pool_forward(x.double(), g, i, i, 1, 1, 2, 2)
    with pytest.raises(RuntimeError, match="contiguous"):
        bev_pool_ext.bev_pool_forward(x.t().contiguous().t(), g, i, i, 1, 1, 2, 2)
    empty = bev_pool_ext.bev_pool_forward(x[:0], g[:0], i[:0], i[:0], 1, 1, 2, 2)
    assert empty.shape == (1, 1, 2, 2, 80) and float(empty.abs().sum()) == 0.0
    with pytest.raises(NotImplementedError):
        QuickCumsumCuda.backward(None, None)


def test_quick_cumsum_golden_on_gpu():
    g = golden("quick_cumsum.npz")
    pooled, pg = QuickCumsum.apply(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["geom"]).cuda(),
                                   torch.from_numpy(g["ranks"]).cuda())
    np.testing.assert_allclose(pooled.cpu().numpy(), g["pooled"], rtol=1e-10, atol=1e-10)
    np.testing.assert_array_equal(pg.cpu().numpy(), g["pooled_geom"])
    out = QuickCumsumTrainingCuda.apply(torch.from_numpy(g["x"]).float().cuda(), torch.from_numpy(g["geom"]).cuda(),
                                        torch.from_numpy(g["ranks"]).cuda(), 1, 1, 5, 8)
    dense = np.zeros((1, 1, 5, 8, g["x"].shape[1]))
    dense[0, 0, g["pooled_geom"][:, 0], g["pooled_geom"][:, 1]] = g["pooled"]
    np.testing.assert_allclose(out.cpu().numpy(), dense, rtol=1e-5, atol=1e-5)


def _view_case(batch, n_cams, feature_size, image_size, D_bound, C, xb, seed, zb=(-10.0, 10.0, 20.0)):
    vt = BaseViewTransform(8, C, image_size, feature_size, xb, xb, list(zb), D_bound).cuda()
    rig = {k: torch.from_numpy(v).cuda() for k, v in synthetic.camera_rig(n_cams=n_cams, image_size=image_size,
                                                                         batch=batch).items()}
    geom = vt.get_geometry(**rig)
    B, N, D, fH, fW, _ = geom.shape
    depth, ctx = synthetic.camera_features(n_cams=N, D=D, C=C, feature_size=feature_size, batch=B, seed=seed)
    return vt, geom, depth, ctx


@pytest.mark.parametrize("case", ["small_b2", "small_nz2_b2", "config_A", "config_C_custom", "stress_D236_b2"])
def test_fused_forward_matches_reference_chain(oracle_mod, case):
    if case == "small_b2":
        vt, geom, depth, ctx = _view_case(2, 3, (8, 22), (64, 176), [1.0, 30.0, 1.0], 16, [-27.0, 27.0, 0.6], 1)
    elif case == "small_nz2_b2":     # two z slabs: the collapsed channel index must be z*C + ch (depth_lss.py:202)
        vt, geom, depth, ctx = _view_case(2, 3, (8, 22), (64, 176), [1.0, 30.0, 1.0], 16, [-27.0, 27.0, 0.6], 5,
                                          zb=(-4.0, 6.0, 5.0))
    elif case == "config_C_custom":  # BASELINE configs[3]: 5 cams, 384x704 image, 48x88 features
        vt, geom, depth, ctx = _view_case(1, 5, (48, 88), (384, 704), [1.0, 60.0, 0.5], 80, [-54.0, 54.0, 0.3], 6)
    elif case == "stress_D236_b2":   # BASELINE configs[4]: 236 depth bins (dbound step 0.25), 2 frames per GPU
        vt, geom, depth, ctx = _view_case(2, 6, (32, 88), (256, 704), [1.0, 60.0, 0.25], 80, [-54.0, 54.0, 0.3], 7)
    else:  # BASELINE configs[0]/[1]: 6 cams x 118 x 32 x 88 x 80 ch -> 360 x 360
        vt, geom, depth, ctx = _view_case(1, 6, (32, 88), (256, 704), [1.0, 60.0, 0.5], 80, [-54.0, 54.0, 0.3], 0)
    B, N, D, fH, fW, _ = geom.shape
    C = ctx.shape[1]
    d_t, c_t = torch.from_numpy(depth).cuda(), torch.from_numpy(ctx).cuda()
    tabs = vt.build_tables(geom)
    out = vt.pool_fused(d_t, c_t)
    # both forward kernels (ray-major runs, cell-major intervals) must agree; `out` above used the default choice
    assert tabs.n_runs > 0
    tabs.use_runs = not tabs.use_runs
    out_other = vt.pool_fused(d_t, c_t)
    tabs.use_runs = not tabs.use_runs
    np.testing.assert_allclose(out_other.cpu().numpy(), out.cpu().numpy(), rtol=RTOL,
                               atol=RTOL * float(out.abs().max()))
    # (1) against the boundary form fed with the materialised frustum tensor (the reference's own data path)
    x = (d_t.unsqueeze(1) * c_t.unsqueeze(2)).view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    vt.eval()
    ref_gpu = vt.bev_pool(x, geom)
    assert out.shape == ref_gpu.shape == (B, C * int(vt.nx[2]), int(vt.nx[0]), int(vt.nx[1]))
    scale = float(ref_gpu.abs().max())
    np.testing.assert_allclose(out.cpu().numpy(), ref_gpu.cpu().numpy(), rtol=RTOL, atol=RTOL * scale)
    # (2) against the CPU oracle of the whole chain
    gf, kept, ranks, indices = vt.bev_pool_aux(geom)
    src = torch.nonzero(kept, as_tuple=False).squeeze(1)[indices].cpu().numpy()
    starts, lengths = oracle_mod.intervals_from_ranks(ranks.cpu().numpy())
    want = oracle_mod.bev_pool_fused(depth, ctx, src, gf.int().cpu().numpy(), starts, lengths, B, int(vt.nx[2]),
                                     int(vt.nx[0]), int(vt.nx[1]))
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=RTOL, atol=RTOL * scale)
    if case == "config_A":
        assert tabs.nk > 1_700_000 and 40_000 < tabs.n_intervals < 50_000
        assert tabs.use_runs and tabs.n_runs < 80_000   # ~N*D*fW runs of ~fH points


@pytest.mark.parametrize("zb", [(-10.0, 10.0, 20.0), (-4.0, 6.0, 5.0)])
def test_fused_backward_matches_autograd_of_reference_chain(zb):
    vt, geom, depth, ctx = _view_case(2, 3, (8, 22), (64, 176), [1.0, 30.0, 1.0], 16, [-27.0, 27.0, 0.6], 2, zb=zb)
    assert int(vt.nx[2]) == (1 if zb[2] == 20.0 else 2)
    B, N, D, fH, fW, _ = geom.shape
    C = ctx.shape[1]
    vt.build_tables(geom)
    d1 = torch.from_numpy(depth).cuda().requires_grad_(True)
    c1 = torch.from_numpy(ctx).cuda().requires_grad_(True)
    out = vt.pool_fused(d1, c1)
    g = torch.randn_like(out)
    out.backward(g)
    # reference chain with torch autograd through the boundary op (QuickCumsumTrainingCuda backward = K2)
    d2 = torch.from_numpy(depth).cuda().requires_grad_(True)
    c2 = torch.from_numpy(ctx).cuda().requires_grad_(True)
    x = (d2.unsqueeze(1) * c2.unsqueeze(2)).view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    vt.train()
    ref = vt.bev_pool(x, geom)
    ref.backward(g)
    for a, b in ((d1.grad, d2.grad), (c1.grad, c2.grad)):
        scale = float(b.abs().max())
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=1e-5 * scale)


def test_linearity_and_zero_cells_full_size():
    """Size-independent properties at BASELINE's full size: pool(a*d, c) == a*pool(d, c); cells no frustum point
    reaches are exactly zero; the sum over the BEV grid equals the sum of depth*ctx over kept points."""
    vt, geom, depth, ctx = _view_case(1, 6, (32, 88), (256, 704), [1.0, 60.0, 0.5], 80, [-54.0, 54.0, 0.3], 3)
    d_t, c_t = torch.from_numpy(depth).cuda(), torch.from_numpy(ctx).cuda()
    tabs = vt.build_tables(geom)
    out = vt.pool_fused(d_t, c_t)
    out2 = vt.pool_fused(d_t * 2.0, c_t)
    np.testing.assert_allclose(out2.cpu().numpy(), 2.0 * out.cpu().numpy(), rtol=1e-6, atol=1e-6)
    occ = torch.zeros(360 * 360, dtype=torch.bool, device="cuda")
    occ[tabs.interval_cell.long()] = True
    assert float(out.view(80, -1)[:, ~occ].abs().max()) == 0.0
    kept_w = torch.zeros(depth.size, device="cuda", dtype=torch.float64)
    kept_w[tabs.src.long()] = 1.0
    total = ((d_t.double() * kept_w.view_as(d_t)).sum(1, keepdim=True) * c_t.double()).sum()
    np.testing.assert_allclose(float(out.double().sum()), float(total), rtol=1e-6)


TABLE_FIELDS = ("src", "interval_starts", "interval_cell", "cell_of_point", "tile_starts", "run_p0", "run_len", "run_pos",
                "col_run_starts", "cell_run_ids", "cell_run_starts")


def _assert_tables_equal(dev_t, ref_t):
    assert (dev_t.nk, dev_t.n_intervals, dev_t.n_runs) == (ref_t.nk, ref_t.n_intervals, ref_t.n_runs)
    assert dev_t.use_runs == ref_t.use_runs
    for f in TABLE_FIELDS:
        a, b = getattr(dev_t, f), getattr(ref_t, f)
        assert a.dtype == torch.int32 and a.shape == b.shape, f
        assert torch.equal(a, b), f"table {f} differs"


@pytest.mark.parametrize("case", ["small_b2", "config_A", "config_C_custom", "stress_D236_b2", "nz2_b2"])
def test_device_built_tables_equal_bev_pool_aux_route(case):
    """csrc/bev_tables.cu against the reference's bev_pool_aux formulation (depth_lss.py:118-176 + the interval
    construction of bev_pool.py:158-166): every table bit for bit (integer work)."""
    if case == "small_b2":
        vt, geom, _, _ = _view_case(2, 3, (8, 22), (64, 176), [1.0, 30.0, 1.0], 16, [-27.0, 27.0, 0.6], 1)
    elif case == "config_C_custom":
        vt, geom, _, _ = _view_case(1, 5, (48, 88), (384, 704), [1.0, 60.0, 0.5], 80, [-54.0, 54.0, 0.3], 6)
    elif case == "stress_D236_b2":
        vt, geom, _, _ = _view_case(2, 6, (32, 88), (256, 704), [1.0, 60.0, 0.25], 80, [-54.0, 54.0, 0.3], 7)
    elif case == "nz2_b2":   # two Z slabs and two samples: rank order (x, y, z, b) != output-memory order (b, z, x, y)
        vt = BaseViewTransform(8, 16, (64, 176), (8, 22), [-27.0, 27.0, 0.6], [-27.0, 27.0, 0.6], [-4.0, 4.0, 4.0],
                               [1.0, 30.0, 1.0]).cuda()
        rig = {k: torch.from_numpy(v).cuda() for k, v in synthetic.camera_rig(n_cams=3, image_size=(64, 176),
                                                                             batch=2).items()}
        geom = vt.get_geometry(**rig)
    else:
        vt, geom, _, _ = _view_case(1, 6, (32, 88), (256, 704), [1.0, 60.0, 0.5], 80, [-54.0, 54.0, 0.3], 0)
    ref_t = vt.build_tables(geom, device_build=False)
    dev_t = vt.build_tables(geom, device_build=True)
    assert dev_t.nk > 0 and dev_t.n_runs > 0
    _assert_tables_equal(dev_t, ref_t)
    dev_again = vt.build_tables(geom)   # default for CUDA geometry; no atomics-order dependence
    _assert_tables_equal(dev_again, ref_t)


def test_device_built_tables_on_reference_golden_geometry(oracle_mod):
    """The geometry the reference's own get_geometry produced (with extra_rots / extra_trans), its kept mask and
    rank order: src must be the reference's kept[indices] up to the order inside one cell."""
    g = golden("view_geometry.npz")
    geom = torch.from_numpy(g["geom"]).cuda()
    B = geom.shape[0]
    nx = [int(v) for v in g["nx"]]
    t = ops.BevPoolTables.from_geometry(geom, B, g["bx"].tolist(), g["dx"].tolist(), nx)
    kept = g["kept"].reshape(-1)
    np.testing.assert_array_equal((t.cell_of_point >= 0).cpu().numpy(), kept)
    assert t.nk == int(kept.sum()) == g["geom_feats"].shape[0]
    gf = g["geom_feats"]                                     # sorted (x, y, z, b) rows, rank order
    cell_ref = (gf[:, 3] * nx[2] + gf[:, 2]) * (nx[0] * nx[1]) + gf[:, 0] * nx[1] + gf[:, 1]
    cop = t.cell_of_point.cpu().numpy()
    np.testing.assert_array_equal(np.sort(cop[cop >= 0]), np.sort(cell_ref))
    src = t.src.cpu().numpy()
    assert (np.diff(cop[src]) >= 0).all()                    # ordered by cell
    same = cop[src][1:] == cop[src][:-1]
    assert (np.diff(src)[same] > 0).all()                    # ties in frustum order
    np.testing.assert_array_equal(np.unique(cell_ref), t.interval_cell.cpu().numpy())


def test_device_built_tables_edge_values():
    """Truncation toward zero keeps (-1, 0) in cell 0 (depth_lss.py:129 `.long()`); NaN / far-away points drop; an
    all-outside frustum gives empty tables."""
    vt = BaseViewTransform(8, 16, (64, 176), (8, 22), [-27.0, 27.0, 0.6], [-27.0, 27.0, 0.6], [-10.0, 10.0, 20.0],
                           [1.0, 4.0, 1.0]).cuda()
    geom = torch.full((1, 1, 3, 8, 22, 3), 1e9, device="cuda")
    lo = float(vt.bx[0] - vt.dx[0] / 2)
    geom[0, 0, 0, 0, 0] = torch.tensor([lo - 0.3, lo - 0.59, 0.0])   # u in (-1, 0) on x and y -> cell (0, 0)
    geom[0, 0, 0, 1, 0] = torch.tensor([lo - 0.6001, lo, 0.0])       # u <= -1 -> dropped
    geom[0, 0, 0, 2, 0] = torch.tensor([float("nan"), lo, 0.0])
    geom[0, 0, 1, 0, 5] = torch.tensor([lo + 53.99, lo + 53.99, 9.9])  # last cell
    ref_t = vt.build_tables(geom, device_build=False)
    dev_t = vt.build_tables(geom, device_build=True)
    _assert_tables_equal(dev_t, ref_t)
    assert dev_t.nk == 2 and dev_t.interval_cell.tolist() == [0, 90 * 90 - 1]
    empty = vt.build_tables(torch.full((1, 1, 3, 8, 22, 3), 1e9, device="cuda"), device_build=True)
    assert (empty.nk, empty.n_intervals, empty.n_runs) == (0, 0, 0) and not empty.use_runs
    out = vt.pool_fused(torch.rand(1, 3, 8, 22, device="cuda"), torch.rand(1, 16, 8, 22, device="cuda"), empty)
    assert out.shape == (1, 16, 90, 90) and float(out.abs().sum()) == 0.0
