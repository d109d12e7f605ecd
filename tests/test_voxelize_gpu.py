"""CUDA voxelization (through the C ABI) against the CPU oracle and the reference golden vectors: bit-exact."""
import numpy as np
import pytest
import torch

from conftest import golden
from bevfusion_3d_object_detection_b200 import ops, synthetic
from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer

pytestmark = pytest.mark.gpu


def _run_hard(points, vs, cr, mp, mv, deterministic=True):
    p = torch.from_numpy(points).cuda()
    v, c, n = ops.voxelization(p, list(vs), list(cr), int(mp), int(mv), deterministic)
    torch.cuda.synchronize()
    return v.cpu().numpy(), c.cpu().numpy(), n.cpu().numpy()


@pytest.mark.parametrize("name", ["voxel_numba_sweep_nocap.npz", "voxel_numba_sweep_cap.npz",
                                  "voxel_numba_sweep_coarse.npz", "voxel_ref_cpp_asis.npz", "voxel_ref_cpp_fixed.npz"])
def test_hard_voxelize_matches_reference_golden(name):
    g = golden(name)
    v, c, n = _run_hard(g["points"].astype(np.float32), g["voxel_size"], g["coors_range"], g["max_points"],
                        g["max_voxels"])
    np.testing.assert_array_equal(c, g["coors"])
    np.testing.assert_array_equal(n, g["npv"])
    np.testing.assert_array_equal(v, g["voxels"].astype(np.float32))


def test_reference_known_answer_on_gpu():
    np.random.seed(0)
    points = np.random.uniform(0, 4, (20, 3)).astype(np.float32)
    v, c, n = _run_hard(points, [5, 5, 1], [0, 0, 0, 20, 40, 4], 5, 20000)
    assert v.shape == (4, 5, 3)
    np.testing.assert_array_equal(c[:, ::-1], [[2, 0, 0], [3, 0, 0], [0, 0, 0], [1, 0, 0]])
    np.testing.assert_array_equal(n, [5, 5, 5, 3])


@pytest.mark.parametrize("cfg", ["cpu_ref_34k", "full_10sweep_cap", "three_dim_custom", "stress_1m"])
def test_hard_voxelize_bit_exact_vs_oracle(oracle_mod, cfg):
    if cfg == "cpu_ref_34k":  # BASELINE configs[0]
        pts, vs, mv = synthetic.lidar_sweeps(n_sweeps=1, azimuth=1150), synthetic.NUSCENES_VOXEL, 160000
    elif cfg == "full_10sweep_cap":  # BASELINE configs[1]: ~320 k points, hits the 160 000 voxel cap
        pts, vs, mv = synthetic.lidar_sweeps(), synthetic.NUSCENES_VOXEL, 160000
    elif cfg == "three_dim_custom":  # configs[3]: 3-dim points, training cap
        pts, vs, mv = synthetic.lidar_sweeps(n_sweeps=3, dims=3, seed=5), synthetic.NUSCENES_VOXEL, 120000
    else:  # configs[4]: ~0.9 M points at 0.05 m
        pts, vs, mv = synthetic.stress_sweep(), [0.05, 0.05, 0.2], 500000
    ov, oc, on = oracle_mod.hard_voxelize(pts, vs, synthetic.NUSCENES_RANGE, 10, mv)
    v, c, n = _run_hard(pts, vs, synthetic.NUSCENES_RANGE, 10, mv)
    assert v.shape == ov.shape
    np.testing.assert_array_equal(c, oc)
    np.testing.assert_array_equal(n, on)
    np.testing.assert_array_equal(v, ov)
    if cfg == "full_10sweep_cap":
        assert v.shape[0] == 160000


def test_edge_cases(oracle_mod):
    vs, cr = [1.0, 1.0, 1.0], [0, 0, 0, 4, 4, 4]
    # empty input
    v, c, n = _run_hard(np.zeros((0, 4), np.float32), vs, cr, 5, 100)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3) and n.shape == (0,)
    # every point out of range
    v, c, n = _run_hard(np.full((50, 4), 9.0, np.float32), vs, cr, 5, 100)
    assert v.shape[0] == 0
    # all points in ONE voxel (the reference's O(N^2) worst case), more points than max_points
    pts = np.random.default_rng(0).uniform(1.0, 1.99, (5000, 4)).astype(np.float32)
    v, c, n = _run_hard(pts, vs, cr, 7, 100)
    assert v.shape[0] == 1 and n[0] == 7
    np.testing.assert_array_equal(v[0], pts[:7])
    # points exactly on cell and range boundaries, negative zero, max_points = 1, max_voxels = 1
    pts = np.array([[0, 0, 0, 1], [1, 1, 1, 2], [4, 0, 0, 3], [3.9999998, 3.9999998, 3.9999998, 4],
                    [-0.0, 0.5, 0.5, 5], [2, 2, 2, 6], [-1e-9, 0, 0, 7]], np.float32)
    for mp, mv in ((1, 100), (5, 1), (5, 3)):
        ov, oc, on = oracle_mod.hard_voxelize(pts, vs, cr, mp, mv)
        v, c, n = _run_hard(pts, vs, cr, mp, mv)
        np.testing.assert_array_equal(c, oc)
        np.testing.assert_array_equal(n, on)
        np.testing.assert_array_equal(v, ov)


def test_non_deterministic_flag_and_module(oracle_mod):
    pts = synthetic.lidar_sweeps(n_sweeps=2, seed=9)
    ov, oc, on = oracle_mod.hard_voxelize(pts, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10, 30000)
    v, c, n = _run_hard(pts, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10, 30000, deterministic=False)
    np.testing.assert_array_equal(c, oc)
    np.testing.assert_array_equal(v, ov)
    mod = ops.Voxelization(synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10, (120000, 30000)).cuda().eval()
    mv, mc, mn = mod(torch.from_numpy(pts).cuda())
    assert mc.dtype == torch.int32 and mn.dtype == torch.int32
    np.testing.assert_array_equal(mc.cpu().numpy(), oc)
    mod.train()
    tv, tc, tn = mod(torch.from_numpy(pts).cuda())
    assert tv.shape[0] == min(120000, oracle_mod.hard_voxelize(pts, synthetic.NUSCENES_VOXEL,
                                                               synthetic.NUSCENES_RANGE, 10, 120000)[0].shape[0])


def test_async_zero_fill_variant(oracle_mod):
    pts = synthetic.lidar_sweeps(n_sweeps=2, seed=2)
    mp, mv = 10, 40000
    p = torch.from_numpy(pts).cuda()
    voxels = torch.full((mv, mp, 5), float("nan"), device="cuda")
    coors = torch.full((mv, 3), -7, dtype=torch.int32, device="cuda")
    npv = torch.full((mv,), -7, dtype=torch.int32, device="cuda")
    m = voxel_layer.hard_voxelize_async(p, voxels, coors, npv, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, mp, mv,
                                        zero_fill=True)
    m = int(m.item())
    ov, oc, on = oracle_mod.hard_voxelize(pts, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, mp, mv)
    assert m == ov.shape[0]
    np.testing.assert_array_equal(voxels[:m].cpu().numpy(), ov)
    np.testing.assert_array_equal(coors[:m].cpu().numpy(), oc)
    np.testing.assert_array_equal(npv[:m].cpu().numpy(), on)


def test_voxelize_mean_batch(oracle_mod):
    """Extension: fused voxelize + mean + batch pad == BEVFusion.voxelize (bevfusion.py:227-255)."""
    clouds = [synthetic.lidar_sweeps(n_sweeps=2, seed=s) for s in (1, 2, 3)]
    mod = ops.Voxelization(synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10, (120000, 25000)).cuda().eval()
    feats, coords, sizes = mod.forward_mean([torch.from_numpy(c).cuda() for c in clouds])
    ref_f, ref_c, ref_s = [], [], []
    for k, pts in enumerate(clouds):
        ov, oc, on = oracle_mod.hard_voxelize(pts, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10, 25000)
        ref_f.append(oracle_mod.voxel_mean(ov, on))
        ref_c.append(np.concatenate([np.full((oc.shape[0], 1), k, np.int32), oc], 1))
        ref_s.append(on)
    np.testing.assert_array_equal(coords.cpu().numpy(), np.concatenate(ref_c))
    np.testing.assert_array_equal(sizes.cpu().numpy(), np.concatenate(ref_s))
    np.testing.assert_allclose(feats.cpu().numpy(), np.concatenate(ref_f), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["voxel_ref_cpp_asis.npz", "voxel_ref_cpp_fixed.npz"])
def test_dynamic_voxelize(oracle_mod, name):
    g = golden(name)
    pts = g["points"].astype(np.float32)
    p = torch.from_numpy(pts).cuda()
    coors = ops.voxelization(p, list(g["voxel_size"]), list(g["coors_range"]), -1, -1)
    got = coors.cpu().numpy()
    want = oracle_mod.dynamic_voxelize(pts, g["voxel_size"], g["coors_range"], gpu_partial=True)
    np.testing.assert_array_equal(got, want)  # includes the reference kernel's partial -1 rows
    ok = g["dyn_coors_cpu"][:, 0] >= 0
    np.testing.assert_array_equal(got[ok], g["dyn_coors_cpu"][ok])  # reference C++ CPU op on in-range rows
    # idempotent on a pre-filled buffer: untouched components keep the caller's value
    pre = torch.full((pts.shape[0], 3), 5, dtype=torch.int32, device="cuda")
    voxel_layer.dynamic_voxelize(p, pre, list(g["voxel_size"]), list(g["coors_range"]), 3)
    want5 = oracle_mod.dynamic_voxelize(pts, g["voxel_size"], g["coors_range"], gpu_partial=True,
                                        coors_init=np.full((pts.shape[0], 3), 5, np.int32))
    np.testing.assert_array_equal(pre.cpu().numpy(), want5)


@pytest.mark.parametrize("reduce_type", ["max", "sum", "mean"])
@pytest.mark.parametrize("ndim", [3, 4])
def test_dynamic_scatter_forward_backward(oracle_mod, reduce_type, ndim):
    rng = np.random.default_rng(3)
    n, c = 20000, 5
    coors = rng.integers(-1, 24, (n, ndim)).astype(np.int32)
    feats = rng.standard_normal((n, c)).astype(np.float32)
    f = torch.from_numpy(feats).cuda().requires_grad_(True)
    vf, vc = ops.dynamic_scatter(f, torch.from_numpy(coors).cuda(), reduce_type)
    r, oc, cmap, rc = oracle_mod.dynamic_scatter(feats, coors, reduce_type)
    np.testing.assert_array_equal(vc.cpu().numpy(), oc)
    if reduce_type == "max":
        np.testing.assert_array_equal(vf.detach().cpu().numpy(), r)
    else:
        np.testing.assert_allclose(vf.detach().cpu().numpy(), r, rtol=1e-5, atol=1e-5)
    g = rng.standard_normal(r.shape).astype(np.float32)
    vf.backward(torch.from_numpy(g).cuda())
    want = oracle_mod.dynamic_scatter_backward(g, feats, r if reduce_type != "max" else vf.detach().cpu().numpy(),
                                               cmap, rc, reduce_type)
    np.testing.assert_allclose(f.grad.cpu().numpy(), want, rtol=1e-6, atol=1e-7)


def test_dynamic_scatter_module_and_edges(oracle_mod):
    ds = ops.DynamicScatter([0.1, 0.1, 0.1], [0, 0, 0, 1, 1, 1], True)
    assert "average_points=True" in repr(ds)
    rng = np.random.default_rng(4)
    coors = np.concatenate([rng.integers(0, 3, (3000, 1)), rng.integers(0, 12, (3000, 3))], 1).astype(np.int32)
    coors = coors[np.argsort(coors[:, 0], kind="stable")]
    feats = rng.standard_normal((3000, 4)).astype(np.float32)
    vf, vc = ds(torch.from_numpy(feats).cuda(), torch.from_numpy(coors).cuda())
    r, oc, _, _ = oracle_mod.dynamic_scatter(feats, coors, "mean")
    np.testing.assert_array_equal(vc.cpu().numpy(), oc)  # batch-major lexicographic == per-sample concat
    np.testing.assert_allclose(vf.cpu().numpy(), r, rtol=1e-5, atol=1e-5)
    # empty input, and all rows invalid
    e = ops.dynamic_scatter(torch.zeros(0, 4, device="cuda"), torch.zeros(0, 3, dtype=torch.int32, device="cuda"), "max")
    assert e[0].shape == (0, 4)
    vf, vc = ops.dynamic_scatter(torch.ones(10, 4, device="cuda"), torch.full((10, 3), -1, dtype=torch.int32, device="cuda"),
                                 "sum")
    assert vf.shape[0] == 0 and vc.shape[0] == 0
    with pytest.raises(RuntimeError, match="reduce type"):
        ops.dynamic_scatter(torch.ones(10, 4, device="cuda"), torch.zeros(10, 3, dtype=torch.int32, device="cuda"), "min")


def test_mmcv_style_zyx_spelling(oracle_mod):
    """mmdet3d's data-preprocessor spelling (coords zyx, count through a tensor) == the project spelling flipped;
    the reference's known answer (test_voxel_generator.py:7-20) is stated in zyx."""
    from bevfusion_3d_object_detection_b200.ops.voxel import mmcv_style

    np.random.seed(0)
    points = np.random.uniform(0, 4, (20, 3)).astype(np.float32)
    v, c, n = mmcv_style.voxelization(torch.from_numpy(points).cuda(), [5, 5, 1], [0, 0, 0, 20, 40, 4], 5, 20000, True)
    np.testing.assert_array_equal(c.cpu().numpy(), [[2, 0, 0], [3, 0, 0], [0, 0, 0], [1, 0, 0]])
    np.testing.assert_array_equal(n.cpu().numpy(), [5, 5, 5, 3])
    pts = synthetic.lidar_sweeps(n_sweeps=1, seed=9)
    layer = mmcv_style.VoxelizationByGridShape(synthetic.NUSCENES_RANGE, 10, voxel_size=synthetic.NUSCENES_VOXEL,
                                               max_voxels=(120000, 160000)).eval()
    assert layer.grid_shape == [1440, 1440, 40]
    v, c, n = layer(torch.from_numpy(pts).cuda())
    ov, oc, on = oracle_mod.hard_voxelize(pts, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10, 160000)
    np.testing.assert_array_equal(c.cpu().numpy(), oc[:, ::-1])
    np.testing.assert_array_equal(n.cpu().numpy(), on)
    np.testing.assert_array_equal(v.cpu().numpy(), ov)
    # dynamic form: CPU contract of mmcv (a failed point is (-1, -1, -1)), zyx
    dyn = mmcv_style.voxelization(torch.from_numpy(pts).cuda(), synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, -1, -1)
    want = oracle_mod.dynamic_voxelize(pts, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, gpu_partial=False)
    np.testing.assert_array_equal(dyn.cpu().numpy(), want[:, ::-1])
    # scatter with return_map
    feats = torch.from_numpy(pts).cuda()
    vf, vc, pmap = mmcv_style.dynamic_scatter_3d(feats, dyn.contiguous(), "mean", True)
    assert pmap.shape[0] == pts.shape[0] and vf.shape[0] == vc.shape[0] == int(pmap.max()) + 1
    scat = mmcv_style.DynamicScatter3D(synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, True)
    vf2, vc2 = scat(feats, dyn.contiguous())
    # float atomics: the sum order (and so the last bits of a mean) differs run to run
    assert torch.equal(vc2, vc) and torch.allclose(vf2, vf, rtol=1e-5, atol=1e-5)
