"""The library against the REFERENCE'S OWN CUDA OPS run on the same B200 (oracle/_ref/ref_voxel_cuda, ref_bev_pool_cuda:
the unmodified sources of projects/BEVFusion/bevfusion/ops/{voxel,bev_pool}/src compiled for sm_100a by
oracle/build_ref.py).  This pins what no CPU oracle can: the GPU kernel's partial -1 rows of dynamic_voxelize
(voxelization_cuda.cu:25-61), the set semantics of the non-deterministic voxelizer (:375-483), dynamic scatter
(scatter_points_cuda.cu:183-308, which has no CPU path and no reference test) and K1 / K2 of bev_pool
(bev_pool_cuda.cu:20-42, 61-84).

Bar: indices / coordinates / counts / voxels bit-exact; fp32 sums 1e-5 relative (+1e-6 * sum|terms| for summation order).
"""
import numpy as np
import pytest
import torch

from bevfusion_3d_object_detection_b200 import ops, synthetic
from bevfusion_3d_object_detection_b200.ops.bev_pool import bev_pool_ext
from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer
from bevfusion_3d_object_detection_b200.view_transform import BaseViewTransform

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_voxel():
    from oracle import build_ref

    mod = build_ref.load_ref("ref_voxel_cuda")
    if mod is None:
        pytest.skip("oracle/_ref/ref_voxel_cuda not built (python oracle/build_ref.py where /root/reference is visible)")
    return mod


@pytest.fixture(scope="module")
def ref_pool():
    from oracle import build_ref

    mod = build_ref.load_ref("ref_bev_pool_cuda")
    if mod is None:
        pytest.skip("oracle/_ref/ref_bev_pool_cuda not built")
    return mod


def _ref_hard(mod, pts, vs, cr, mp, mv, deterministic):
    p = torch.from_numpy(pts).cuda()
    voxels = p.new_zeros((mv, mp, p.shape[1]))
    coors = p.new_zeros((mv, 3), dtype=torch.int32)
    npv = p.new_zeros((mv,), dtype=torch.int32)
    m = mod.hard_voxelize(p, voxels, coors, npv, [float(v) for v in vs], [float(v) for v in cr], mp, mv, 3,
                          deterministic)
    torch.cuda.synchronize()
    return voxels[:m].cpu().numpy(), coors[:m].cpu().numpy(), npv[:m].cpu().numpy()


def _ours_hard(pts, vs, cr, mp, mv, deterministic=True):
    v, c, n = ops.voxelization(torch.from_numpy(pts).cuda(), list(vs), list(cr), mp, mv, deterministic)
    torch.cuda.synchronize()
    return v.cpu().numpy(), c.cpu().numpy(), n.cpu().numpy()


@pytest.mark.parametrize("cfg", ["one_sweep", "one_sweep_cap", "three_dim", "two_sweeps_coarse"])
def test_hard_voxelize_deterministic_bit_exact_vs_reference_gpu(ref_voxel, cfg):
    """hard_voxelize_gpu (voxelization_cuda.cu:231-373: the O(N^2) scan + single-thread walk) on the B200."""
    vs, cr, mp = synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 10
    if cfg == "one_sweep":          # BASELINE configs[0]: ~34 k points, 5 dims
        pts, mv = synthetic.lidar_sweeps(n_sweeps=1, azimuth=1150), 160000
    elif cfg == "one_sweep_cap":    # the max_voxels cap bites: voxels past the cap and their points are dropped
        pts, mv = synthetic.lidar_sweeps(n_sweeps=1, seed=2), 9000
    elif cfg == "three_dim":        # configs[3]: 3-dim points
        pts, mv = synthetic.lidar_sweeps(n_sweeps=1, dims=3, seed=5), 120000
    else:                           # coarse voxels: many voxels overflow max_points
        pts, mv, vs, mp = synthetic.lidar_sweeps(n_sweeps=2, seed=7), 40000, [0.6, 0.6, 0.8], 5
    rv, rc, rn = _ref_hard(ref_voxel, pts, vs, cr, mp, mv, True)
    v, c, n = _ours_hard(pts, vs, cr, mp, mv)
    assert v.shape == rv.shape
    np.testing.assert_array_equal(c, rc)
    np.testing.assert_array_equal(n, rn)
    np.testing.assert_array_equal(v, rv)


def test_hard_voxelize_nondeterministic_same_voxel_set(ref_voxel):
    """nondisterministic_hard_voxelize_gpu (:375-483): voxel order, which points survive an overflow and which voxels
    survive the cap are race-dependent; what is defined is the voxel SET (no cap), min(count, max_points) per voxel
    and, for voxels that do not overflow, the multiset of their points.  Our deterministic answer must be one of the
    legal outcomes."""
    vs, cr, mp, mv = [0.3, 0.3, 0.4], synthetic.NUSCENES_RANGE, 10, 200000
    pts = synthetic.lidar_sweeps(n_sweeps=2, seed=9)
    rv, rc, rn = _ref_hard(ref_voxel, pts, vs, cr, mp, mv, False)
    v, c, n = _ours_hard(pts, vs, cr, mp, mv, deterministic=False)
    assert c.shape == rc.shape

    def canon(v, c, n):
        key = (c[:, 0].astype(np.int64) * 4096 + c[:, 1]) * 4096 + c[:, 2]
        o = np.argsort(key, kind="stable")
        return key[o], n[o], v[o]

    k0, n0, v0 = canon(rv, rc, rn)
    k1, n1, v1 = canon(v, c, n)
    np.testing.assert_array_equal(k1, k0)
    np.testing.assert_array_equal(n1, n0)
    small = n0 < mp   # no overflow: every point of the voxel is kept by both; compare as sorted multisets
    a = np.sort(v0[small].reshape(int(small.sum()), -1), axis=1)
    b = np.sort(v1[small].reshape(int(small.sum()), -1), axis=1)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("seed", [0, 1])
def test_dynamic_voxelize_partial_rows_vs_reference_gpu(ref_voxel, seed):
    """dynamic_voxelize_gpu leaves PARTIAL rows for out-of-range points ((-1, old, old), (-1, -1, old), ...); the
    pre-filled buffer shows which components the kernel touches."""
    rng = np.random.default_rng(seed)
    pts = synthetic.lidar_sweeps(n_sweeps=1, seed=seed)
    noise = rng.uniform(-70, 70, (5000, pts.shape[1])).astype(np.float32)   # plenty of out-of-range rows on every axis
    noise[:, 2] = rng.uniform(-8, 6, 5000)
    pts = np.concatenate([pts, noise])[rng.permutation(pts.shape[0] + 5000)]
    p = torch.from_numpy(pts).cuda()
    for fill in (0, 7):
        ref_c = torch.full((pts.shape[0], 3), fill, dtype=torch.int32, device="cuda")
        ref_voxel.dynamic_voxelize(p, ref_c, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 3)
        our_c = torch.full((pts.shape[0], 3), fill, dtype=torch.int32, device="cuda")
        voxel_layer.dynamic_voxelize(p, our_c, synthetic.NUSCENES_VOXEL, synthetic.NUSCENES_RANGE, 3)
        torch.cuda.synchronize()
        assert int((ref_c[:, 0] < 0).sum()) > 100
        assert torch.equal(our_c, ref_c)


@pytest.mark.parametrize("reduce_type", ["max", "sum", "mean"])
@pytest.mark.parametrize("ndim", [3, 4])
def test_dynamic_scatter_vs_reference_gpu(ref_voxel, reduce_type, ndim):
    rng = np.random.default_rng(10 + ndim)
    n, c = 30000, 5
    coors = rng.integers(-1, 28, (n, ndim)).astype(np.int32)
    feats = rng.standard_normal((n, c)).astype(np.float32)
    f, co = torch.from_numpy(feats).cuda(), torch.from_numpy(coors).cuda()
    r_red, r_coors, r_map, r_cnt = ref_voxel.dynamic_point_to_voxel_forward(f, co, reduce_type)
    o_red, o_coors, o_map, o_cnt = voxel_layer.dynamic_point_to_voxel_forward(f, co, reduce_type)
    torch.cuda.synchronize()
    assert torch.equal(o_coors, r_coors)            # unique_dim order: lexicographic
    assert torch.equal(o_map, r_map)
    assert torch.equal(o_cnt, r_cnt)
    if reduce_type == "max":
        assert torch.equal(o_red, r_red)
    else:   # the reference sums with float atomics in race order
        np.testing.assert_allclose(o_red.cpu().numpy(), r_red.cpu().numpy(), rtol=1e-5, atol=1e-5)
    g = torch.from_numpy(rng.standard_normal(tuple(r_red.shape)).astype(np.float32)).cuda()
    r_g = torch.zeros_like(f)
    ref_voxel.dynamic_point_to_voxel_backward(r_g, g, f, r_red, r_map, r_cnt, reduce_type)
    o_g = torch.zeros_like(f)
    voxel_layer.dynamic_point_to_voxel_backward(o_g, g, f, r_red, r_map, r_cnt, reduce_type)
    torch.cuda.synchronize()
    np.testing.assert_allclose(o_g.cpu().numpy(), r_g.cpu().numpy(), rtol=1e-6, atol=1e-7)


def _pool_inputs(c=80, seed=0):
    """config-A geometry through bev_pool_aux -> the exact arguments bev_pool() hands to bev_pool_ext."""
    vt = BaseViewTransform(8, c, (256, 704), (32, 88), [-54.0, 54.0, 0.3], [-54.0, 54.0, 0.3], [-10.0, 10.0, 20.0],
                           [1.0, 60.0, 0.5]).cuda()
    rig = {k: torch.from_numpy(v).cuda() for k, v in synthetic.camera_rig(n_cams=6, image_size=(256, 704)).items()}
    geom = vt.get_geometry(**rig)
    geom_feats, kept, ranks, indices = vt.bev_pool_aux(geom)
    nk = int(ranks.shape[0])
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((nk, c), device="cuda", generator=g)
    keep = torch.ones(nk, dtype=torch.bool, device="cuda")
    keep[1:] = ranks[1:] != ranks[:-1]
    starts = torch.where(keep)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = nk - starts[-1]
    return x, geom_feats.int().contiguous(), lengths.contiguous(), starts.contiguous()


def test_bev_pool_forward_backward_vs_reference_kernels(ref_pool):
    """K1 / K2 at BASELINE's full size (1.8 M sorted rows x 80 channels -> 360 x 360)."""
    x, geom, lengths, starts = _pool_inputs()
    ref = ref_pool.bev_pool_forward(x, geom, lengths, starts, 1, 1, 360, 360)
    out = bev_pool_ext.bev_pool_forward(x, geom, lengths, starts, 1, 1, 360, 360)
    l1 = ref_pool.bev_pool_forward(x.abs(), geom, lengths, starts, 1, 1, 360, 360)
    torch.cuda.synchronize()
    assert out.shape == ref.shape == (1, 1, 360, 360, 80)
    err = (out - ref).abs()
    assert bool((err <= 1e-5 * ref.abs() + 1e-6 * l1).all()), float(err.max())
    og = torch.randn_like(ref)
    r_g = ref_pool.bev_pool_backward(og, geom, lengths, starts, 1, 1, 360, 360)
    o_g = bev_pool_ext.bev_pool_backward(og, geom, lengths, starts, 1, 1, 360, 360)
    torch.cuda.synchronize()
    assert torch.equal(o_g, r_g)   # pure broadcast: exact
