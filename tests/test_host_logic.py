"""Host-side mirror of the reference interface: attributes, geometry, table construction, error behaviour.
CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden
from bevfusion_3d_object_detection_b200 import ops, synthetic
from bevfusion_3d_object_detection_b200.view_transform import BaseViewTransform, gen_dx_bx


def test_voxelization_module_surface():
    v = ops.Voxelization([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, (120000, 160000))
    assert v.max_voxels == (120000, 160000) and v.deterministic is True
    assert v.grid_size.tolist() == [1440, 1440, 40]
    assert [int(x) for x in v.pcd_shape] == [1440, 1440, 1]
    assert "max_voxels=(120000, 160000)" in repr(v) and repr(v).startswith("Voxelization(")
    assert ops.Voxelization([1, 1, 1], [0, 0, 0, 4, 4, 4], 5, 300).max_voxels == (300, 300)
    assert len(list(v.parameters())) == 0


def test_cpu_tensors_fail_loudly():
    v = ops.Voxelization([1, 1, 1], [0, 0, 0, 4, 4, 4], 5, 300)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        v(torch.zeros(10, 4))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.dynamic_scatter(torch.zeros(10, 4), torch.zeros(10, 3, dtype=torch.int32), "max")
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.bev_pool(torch.zeros(4, 8), torch.zeros(4, 4, dtype=torch.int32), torch.zeros(4, dtype=torch.long), 1, 1,
                     2, 2, False)
    with pytest.raises(AssertionError):
        ops.bev_pool(torch.zeros(4, 8), torch.zeros(3, 4, dtype=torch.int32), torch.zeros(4), 1, 1, 2, 2, True)


def test_gen_dx_bx():
    dx, bx, nx = gen_dx_bx([-54.0, 54.0, 0.3], [-54.0, 54.0, 0.3], [-10.0, 10.0, 20.0])
    assert nx.tolist() == [360, 360, 1]
    np.testing.assert_allclose(dx.numpy(), [0.3, 0.3, 20.0])
    np.testing.assert_allclose(bx.numpy(), [-53.85, -53.85, 0.0], rtol=1e-6)


def _small_vt():
    return BaseViewTransform(8, 6, (64, 96), (4, 6), [-12.0, 12.0, 0.75], [-12.0, 12.0, 0.75], [-10.0, 10.0, 20.0],
                             [1.0, 15.0, 1.0])


def test_geometry_and_aux_match_reference_golden():
    g = golden("view_geometry.npz")
    vt = _small_vt()
    np.testing.assert_array_equal(vt.frustum.detach().numpy(), g["frustum"])
    rig = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("rig_")}
    geom = vt.get_geometry(**rig, extra_rots=torch.from_numpy(g["extra_rots"]),
                           extra_trans=torch.from_numpy(g["extra_trans"]))
    np.testing.assert_array_equal(geom.numpy(), g["geom"])  # same torch ops on the same CPU: bit-equal
    gf, kept, ranks, indices = vt.bev_pool_aux(geom)
    np.testing.assert_array_equal(kept.numpy(), g["kept"])
    np.testing.assert_array_equal(ranks.numpy(), g["ranks"])
    np.testing.assert_array_equal(gf.numpy(), g["geom_feats"])


def test_fused_tables_describe_the_same_pooling(oracle_mod):
    """BevPoolTables (src / CSR starts / cells / cell_of_point) reproduce, through the fused oracle, what the
    reference chain (outer product -> x[kept] -> x[indices] -> bev_pool -> collapse Z) computes."""
    vt = _small_vt()
    rig = {k: torch.from_numpy(v) for k, v in synthetic.camera_rig(n_cams=3, image_size=(64, 96), batch=2,
                                                                  src_size=(200, 300), resize=0.4).items()}
    geom = vt.get_geometry(**rig)
    B, N, D, fH, fW, _ = geom.shape
    C = 8
    depth, ctx = synthetic.camera_features(n_cams=N, D=D, C=C, feature_size=(fH, fW), batch=B, seed=4)
    gf, kept, ranks, indices = vt.bev_pool_aux(geom)
    # reference chain on CPU (oracle restatement of K1)
    x = torch.from_numpy(depth).unsqueeze(1) * torch.from_numpy(ctx).unsqueeze(2)       # [BN, C, D, fH, fW]
    x = x.view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2).reshape(-1, C)[kept][indices]
    ref = oracle_mod.bev_pool(x.numpy(), gf.numpy(), ranks.numpy(), B, int(vt.nx[2]), int(vt.nx[0]), int(vt.nx[1]))
    ref = np.concatenate([ref[:, :, z] for z in range(ref.shape[2])], 1)
    tabs = ops.BevPoolTables(gf, kept, ranks, indices, B, int(vt.nx[2]), int(vt.nx[0]), int(vt.nx[1]))
    assert tabs.interval_starts[-1].item() == tabs.nk == int(kept.sum())
    assert (tabs.interval_cell[1:] > tabs.interval_cell[:-1]).all()
    starts = tabs.interval_starts[:-1].numpy()
    lengths = np.diff(tabs.interval_starts.numpy())
    cell = tabs.interval_cell.long()
    nz, nx, ny = tabs.nz, tabs.nx, tabs.ny
    geom4 = torch.stack([(cell // ny) % nx, cell % ny, (cell // (nx * ny)) % nz, cell // (nz * nx * ny)], 1)
    geom4_rows = np.zeros((tabs.nk, 4), np.int32)
    geom4_rows[starts] = geom4.numpy()
    got = oracle_mod.bev_pool_fused(depth, ctx, tabs.src.numpy(), geom4_rows, starts, lengths, B, nz, nx, ny)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    cop = tabs.cell_of_point.numpy()
    assert (cop >= 0).sum() == tabs.nk and (cop[~kept.numpy()] == -1).all()


def test_run_tables_describe_the_same_pooling(oracle_mod):
    """The ray-major run tables (run_p0 / run_len / cell CSR) partition exactly the kept points and, summed run by
    run and then cell by cell (what the two-phase kernel does), give the reference chain's BEV map."""
    vt = _small_vt()
    rig = {k: torch.from_numpy(v) for k, v in synthetic.camera_rig(n_cams=3, image_size=(64, 96), batch=2,
                                                                  src_size=(200, 300), resize=0.4).items()}
    geom = vt.get_geometry(**rig)
    B, N, D, fH, fW, _ = geom.shape
    C = 8
    depth, ctx = synthetic.camera_features(n_cams=N, D=D, C=C, feature_size=(fH, fW), batch=B, seed=5)
    tabs = vt.build_tables(geom)
    assert tabs.n_runs > 0 and int(tabs.run_len.sum()) == tabs.nk
    assert int(tabs.cell_run_starts[-1]) == tabs.n_runs and tabs.cell_run_starts.numel() == tabs.n_intervals + 1
    plane = fH * fW
    p0, ln = tabs.run_p0.numpy().astype(np.int64), tabs.run_len.numpy()
    # every run stays inside one (camera, depth bin, column): rows h0 .. h0+len-1
    h0 = (p0 // fW) % fH
    assert (h0 + ln <= fH).all()
    cop = tabs.cell_of_point.numpy()
    covered = np.zeros(cop.shape[0], bool)
    dflat = depth.reshape(-1)
    ctx_nhwc = np.ascontiguousarray(ctx.transpose(0, 2, 3, 1)).reshape(-1, C)
    partial = np.zeros((tabs.n_runs, C), np.float64)
    run_cell = np.zeros(tabs.n_runs, np.int64)
    for r in range(tabs.n_runs):
        pts = p0[r] + np.arange(ln[r]) * fW
        assert not covered[pts].any()
        covered[pts] = True
        cells = cop[pts]
        assert (cells == cells[0]).all() and cells[0] >= 0
        run_cell[r] = cells[0]
        pix = (pts // (D * plane)) * plane + pts % plane
        partial[r] = (dflat[pts][:, None].astype(np.float64) * ctx_nhwc[pix]).sum(0)
    assert (covered == (cop >= 0)).all()
    nz, nx, ny = tabs.nz, tabs.nx, tabs.ny
    out = np.zeros((B * nz * nx * ny, C))
    starts, ids, icell = tabs.cell_run_starts.numpy(), tabs.cell_run_ids.numpy(), tabs.interval_cell.numpy()
    for t in range(tabs.n_intervals):
        rr = ids[starts[t]:starts[t + 1]]
        assert (run_cell[rr] == icell[t]).all()
        out[icell[t]] = partial[rr].sum(0)
    out = out.reshape(B, nz, nx, ny, C).transpose(0, 4, 1, 2, 3).reshape(B, C * nz, nx, ny)
    # reference chain
    gf, kept, ranks, indices = vt.bev_pool_aux(geom)
    src = torch.nonzero(kept, as_tuple=False).squeeze(1)[indices].numpy()
    st, lg = oracle_mod.intervals_from_ranks(ranks.numpy())
    ref = oracle_mod.bev_pool_fused(depth, ctx, src, gf.numpy().astype(np.int32), st, lg, B, nz, nx, ny)
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)


def test_sparse_conv_module_surface_and_checkpoint_shim():
    """spconv-2.x constructor / parameter layout, the registry route the reference builds its layers through
    (sparse_block.py:201-217), and the checkpoint weight-layout shim of write_spconv2.py:43-104."""
    import pytest
    from bevfusion_3d_object_detection_b200 import registry, spconv

    registry.register_all()
    conv = registry.build_conv_layer(dict(type="SubMConv3d", indice_key="subm1"), 5, 16, 3, padding=1, bias=False)
    assert isinstance(conv, spconv.SubMConv3d) and conv.subm and conv.indice_key == "subm1"
    assert tuple(conv.weight.shape) == (16, 3, 3, 3, 5) and conv.bias is None        # (Cout, kD, kH, kW, Cin)
    down = registry.build_conv_layer(dict(type="SparseConv3d", indice_key="spconv3"), 64, 128, 3, stride=2,
                                     padding=(1, 1, 0), bias=False)
    assert not down.subm and down.stride == [2, 2, 2] and down.padding == [1, 1, 0]
    with pytest.raises(KeyError):
        registry.build_conv_layer(dict(type="NoSuchConv"), 1, 1, 3)
    for bad in (dict(groups=2), dict(transposed=True), dict(inverse=True)):
        with pytest.raises(NotImplementedError):
            spconv.SparseConvolution(3, 4, 4, **bad)
    with pytest.raises(NotImplementedError):
        spconv.SparseConvolution(2, 4, 4)
    # a checkpoint saved by spconv 1.x / mmcv stores (kD, kH, kW, Cin, Cout): no version in the metadata -> permuted
    w_old = torch.randn(3, 3, 3, 5, 16)
    sd = {"weight": w_old.clone()}
    conv.load_state_dict(sd)
    assert torch.equal(conv.weight.data, w_old.permute(4, 0, 1, 2, 3))
    # a version-2 checkpoint (what this module itself saves) loads unchanged
    sd2 = conv.state_dict()
    assert sd2._metadata[""]["version"] == 2
    conv2 = spconv.SubMConv3d(5, 16, 3, padding=1, bias=False)
    conv2.load_state_dict(sd2)
    assert torch.equal(conv2.weight.data, conv.weight.data)
    # wrong shape is reported, not silently accepted
    with pytest.raises(RuntimeError):
        conv.load_state_dict({"weight": torch.randn(3, 3, 3, 4, 16)})


def test_encoder_layout_and_state_dict_keys():
    """BEVFusionSparseEncoder with the nuScenes config: 21 sparse convs, the reference's module names (so its
    checkpoints load by key), strided convs carry the reference's indice keys, only conv_out keeps fp32 output."""
    from bevfusion_3d_object_detection_b200 import spconv
    from bevfusion_3d_object_detection_b200.sparse_encoder import NUSCENES_ENCODER_CFG, BEVFusionSparseEncoder

    enc = BEVFusionSparseEncoder(**NUSCENES_ENCODER_CFG)
    convs = [(n, m) for n, m in enc.named_modules() if isinstance(m, spconv.SparseConvolution)]
    assert len(convs) == 21
    keys = set(enc.state_dict().keys())
    for k in ("conv_input.0.weight", "conv_input.1.running_mean", "encoder_layers.encoder_layer1.0.conv1.weight",
              "encoder_layers.encoder_layer1.0.bn1.weight", "encoder_layers.encoder_layer1.1.conv2.weight",
              "encoder_layers.encoder_layer1.2.0.weight", "encoder_layers.encoder_layer3.2.0.weight",
              "encoder_layers.encoder_layer4.1.bn2.running_var", "conv_out.0.weight", "conv_out.1.bias"):
        assert k in keys, k
    assert not any(k.endswith(".bias") and ".0." in k and "conv" in k.split(".")[-2] for k in keys)
    by_name = dict(convs)
    assert by_name["conv_input.0"].indice_key == "subm1" and by_name["conv_input.0"].in_channels == 5
    assert [by_name[f"encoder_layers.encoder_layer{i}.2.0"].indice_key for i in (1, 2, 3)] == ["spconv1", "spconv2",
                                                                                                "spconv3"]
    assert by_name["encoder_layers.encoder_layer3.2.0"].padding == [1, 1, 0]
    out = by_name["conv_out.0"]
    assert out.kernel_size == [1, 1, 3] and out.stride == [1, 1, 2] and out.indice_key == "spconv_down2"
    assert out.need_f32 and sum(m.need_f32 for _, m in convs) == 1
    assert tuple(by_name["encoder_layers.encoder_layer4.0.conv1"].weight.shape) == (128, 3, 3, 3, 128)
    n_params = sum(p.numel() for p in enc.parameters())
    assert 2.5e6 < n_params < 3.0e6   # ~2.7 M parameters


def test_rows_result_rebuilds_dense_maps_on_the_host():
    """frontend.RowsResult (host side of HostPipeline's lossless "rows" output): header | indices | rows buffer layout of
    bevf_pack_sparse_rows and the channel-major camera columns of bevf_pack_cells scatter back into
    [B, C*Z, X, Y] / [B, nz*C, nx, ny] exactly as SparseConvTensor.dense() + permute + view (sparse_encoder.py:147-151) and
    bev_pool's output layout (depth_lss.py:199-204) define them."""
    import torch

    from bevfusion_3d_object_detection_b200.frontend import RowsResult

    rng = np.random.default_rng(0)
    B, C, Z, X, Y, cap, n = 2, 8, 2, 5, 6, 40, 23
    lin = rng.choice(B * X * Y * Z, size=n, replace=False)
    idx = np.stack([lin // (X * Y * Z), (lin // (Y * Z)) % X, (lin // Z) % Y, lin % Z], 1).astype(np.int32)
    rows = rng.standard_normal((n, C)).astype(np.float32)
    raw = torch.zeros(16 + cap * 16 + cap * C * 4, dtype=torch.uint8)
    raw[:16].view(torch.int32)[:] = torch.tensor([n, C, cap, 0], dtype=torch.int32)
    raw[16:16 + cap * 16].view(torch.int32).view(cap, 4)[:n] = torch.from_numpy(idx)
    raw[16 + cap * 16:].view(torch.float32).view(cap, C)[:n] = torch.from_numpy(rows)
    nz, Cc, nx, ny = 2, 3, 4, 7
    cells_all = B * nz * nx * ny
    cells = np.sort(rng.choice(cells_all, size=17, replace=False)).astype(np.int32)
    pitch = (len(cells) + 3) // 4 * 4
    cols = torch.zeros((Cc, pitch), dtype=torch.float32)
    cols[:, :len(cells)] = torch.from_numpy(rng.standard_normal((Cc, len(cells))).astype(np.float32))
    r = RowsResult(raw, cols, cap, C, torch.from_numpy(cells), len(cells), (B, C * Z, X, Y), (B, nz * Cc, nx, ny), nz)
    gi, gr = r.lidar()
    assert gi.shape == (n, 4) and torch.equal(gr, torch.from_numpy(rows))
    lidar, cam = r.dense()
    want_l = np.zeros((B, C, Z, X, Y), np.float32)
    for (b, x, y, z), row in zip(idx, rows):
        want_l[b, :, z, x, y] = row                      # dense() [N, C, X, Y, Z] -> permute(0, 1, 4, 2, 3)
    np.testing.assert_array_equal(lidar.numpy(), want_l.reshape(B, C * Z, X, Y))
    want_c = np.zeros((B, nz * Cc, nx, ny), np.float32)
    for i, cell in enumerate(cells):
        bz, xy = divmod(int(cell), nx * ny)
        b, z = divmod(bz, nz)
        want_c[b, z * Cc:(z + 1) * Cc, xy // ny, xy % ny] = cols[:, i].numpy()
    np.testing.assert_array_equal(cam.numpy(), want_c)
    assert r.nbytes() == 16 + n * 16 + n * C * 4 + Cc * pitch * 4
