"""The sync-free execution plan (StaticFrontEnd, eager and as a CUDA graph) gives bit-identical BEV maps to the
module path (BEVFrontEnd.forward), frame after frame with changing point counts."""
import pytest
import torch

from bevfusion_3d_object_detection_b200 import frontend, synthetic
from bevfusion_3d_object_detection_b200.static_frontend import StaticFrontEnd

pytestmark = pytest.mark.gpu


def _frames(n, sweeps):
    out = []
    for i in range(n):
        pts = synthetic.lidar_sweeps(n_sweeps=sweeps[i % len(sweeps)], seed=20 + i)
        depth, ctx = synthetic.camera_features(6, 118, 80, (32, 88), batch=1, seed=20 + i)
        out.append((torch.from_numpy(pts).cuda(), torch.from_numpy(depth).cuda(), torch.from_numpy(ctx).cuda()))
    return out


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_static_plan_matches_module_path(precision):
    torch.manual_seed(0)
    model = frontend.BEVFrontEnd(precision=precision).cuda().eval()
    rig = {k: torch.from_numpy(v).cuda() for k, v in synthetic.camera_rig(6, (256, 704), 1).items()}
    tables = model.set_calibration(rig)
    plan = StaticFrontEnd(model, tables, "cuda", batch=1, max_points=360000)
    frames = _frames(3, sweeps=[3, 10, 1])   # growing and shrinking point counts exercise the sentinel tail
    with torch.no_grad():
        want = [model([p], d, c, tables) for p, d, c in frames]
    # eager
    for (p, d, c), (wl, wc) in zip(frames, want):
        plan.load_inputs([p], d, c)
        gl, gc = plan.run()
        torch.cuda.synchronize()
        assert plan.error_codes() == [0] * len(plan.levels)
        assert torch.equal(gl, wl), float((gl - wl).abs().max())
        assert torch.equal(gc, wc)
    # captured graph, replayed per frame
    plan.capture()
    for (p, d, c), (wl, wc) in zip(frames, want):
        plan.load_inputs([p], d, c)
        gl, gc = plan.replay()
        torch.cuda.synchronize()
        assert torch.equal(gl, wl) and torch.equal(gc, wc)
    counts = plan.counts()
    assert counts[0] > 0 and all(c <= lv.cap for c, lv in zip(counts, plan.levels))


def test_static_plan_batch2():
    torch.manual_seed(1)
    model = frontend.BEVFrontEnd(precision="bf16").cuda().eval()
    rig = {k: torch.from_numpy(v).cuda() for k, v in synthetic.camera_rig(6, (256, 704), 2).items()}
    tables = model.set_calibration(rig)
    plan = StaticFrontEnd(model, tables, "cuda", batch=2, max_points=120000)
    pts = [torch.from_numpy(synthetic.lidar_sweeps(n_sweeps=2, seed=40 + k)).cuda() for k in range(2)]
    depth, ctx = synthetic.camera_features(6, 118, 80, (32, 88), batch=2, seed=41)
    depth, ctx = torch.from_numpy(depth).cuda(), torch.from_numpy(ctx).cuda()
    with torch.no_grad():
        wl, wc = model(pts, depth, ctx, tables)
    plan.load_inputs(pts, depth, ctx)
    gl, gc = plan.run()
    assert gl.shape == (2, 256, 180, 180) and torch.equal(gl, wl) and torch.equal(gc, wc)


def test_empty_and_tiny_frames():
    """A frame with no point inside the range, then one with a single voxel: every level empty / one site; both
    paths return the same (zero / single-column) maps without tripping on zero-sized launches."""
    torch.manual_seed(2)
    model = frontend.BEVFrontEnd(precision="bf16").cuda().eval()
    rig = {k: torch.from_numpy(v).cuda() for k, v in synthetic.camera_rig(6, (256, 704), 1).items()}
    tables = model.set_calibration(rig)
    depth, ctx = synthetic.camera_features(6, 118, 80, (32, 88), batch=1, seed=3)
    depth, ctx = torch.from_numpy(depth).cuda(), torch.from_numpy(ctx).cuda()
    plan = StaticFrontEnd(model, tables, "cuda", batch=1, max_points=5000)
    far = torch.full((100, 5), 500.0, device="cuda")                       # all out of range
    one = torch.tensor([[1.0, 2.0, -1.0, 7.0, 0.0]] * 3, device="cuda")     # three points, one voxel
    for pts, n_sites in ((far, 0), (one, 1), (far, 0)):
        with torch.no_grad():
            wl, wc = model([pts], depth, ctx, tables)
        plan.load_inputs([pts], depth, ctx)
        gl, gc = plan.run()
        torch.cuda.synchronize()
        assert plan.counts()[0] == n_sites
        assert torch.equal(gl, wl) and torch.equal(gc, wc)
        if n_sites == 0:
            assert float(gl.abs().max()) == 0.0
        else:
            assert float(gl.abs().max()) > 0.0


@pytest.mark.parametrize("batch,mode", [(1, "rows"), (2, "rows"), (1, "rows_zero_copy")])
def test_host_pipeline_rows_output_is_lossless(batch, mode):
    """HostPipeline's "rows" output (active rows + coordinates written into pinned host buffers by the pack kernels)
    rebuilds, on the host, the dense fp32 maps of the default output bit for bit; frame after frame on every slot."""
    torch.manual_seed(2)
    model = frontend.BEVFrontEnd(precision="bf16").cuda().eval()
    rig = {k: torch.from_numpy(v).cuda() for k, v in synthetic.camera_rig(6, (256, 704), batch).items()}
    tables = model.set_calibration(rig)
    frames = []
    for i in range(4):
        pts = [torch.from_numpy(synthetic.lidar_sweeps(n_sweeps=[2, 5, 1, 3][i], seed=60 + 7 * i + k)).pin_memory()
               for k in range(batch)]
        depth, ctx = synthetic.camera_features(6, 118, 80, (32, 88), batch=batch, seed=60 + i)
        frames.append((pts, torch.from_numpy(depth).pin_memory(), torch.from_numpy(ctx).pin_memory()))
    pipe = frontend.HostPipeline(model, tables, "cuda", depth=2, batch=batch, max_points=200000,
                                 example=([p.cuda() for p in frames[1][0]], frames[1][1].cuda(), frames[1][2].cuda()))
    dense = []
    for pts, depth, ctx in frames:
        lid, cam = pipe.result(pipe.submit(pts, depth, ctx))
        dense.append((lid.clone(), cam.clone()))
    slots = [pipe.submit(pts, depth, ctx, output=mode) for pts, depth, ctx in frames[:2]]   # two frames in flight
    got = [pipe.result(s).dense() for s in slots]
    for pts, depth, ctx in frames[2:]:
        got.append(pipe.result(pipe.submit(pts, depth, ctx, output=mode)).dense())
    for (wl, wc), (gl, gc) in zip(dense, got):
        assert torch.equal(gl, wl) and torch.equal(gc, wc)
    r = pipe.result(pipe.submit(*frames[0], output=mode))
    idx, rows = r.lidar()
    assert idx.shape[0] == rows.shape[0] > 0 and rows.shape[1] == 128 and int(idx[:, 0].max()) == batch - 1
    assert r.nbytes() < 0.6 * (dense[0][0].numel() + dense[0][1].numel()) * 4
    # mixing output forms on one pipeline keeps working
    lid, cam = pipe.result(pipe.submit(*frames[3]))
    assert torch.equal(lid, dense[3][0]) and torch.equal(cam, dense[3][1])
