#!/usr/bin/env python
"""BEV front-end throughput on B200 (BASELINE.json metric: frames/s; bev_pool / voxelize GB/s and sparse-conv TFLOP/s
against the measured peaks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--precision bf16|fp32]
                    [--config infer|train|custom|stress]

--config infer  (default) BASELINE configs[1]: one nuScenes-shaped frame per GPU per step: 10-sweep LiDAR voxelize(+mean)
                -> 21-conv sparse encoder -> dense BEV, and 6-camera (depth, context) -> fused bev_pool -> BEV.
--config train  configs[2]: forward + backward of the sparse encoder and the fused bev_pool, batch 4 per GPU, bucketed
                NCCL gradient all-reduce overlapped with backward + SGD update.
--config custom configs[3]: 5 cameras 48x88, 3-dim points, batch 8 per GPU inference.
--config stress configs[4]: ~0.9 M-point 128-beam sweep at 0.05 m voxels + 236-bin bev_pool, 2 frames per GPU
                (16 over 8 GPUs).
Weak scaling: frames are independent, no data-path collective.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the CPU arm mixes two OpenMP runtimes (torch's and the C oracle's); spinning idle threads of one steal the cores of
# the other (3x slower), so both are told to sleep when idle.  Must be set before either runtime loads.
os.environ.setdefault("OMP_WAIT_POLICY", "passive")
os.environ.setdefault("GOMP_SPINCOUNT", "0")
if "--impl" in sys.argv and "reference" in sys.argv and int(os.environ.get("RANK", "0")) == 0:
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the CPU arm runs on rank 0 alone with all host threads
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

METRIC = "bev_frontend_frames_per_sec"
L2_BYTES = 126e6
NUS_RANGE = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]

CONFIGS = {
    "infer": dict(
        key="configs[1]", batch=1, n_cams=6, image=(256, 704), feat=(32, 88), dbound=[1.0, 60.0, 0.5], c_ctx=80,
        dims=5, voxel=[0.075, 0.075, 0.2], grid=[1440, 1440, 41], max_voxels=(120000, 160000),
        lidar=dict(n_sweeps=10),
        workload=("configs[1]: BEVFusion camera+LiDAR nuScenes front end, batch 1, 10-sweep ~320k pts x 5 dims, "
                  "1440x1440x41 sparse grid (hard voxelize max 10 pts / 160000 voxels + mean, 21-conv sparse encoder), "
                  "6 cams x 118 depth x 32x88 x 80 ch fused bev_pool -> 360x360")),
    "train": dict(
        key="configs[2]", batch=4, n_cams=6, image=(256, 704), feat=(32, 88), dbound=[1.0, 60.0, 0.5], c_ctx=80,
        dims=5, voxel=[0.075, 0.075, 0.2], grid=[1440, 1440, 41], max_voxels=(120000, 160000),
        lidar=dict(n_sweeps=10),
        workload=("configs[2]: BEVFusion training forward+backward of the fused bev_pool and the 21-conv sparse encoder "
                  "(train-mode BatchNorm), batch 4 frames per GPU (10-sweep ~320k pts, 120000-voxel training cap, 6 cams "
                  "x 118 x 32x88 x 80 ch), frame-parallel, bucketed NCCL gradient all-reduce overlapped with backward, "
                  "SGD update")),
    "custom": dict(
        key="configs[3]", batch=8, n_cams=5, image=(384, 704), feat=(48, 88), dbound=[1.0, 60.0, 0.5], c_ctx=80,
        dims=3, voxel=[0.075, 0.075, 0.2], grid=[1440, 1440, 41], max_voxels=(120000, 160000),
        lidar=dict(n_sweeps=3, dims=3),
        workload=("configs[3]: the repo's 5-class custom-dataset BEVFusion config (configs/custom_data/"
                  "lidar-cam_custom.py:47-57, lidar_custom.py:43-53: 5 cams 384x704 -> 48x88 features, 3-dim points, "
                  "range/voxel size as nuScenes), batch 8 inference per GPU, 3-sweep ~96k pts per frame")),
    "stress": dict(
        key="configs[4]", batch=2, n_cams=6, image=(256, 704), feat=(32, 88), dbound=[1.0, 60.0, 0.25], c_ctx=80,
        dims=5, voxel=[0.05, 0.05, 0.2], grid=[2160, 2160, 41], max_voxels=(500000, 500000),
        lidar=dict(n_sweeps=1, beams=128, azimuth=8192, elev_deg=(-25.0, 15.0)),
        workload=("configs[4]: stress -- 128-beam ~0.9 M-point sweep at 0.05 m voxels (2160x2160x41 grid, 500000-voxel "
                  "cap) + 236-bin depth bev_pool (6 cams x 236 x 32x88 x 80 ch), 2 frames per GPU (batch 16 over 8 "
                  "GPUs)")),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tc=float(d["bf16_tflops"]), tc_sustained=float(d["bf16_tflops_sustained"]),
                    src="measured")
    return dict(hbm=6650.0, tc=1590.0, tc_sustained=1400.0, src="fallback")


def n_depth_bins(cfg):
    d0, d1, dd = cfg["dbound"]
    return int(np.arange(d0, d1, dd).shape[0])


def make_frames(cfg, n, seed0=0):
    from bevfusion_3d_object_detection_b200 import synthetic

    frames = []
    for i in range(n):
        pts = synthetic.lidar_sweeps(seed=seed0 + i, **cfg["lidar"])
        depth, ctx = synthetic.camera_features(cfg["n_cams"], n_depth_bins(cfg), cfg["c_ctx"], cfg["feat"], batch=1,
                                               seed=seed0 + i)
        frames.append(dict(points=pts, depth=depth, ctx=ctx))
    return frames


def batches_of(frames, batch):
    """[frame] -> [dict(points=[...], depth=[B*N,...], ctx=[B*N,...])] numpy."""
    out = []
    for b0 in range(0, len(frames) - batch + 1, batch):
        fs = frames[b0:b0 + batch]
        out.append(dict(points=[f["points"] for f in fs], depth=np.concatenate([f["depth"] for f in fs], 0),
                        ctx=np.concatenate([f["ctx"] for f in fs], 0)))
    return out


def build_model(cfg, precision, dev):
    import torch

    from bevfusion_3d_object_detection_b200 import frontend, synthetic
    from bevfusion_3d_object_detection_b200.sparse_encoder import NUSCENES_ENCODER_CFG

    enc_cfg = dict(NUSCENES_ENCODER_CFG)
    enc_cfg.update(in_channels=cfg["dims"], sparse_shape=list(cfg["grid"]))
    vox_cfg = dict(max_num_points=10, point_cloud_range=NUS_RANGE, voxel_size=cfg["voxel"], max_voxels=cfg["max_voxels"])
    view_cfg = dict(in_channels=256, out_channels=cfg["c_ctx"], image_size=cfg["image"], feature_size=cfg["feat"],
                    xbound=[-54.0, 54.0, 0.3], ybound=[-54.0, 54.0, 0.3], zbound=[-10.0, 10.0, 20.0],
                    dbound=cfg["dbound"])
    torch.manual_seed(0)
    model = frontend.BEVFrontEnd(voxelize_cfg=vox_cfg, encoder_cfg=enc_cfg, view_cfg=view_cfg, precision=precision)
    synthetic.init_encoder_weights(model.pts_middle_encoder, seed=0)   # no checkpoints offline: seeded He-style init
    return model.to(dev)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax.append(float(p[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU formulation on the host cores
# ---------------------------------------------------------------------------------------------------------------
class CpuFrontEnd:
    """oracle/cpu_frontend (reference C++ hard_voxelize_cpu + torch gather-mm-scatter sparse encoder + torch
    outer-product / index_add_ bev_pool) with the weights of `model` (or a freshly seeded model of the same config)."""

    def __init__(self, cfg, model=None):
        import torch

        import oracle
        from oracle import cpu_frontend
        from bevfusion_3d_object_detection_b200 import synthetic

        oracle.build()
        self.t, self.cf, self.cfg = torch, cpu_frontend, cfg
        if model is None:
            model = build_model(cfg, "fp32", "cpu")
        self.model = model
        self.plan = cpu_frontend.encoder_plan(model.pts_middle_encoder)
        self.vox = model.pts_voxel_layer
        self.n_cams = cfg["n_cams"]
        self.vt = build_model(cfg, "fp32", "cpu").view_transform
        self.nx = [int(v) for v in self.vt.nx]
        self._aux = {}
        self.cores = torch.get_num_threads()
        self.kind = "reference" if cpu_frontend.ref_voxel_module() is not None else "port"
        self.synthetic = synthetic

    def aux(self, batch):
        if batch not in self._aux:
            t = self.t
            rig = {k: t.from_numpy(v) for k, v in self.synthetic.camera_rig(self.n_cams, self.cfg["image"], batch).items()}
            with t.no_grad():
                geom = self.vt.get_geometry(**rig)
                gf, kept, _, indices = self.vt.bev_pool_aux(geom)
            self._aux[batch] = (gf, kept, indices)
        return self._aux[batch]

    def voxelize(self, points_list, training=False):
        t, cf = self.t, self.cf
        feats, coords = [], []
        cap = self.vox.max_voxels[0 if training else 1]
        for k, pts in enumerate(points_list):
            f, c, _ = cf.cpu_voxelize_mean(pts, self.vox.voxel_size, self.vox.point_cloud_range, self.vox.max_num_points,
                                           cap)
            c = c.clone()
            c[:, 0] = k
            feats.append(f)
            coords.append(c)
        return t.cat(feats, 0), t.cat(coords, 0)

    def step(self, batch):
        """batch: dict(points=[np], depth=np, ctx=np) -> (lidar_bev, cam_bev) torch CPU."""
        t, cf = self.t, self.cf
        B = len(batch["points"])
        gf, kept, indices = self.aux(B)
        with t.no_grad():
            feats, coords = self.voxelize(batch["points"])
            lidar = cf.cpu_sparse_encoder(self.plan, feats, coords.numpy(), list(self.cfg["grid"]), B)
            cam = cf.cpu_bev_pool(t.from_numpy(batch["depth"]), t.from_numpy(batch["ctx"]), kept, indices, gf, B,
                                  self.n_cams, self.nx[2], self.nx[0], self.nx[1])
        return lidar, cam

    def train_step(self, batch):
        """forward + backward (torch autograd through the same formulation, batch-statistics BatchNorm)."""
        t, cf = self.t, self.cf
        B = len(batch["points"])
        gf, kept, indices = self.aux(B)
        leaves = []
        for L in self.plan:
            for k, v in L.items():
                if t.is_tensor(v) and v.dim() > 1:
                    v.requires_grad_(True)
                    v.grad = None
                    leaves.append(v)
        feats, coords = self.voxelize(batch["points"], training=True)
        lidar = cf.cpu_sparse_encoder(self.plan, feats, coords.numpy(), list(self.cfg["grid"]), B, train=True)
        d = t.from_numpy(batch["depth"]).requires_grad_(True)
        c = t.from_numpy(batch["ctx"]).requires_grad_(True)
        cam = cf.cpu_bev_pool(d, c, kept, indices, gf, B, self.n_cams, self.nx[2], self.nx[0], self.nx[1])
        (lidar.sum() + cam.sum()).backward()
        return float(lidar.detach().abs().sum())

    def describe(self):
        v = "reference C++ hard_voxelize_cpu (oracle/_ref)" if self.kind == "reference" else "C port of hard_voxelize"
        return (f"{v} + torch gather-mm-scatter sparse encoder (mmcv CPU indice_conv formulation, rulebook from the C "
                f"oracle) + torch outer-product/index_add_ bev_pool")


def arm_config(cfg, args, frame_bytes):
    """The `config` object of the JSON line: built the same way by BOTH arms (the reference arm reports the configuration
    of the comparison, i.e. this arm's, verbatim: same workload, same batch per step, same input rotation)."""
    B = cfg["batch"]
    if args.config == "train":
        return dict(workload=cfg["workload"], frames_per_gpu_per_step=B, precision=args.precision,
                    l2="inputs rotate over 2 distinct batches (%.0f MB > 126 MB L2)" % (2 * B * frame_bytes / 1e6),
                    parallelism="frame-parallel; gradient buckets all-reduced over NCCL as they complete")
    ring = max(2, int(np.ceil(1.25 * L2_BYTES / (frame_bytes * B))))
    return dict(workload=cfg["workload"], frames_per_gpu_per_step=B, precision=args.precision,
                mode="one CUDA graph per batch, device-side row counts, %d batch(es) in flight"
                     % min(args.inflight, args.inflight_device),
                l2="inputs rotate over %d distinct batches (%.0f MB > 126 MB L2)" % (ring, ring * B * frame_bytes / 1e6),
                parallelism="frame-parallel, no data-path collective")


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch

    torch.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; this arm uses the whole host
    cfg = CONFIGS[args.config]
    cpu = CpuFrontEnd(cfg)
    train = args.config == "train"
    # bounded sample per step: one frame of the batch (the formulation is per-sample; frames/s does not depend on it)
    frames = make_frames(cfg, 2)
    batches = batches_of(frames, 1)
    fn = cpu.train_step if train else cpu.step
    for i in range(args.warmup):
        fn(batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        fn(batches[i % len(batches)])
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    sample = ("one frame per step (1 of the %d frames of a batch; per-sample formulation): " % cfg["batch"]) + cpu.describe()
    if train:
        sample += "; forward + backward by torch autograd through the same formulation"
    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=arm_config(cfg, args, sum(int(v.nbytes) for v in frames[0].values())),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=cpu.cores, kind=cpu.kind, sample=sample),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# timing helpers
# ---------------------------------------------------------------------------------------------------------------
def rel_err(got, ref):
    """max / mean |got - ref| relative to the reference map's scale (max |ref|)."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    s = float(np.abs(ref).max()) or 1.0
    return dict(max_rel=float(np.abs(got - ref).max() / s), mean_rel=float(np.abs(got - ref).mean() / s), scale=s)


def graph_ms(torch, fn, n_inner, reps):
    """Device time per call of a sync-free `fn(i)`: a CUDA graph of n_inner consecutive calls, replayed reps times."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for i in range(n_inner):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        keep = [fn(i) for i in range(n_inner)]
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / (reps * n_inner)


def event_ms(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def latest_traffic():
    """DRAM bytes of the frame's GEMM launches from the newest committed `ncu --set full` capture (profiles/r*_traffic.json)."""
    pd = os.path.join(ROOT, "profiles")
    best = None
    if os.path.isdir(pd):
        for f in sorted(os.listdir(pd)):
            if f.endswith("_traffic.json"):
                best = os.path.join(pd, f)
    return (json.load(open(best)), os.path.basename(best)) if best else (None, None)


def hostlink_gbs(torch, dev, h2d_bytes, d2h_bytes, reps=6):
    """Raw pinned-memory copy bandwidth of this rank with ALL ranks copying at once: one H2D of h2d_bytes and one D2H of
    d2h_bytes per 'frame' on two streams (what the e2e pipeline moves), device-timed.  -> GB/s over both directions."""
    hin = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    hout = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    dout = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for warm in (True, False):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(2 if warm else reps):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return (h2d_bytes + d2h_bytes) / ms / 1e6, ms


# ---------------------------------------------------------------------------------------------------------------
# B200 arm: inference configs (infer / custom / stress)
# ---------------------------------------------------------------------------------------------------------------
def run_frontend(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from bevfusion_3d_object_detection_b200 import _lib, frontend, parallel, synthetic
    from bevfusion_3d_object_detection_b200.static_frontend import StaticFrontEnd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    pk = peaks()
    cfg = CONFIGS[args.config]
    B = cfg["batch"]
    D_BINS = n_depth_bins(cfg)
    fh, fw = cfg["feat"]

    model = build_model(cfg, args.precision, dev).eval()
    rig = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(cfg["n_cams"], cfg["image"], B).items()}
    tables = model.set_calibration(rig)

    # inputs rotate over `ring` distinct batches whose total size exceeds L2 (every rank runs the same batches: per-GPU
    # work is identical, weak scaling)
    probe = make_frames(cfg, 1)[0]
    frame_bytes = sum(int(v.nbytes) for v in probe.values())
    ring = max(2, int(np.ceil(1.25 * L2_BYTES / (frame_bytes * B))))
    frames = [probe] + make_frames(cfg, ring * B, seed0=0)[1:]
    batches = batches_of(frames, B)
    dev_b = [dict(points=[torch.from_numpy(p).to(dev) for p in b["points"]], depth=torch.from_numpy(b["depth"]).to(dev),
                  ctx=torch.from_numpy(b["ctx"]).to(dev)) for b in batches]
    pin_b = [dict(points=[torch.from_numpy(p).pin_memory() for p in b["points"]],
                  depth=torch.from_numpy(b["depth"]).pin_memory(), ctx=torch.from_numpy(b["ctx"]).pin_memory())
             for b in batches]
    max_pts = max(int(p.shape[0]) for b in batches for p in b["points"]) + 4096
    ex = dev_b[0]
    growth = None if args.config == "infer" else 3.0   # level capacities: worst case for configs[1], 3x growth else

    plan = StaticFrontEnd(model, tables, dev, batch=B, max_points=max_pts, level_growth=growth)
    plan.load_inputs(ex["points"], ex["depth"], ex["ctx"])
    n0 = L.bevf_launch_count()
    plan.run()
    launches_per_frame = int(L.bevf_launch_count() - n0)
    plan.capture()
    plan.check_capacity()
    pipe = frontend.HostPipeline(model, tables, dev, depth=max(2, args.inflight), batch=B, max_points=max_pts,
                                 example=(ex["points"], ex["depth"], ex["ctx"]), level_growth=growth)
    out_host = {}

    def step_dev(i):
        f = dev_b[i % ring]
        if args.inflight > 1:   # plans on their own streams: consecutive frames overlap on the GPU
            return pipe.submit_device(f["points"], f["depth"], f["ctx"], ways=args.inflight_device)
        plan.load_inputs(f["points"], f["depth"], f["ctx"])   # device -> static input buffers
        return plan.replay()                                    # the whole batch: one CUDA graph

    def step_e2e(i):
        # host buffers in, host buffers out: H2D / compute / D2H on three streams, outputs multi-buffered
        f = pin_b[i % ring]
        out_host["slot"] = pipe.submit(f["points"], f["depth"], f["ctx"], output=args.output)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, fin=None):
        with torch.no_grad():
            for i in range(warmup):
                fn(i)
            if fin is not None:
                fin()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(warmup + i)
            if fin is not None:
                fin()   # the timing stream waits for the last frame (and its device->host copy)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
        _, ms, _ = parallel.job_throughput(steps, ms, device=dev)   # max over ranks
        return ms

    fin = (lambda: pipe.join()) if args.inflight > 1 else None
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, args.steps, args.warmup, fin=fin)
    clocks = sampler.stop() if rank == 0 else None
    # repeated timed regions (same K steps each): the headline region above is the first; report the spread
    rep_ms = [ms] + [timed(step_dev, args.steps, 0, fin=fin) for _ in range(max(0, args.repeats - 1))]
    ms_e2e = timed(step_e2e, args.steps, args.warmup, fin=lambda: pipe.join())
    rep_e2e = [ms_e2e] + [timed(step_e2e, args.steps, 0, fin=lambda: pipe.join()) for _ in range(max(0, args.repeats - 1))]
    res_host = pipe.result(out_host["slot"])
    rows_mode = args.output == "rows"
    if rows_mode:   # --output rows: the headline e2e itself uses the lossless rows form
        d2h_rows = int(res_host.nbytes())
        lid_h, cam_h = res_host.dense()
    else:
        lid_h, cam_h = res_host[0], res_host[1]
    # the same pipeline with the LOSSLESS "rows" output (active rows + coordinates instead of the dense maps; the pack
    # kernels store straight into the pinned host buffers): reported beside the headline e2e, which keeps the dense maps
    e2e_rows = None
    try:
        def step_rows(i):
            f = pin_b[i % ring]
            out_host["rows_slot"] = pipe.submit(f["points"], f["depth"], f["ctx"], output="rows")

        ms_rows = timed(step_rows, args.steps, args.warmup, fin=lambda: pipe.join())
        rr = pipe.result(out_host["rows_slot"])
        lid_r, cam_r = rr.dense()
        f_last = pin_b[(args.warmup + args.steps - 1) % ring]
        lid_d, cam_d = pipe.result(pipe.submit(f_last["points"], f_last["depth"], f_last["ctx"]))   # dense maps of the same frame
        e2e_rows = dict(value=world * B * args.steps / (ms_rows / 1e3), unit="frames/s", ms_per_step=ms_rows / args.steps,
                        d2h_bytes_per_step=int(rr.nbytes()), h2d_bytes_per_step=None,
                        lossless=bool(torch.equal(lid_r, lid_d) and torch.equal(cam_r, cam_d)),
                        output="fp32 active rows + coordinates (LiDAR map), fp32 columns of the reached cells (camera map); "
                               "dense maps rebuilt on the host bit for bit (`lossless`)")
    except Exception as exc:   # never let the extra block take the headline line down
        print(f"[bench] rows-output timing unavailable: {exc}", file=sys.stderr)
    assert bool(torch.isfinite(lid_h.float()).all()) and float(cam_h.float().abs().sum()) > 0.0
    h2d = sum(int(v.numel() * v.element_size()) for v in pin_b[0]["points"]) + \
        int(pin_b[0]["depth"].numel() * 4 + pin_b[0]["ctx"].numel() * 4)
    if e2e_rows is not None:
        e2e_rows["h2d_bytes_per_step"] = h2d
    d2h = d2h_rows if rows_mode else int(sum(t.numel() * t.element_size() for t in res_host))
    # raw host-link ceiling for exactly these bytes, all ranks copying at once (what bounds e2e at N = 8)
    barrier()
    link_gbs, link_ms = hostlink_gbs(torch, dev, h2d, d2h)
    barrier()
    if world > 1:
        t_l = torch.tensor([link_gbs], dtype=torch.float64, device=dev)
        dist.all_reduce(t_l, op=dist.ReduceOp.SUM)
        link_total = float(t_l.item())
    else:
        link_total = link_gbs

    # ---- per-stage device times (same rotating inputs), for the roofline objects -------------------------------
    from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer as _vl

    stages = {}
    cap_v = cfg["max_voxels"][1]
    c_pts = cfg["dims"]
    _vf = torch.empty((cap_v, c_pts), device=dev)
    _vc = torch.empty((cap_v, 4), dtype=torch.int32, device=dev)
    _vs = torch.empty((cap_v,), dtype=torch.int32, device=dev)
    flat_pts = [p for b in dev_b for p in b["points"]]
    n_inner = max(len(flat_pts), int(np.ceil(1.25 * L2_BYTES / max(1, flat_pts[0].numel() * 4))))

    def f_vox_async(i):   # the sync-free C-ABI form the static plan uses (device-side voxel count)
        return _vl.voxelize_mean(flat_pts[i % len(flat_pts)], _vf, _vc, _vs, cfg["voxel"], NUS_RANGE, 10, cap_v)

    vnum = f_vox_async(0)
    torch.cuda.synchronize()
    m_vox, n_pts = int(vnum.item()), int(flat_pts[0].shape[0])
    ms_vox = graph_ms(torch, f_vox_async, min(n_inner, 4 * len(flat_pts)), 5)
    vox_bytes = 4 * c_pts * n_pts + m_vox * (4 * c_pts + 16 + 4)
    stages["voxelize_mean"] = dict(ms=ms_vox, bytes=vox_bytes, gbs=vox_bytes / ms_vox / 1e6,
                                   frac=vox_bytes / ms_vox / 1e6 / pk["hbm"], points=n_pts, voxels=m_vox,
                                   timing="CUDA graph over %d distinct sweeps (sync-free C-ABI form), one frame per call; "
                                          "the sweeps total %.0f MB, i.e. they stay L2-resident between replays (the op is bound "
                                          "by L2 atomics / dependent latency, not by DRAM: profiles/README.md r2p; "
                                          "scripts/vox_times.py times it over 128 MB of sweeps: same number)"
                                          % (len(flat_pts), len(flat_pts) * flat_pts[0].numel() * 4 / 1e6))

    def f_pool(i):
        f = dev_b[i % ring]
        return model.extract_img_bev(f["depth"], f["ctx"], tables)

    ms_pool = graph_ms(torch, f_pool, ring, 5)
    pool_bytes = (4 * B * cfg["n_cams"] * fh * fw * (D_BINS + cfg["c_ctx"]) + 4 * tables.nk + 8 * tables.n_intervals
                  + 4 * cfg["c_ctx"] * B * 360 * 360)
    stages["bev_pool_fused"] = dict(ms=ms_pool, bytes=pool_bytes, gbs=pool_bytes / ms_pool / 1e6,
                                    frac=pool_bytes / ms_pool / 1e6 / pk["hbm"], nk=tables.nk, batch=B,
                                    n_intervals=tables.n_intervals,
                                    timing="CUDA graph of %d calls (batch %d each, inputs rotate: > L2)" % (ring, B))
    # batched camera branch (the >= 60 % bar is meaningful batched): batch 4 for configs[1]
    if args.config == "infer":
        try:
            B4 = 4
            rig4 = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(cfg["n_cams"], cfg["image"], B4).items()}
            vt4 = build_model(cfg, args.precision, dev).view_transform
            tab4 = vt4.build_tables(vt4.get_geometry(**rig4))
            d4 = [torch.cat([dev_b[(j + q) % ring]["depth"] for q in range(B4)], 0) for j in range(2)]
            c4 = [torch.cat([dev_b[(j + q) % ring]["ctx"] for q in range(B4)], 0) for j in range(2)]
            ms_p4 = graph_ms(torch, lambda i: vt4.pool_fused(d4[i % 2], c4[i % 2], tab4), 2, 5)
            by4 = (4 * B4 * cfg["n_cams"] * fh * fw * (D_BINS + cfg["c_ctx"]) + 4 * tab4.nk + 8 * tab4.n_intervals
                   + 4 * cfg["c_ctx"] * B4 * 360 * 360)
            stages["bev_pool_fused_batch4"] = dict(ms=ms_p4, bytes=by4, gbs=by4 / ms_p4 / 1e6,
                                                   frac=by4 / ms_p4 / 1e6 / pk["hbm"], batch=B4,
                                                   timing="CUDA graph, 2 distinct batches of 4 frames (185 MB > L2)")
            del d4, c4, tab4, vt4
        except Exception as exc:
            print(f"[bench] batched bev_pool timing unavailable: {exc}", file=sys.stderr)

    # per-launch device times of the gather-GEMMs, on the same launch sequence the graph holds
    plan.load_inputs(ex["points"], ex["depth"], ex["ctx"])
    layers = plan.profile(reps=5)
    gemm_ms = sum(r["ms"] for r in layers)
    gemm_flops = sum(r["flops"] for r in layers)
    n_gemm = len(layers)
    stages["sparse_encoder"] = dict(gemm_ms=gemm_ms, gemm_launches=n_gemm, gemm_flops=gemm_flops,
                                    gemm_tflops=gemm_flops / max(gemm_ms, 1e-9) / 1e9,
                                    gemm_frac=gemm_flops / max(gemm_ms, 1e-9) / 1e9 / pk["tc"], sites=plan.counts(),
                                    layers=[dict(cin=r["cin"], cout=r["cout"], subm=r["subm"], rows=r["rows"],
                                                 us=round(1e3 * r["ms"], 2),
                                                 tflops=round(r["flops"] / max(r["ms"], 1e-9) / 1e9, 1)) for r in layers])

    if args.config == "infer":
        _boundary_and_reference_gpu(torch, stages, tables, dev_b, cfg, pk, m_vox, dev)
        _depth_prep_stage(torch, stages, dev_b, cfg, pk, dev)
        _tables_stage(torch, stages, model, rig, tables)

    # the dominant kernel family of the step
    if args.precision == "bf16":
        kname = "spconv_ts_kernel x%d (gathered operand in TMEM), aggregate" % n_gemm
    else:
        kname = "spconv_gemm_f32_kernel x%d (fp32 FFMA parity path), aggregate" % n_gemm
    roof = dict(kernel=kname, bound="tensor", achieved=stages["sparse_encoder"]["gemm_tflops"], peak=pk["tc"],
                unit="TFLOP/s", frac=stages["sparse_encoder"]["gemm_frac"], traffic=None, peak_source=pk["src"],
                launches_per_frame=n_gemm, avg_launch_ms=gemm_ms / max(n_gemm, 1),
                note="achieved = useful flops (2 x valid rulebook pairs x Cin x Cout, counted from the rulebooks) / sum "
                     "of per-launch CUDA-event times; peak = measured dense bf16 burst")
    tr, tr_name = latest_traffic()
    if args.precision == "bf16" and args.config == "infer" and tr is not None:
        roof["traffic"] = (tr["dram_read_bytes"] + tr["dram_write_bytes"]) / n_gemm
        roof["traffic_note"] = ("DRAM bytes per launch, mean over the frame's launches (profiles/%s; ncu flushes the caches "
                                "before every kernel: cold-cache figure)" % tr_name)
        warm = os.path.join(ROOT, "profiles", "r2z_gemm_warm_dram.json")
        if os.path.exists(warm):   # same capture with --cache-control none: the launches back to back, as in the frame
            tw = json.load(open(warm))
            roof["traffic_no_flush"] = (tw["dram_read_bytes"] + tw["dram_write_bytes"]) / n_gemm

    # ---- fp32 parity mode (reference precision, bevfusion.py:177,201 runs the encoder with autocast off) -------
    fp32_block = None
    if args.precision == "bf16" and args.config == "infer" and not args.no_fp32:
        try:
            fp32_block = _fp32_line(torch, cfg, dev, rig, dev_b, pin_b, max_pts, ring)
        except Exception as exc:
            print(f"[bench] fp32 measurement unavailable: {exc}", file=sys.stderr)

    # ---- training stage (configs[2]) inside the default line: one GPU's share, all-reduce active when N > 1 ----
    if args.config == "infer" and not args.no_train_stage:
        try:
            torch.cuda.empty_cache()   # the eager training step allocates its temporaries: start from a clean cache
            stages["training"] = _train_measure(torch, dist, CONFIGS["train"], args.precision, dev, rank, world, steps=6,
                                                warmup=3)["summary"]
        except Exception as exc:
            print(f"[bench] training-stage timing unavailable: {exc}", file=sys.stderr)

    # GPU outputs of one batch for the parity check below (eager run of the plan on batch 1 of the ring)
    chk = dev_b[1 % ring]
    plan.load_inputs(chk["points"], chk["depth"], chk["ctx"])
    lid_g, cam_g = plan.run()
    torch.cuda.synchronize()
    lid_g, cam_g = lid_g.cpu().numpy(), cam_g.cpu().numpy()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- cpu_baseline + parity (rank 0, N == 1 only): bounded sample --------------------------------------------
    cpu_baseline, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        cpu = CpuFrontEnd(cfg, model=build_cpu_twin(model, cfg))
        if args.config == "infer":
            cpu.step(batches[0])
        t0 = time.perf_counter()
        lid_c, cam_c = cpu.step(batches[1 % ring])
        n_t = 1
        if args.config == "infer":
            cpu.step(batches[2 % ring])
            n_t = 2
        dt = (time.perf_counter() - t0) / n_t
        cpu_baseline = dict(value=B / dt, unit="frames/s", cores=cpu.cores, kind=cpu.kind,
                            sample=("%d batch(es) of %d frame(s)%s; " % (n_t, B, " after 1 warm-up" if args.config == "infer" else ""))
                            + cpu.describe())
        parity = dict(lidar_bev=rel_err(lid_g, lid_c.numpy()), camera_bev=rel_err(cam_g, cam_c.numpy()),
                      precision=args.precision,
                      against="oracle/cpu_frontend (fp32) on the same batch and weights; error relative to max|ref|; "
                              "north_star bar: 1e-5 (fp32), 2e-2 (bf16)")
        if fp32_block is not None and fp32_block.get("_lidar") is not None:
            fp32_block["parity"] = dict(lidar_bev=rel_err(fp32_block.pop("_lidar"), lid_c.numpy()),
                                        camera_bev=rel_err(fp32_block.pop("_cam"), cam_c.numpy()))
    if fp32_block is not None:
        fp32_block.pop("_lidar", None)
        fp32_block.pop("_cam", None)

    fps = world * B * args.steps / (ms / 1e3)
    fps_e2e = world * B * args.steps / (ms_e2e / 1e3)
    med = float(np.median(rep_ms))
    med_e2e = float(np.median(rep_e2e))
    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=("bf16 sparse conv (fp32 accumulate) + fp32 voxelize/bev_pool" if args.precision == "bf16"
                       else "f32"),
                data="synthetic",
                config=arm_config(cfg, args, frame_bytes),
                e2e=dict(value=fps_e2e, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                         ms_per_step=ms_e2e / args.steps, output={"dense": "fp32 dense BEV maps", "bf16": "compact (bf16 dense maps, lossy)",
                                 "rows": "lossless fp32 rows: active LiDAR rows + coordinates, camera columns of the "
                                         "reached cells"}[args.output],
                         median_value=world * B * args.steps / (med_e2e / 1e3),
                         hostlink=dict(gbs_all_ranks=link_total, gbs_this_rank=link_gbs, ms_per_step_copies_only=link_ms,
                                       ceiling_frames_per_s=world * B * 1e3 / link_ms,
                                       note="raw pinned H2D+D2H of exactly these bytes on two streams, all ranks at "
                                            "once: the host-link ceiling of e2e")),
                e2e_rows=e2e_rows,
                repeats=dict(n=len(rep_ms), value_median=world * B * args.steps / (med / 1e3),
                             value_min=world * B * args.steps / (max(rep_ms) / 1e3),
                             value_max=world * B * args.steps / (min(rep_ms) / 1e3)),
                gpu_launches=int(launches_per_frame * args.steps), roofline=roof, stages=stages, parity=parity,
                fp32=fp32_block, cpu_baseline=cpu_baseline, clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def build_cpu_twin(model, cfg):
    """CPU copy of the model's encoder weights (the CPU arm must multiply the same numbers)."""
    import torch

    twin = build_model(cfg, "fp32", "cpu")
    src = dict(model.pts_middle_encoder.state_dict())
    with torch.no_grad():   # tensor-by-tensor: a plain dict would lose the state dict's version metadata (layout shim)
        for k, v in twin.pts_middle_encoder.state_dict().items():
            v.copy_(src[k].detach().cpu())
    return twin.eval()


def _fp32_line(torch, cfg, dev, rig, dev_b, pin_b, max_pts, ring, steps=6):
    """Same step with the fp32 FFMA sparse-conv path (1e-5 parity with the fp32 reference)."""
    from bevfusion_3d_object_detection_b200 import frontend
    from bevfusion_3d_object_detection_b200.static_frontend import StaticFrontEnd

    B = cfg["batch"]
    m32 = build_model(cfg, "fp32", dev).eval()
    tab = m32.set_calibration(rig)
    ex = dev_b[0]
    p32 = StaticFrontEnd(m32, tab, dev, batch=B, max_points=max_pts)
    p32.load_inputs(ex["points"], ex["depth"], ex["ctx"])
    p32.capture()

    def f(i):
        b = dev_b[i % ring]
        p32.load_inputs(b["points"], b["depth"], b["ctx"])
        p32.replay()

    ms = event_ms(torch, lambda: [f(i) for i in range(steps)], 1, warm=1) / steps
    pipe = frontend.HostPipeline(m32, tab, dev, depth=2, batch=B, max_points=max_pts,
                                 example=(ex["points"], ex["depth"], ex["ctx"]))
    for i in range(2):
        pipe.submit(pin_b[i % ring]["points"], pin_b[i % ring]["depth"], pin_b[i % ring]["ctx"])
    pipe.join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        pipe.submit(pin_b[i % ring]["points"], pin_b[i % ring]["depth"], pin_b[i % ring]["ctx"])
    pipe.join()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / steps
    chk = dev_b[1 % ring]
    p32.load_inputs(chk["points"], chk["depth"], chk["ctx"])
    lid, cam = p32.run()
    torch.cuda.synchronize()
    layers = p32.profile(reps=2)
    out = dict(value=B * 1e3 / ms, unit="frames/s", ms_per_step=ms, e2e_value=B * 1e3 / ms_e2e, steps=steps,
               dtype="f32", gemm_ms=sum(r["ms"] for r in layers),
               note="--precision fp32: spconv_gemm_f32_kernel (FFMA), everything else unchanged; device-resident value "
                    "and end-to-end value from host buffers",
               _lidar=lid.cpu().numpy(), _cam=cam.cpu().numpy())
    del pipe, p32, m32
    torch.cuda.empty_cache()
    return out


def _boundary_and_reference_gpu(torch, stages, tables, dev_b, cfg, pk, m_vox, dev):
    """The reference's own boundary forms (bev_pool_ext.bev_pool_forward / _backward, hard_voxelize): what a reference
    build gets by swapping only the pybind modules -- ours, and the REFERENCE'S OWN CUDA kernels
    (oracle/_ref, compiled unmodified for sm_100a) on the same inputs on the same GPU: the kernels to beat."""
    try:
        from bevfusion_3d_object_detection_b200.ops.bev_pool import bev_pool_ext
        from bevfusion_3d_object_detection_b200.ops.voxel import voxel_layer

        C = cfg["c_ctx"]
        nk = tables.nk
        xs = torch.randn((nk, C), device=dev)
        cellv = torch.repeat_interleave(tables.interval_cell.long(), torch.diff(tables.interval_starts.long()))
        geom4 = torch.stack([(cellv // 360) % 360, cellv % 360, torch.zeros_like(cellv), torch.zeros_like(cellv)],
                            1).int().contiguous()
        starts = tables.interval_starts[:-1].contiguous()
        lengths = torch.diff(tables.interval_starts).int().contiguous()
        og = torch.randn((1, 1, 360, 360, C), device=dev)
        ms_bf = graph_ms(torch, lambda i: bev_pool_ext.bev_pool_forward(xs, geom4, lengths, starts, 1, 1, 360, 360), 4, 3)
        ms_bb = graph_ms(torch, lambda i: bev_pool_ext.bev_pool_backward(og, geom4, lengths, starts, 1, 1, 360, 360), 4, 3)
        by_f = 4 * C * nk + 4 * C * 360 * 360 + 8 * tables.n_intervals
        by_b = 4 * C * tables.n_intervals + 4 * C * nk
        pts0 = dev_b[0]["points"][0]
        c_pts = cfg["dims"]
        cap = cfg["max_voxels"][1]
        vbuf = torch.empty((cap, 10, c_pts), device=dev)
        cbuf = torch.empty((cap, 3), dtype=torch.int32, device=dev)
        nbuf = torch.empty((cap,), dtype=torch.int32, device=dev)
        ms_hv = graph_ms(torch, lambda i: voxel_layer.hard_voxelize_async(pts0, vbuf, cbuf, nbuf, cfg["voxel"], NUS_RANGE,
                                                                         10, cap, True), 4, 3)
        by_hv = 4 * c_pts * int(pts0.shape[0]) + m_vox * (10 * c_pts * 4 + 12 + 4)
        bf = dict(
            bev_pool_forward=dict(ms=ms_bf, bytes=by_f, gbs=by_f / ms_bf / 1e6, frac=by_f / ms_bf / 1e6 / pk["hbm"]),
            bev_pool_backward=dict(ms=ms_bb, bytes=by_b, gbs=by_b / ms_bb / 1e6, frac=by_b / ms_bb / 1e6 / pk["hbm"]),
            hard_voxelize=dict(ms=ms_hv, bytes=by_hv, gbs=by_hv / ms_hv / 1e6, frac=by_hv / ms_hv / 1e6 / pk["hbm"]),
            timing="CUDA graph of 4 calls each (inputs L2-warm: one frame)")
        # the reference's kernels, same inputs, same GPU (event-timed eager calls: their wrappers allocate and sync)
        try:
            from oracle import build_ref

            rp = build_ref.load_ref("ref_bev_pool_cuda")
            rv = build_ref.load_ref("ref_voxel_cuda")
            if rp is not None:
                bf["bev_pool_forward"]["reference_gpu_ms"] = event_ms(
                    torch, lambda: rp.bev_pool_forward(xs, geom4, lengths, starts, 1, 1, 360, 360), 5)
                bf["bev_pool_backward"]["reference_gpu_ms"] = event_ms(
                    torch, lambda: rp.bev_pool_backward(og, geom4, lengths, starts, 1, 1, 360, 360), 5)
                bf["bev_pool_forward"]["ms_eager"] = event_ms(
                    torch, lambda: bev_pool_ext.bev_pool_forward(xs, geom4, lengths, starts, 1, 1, 360, 360), 5)
                bf["bev_pool_backward"]["ms_eager"] = event_ms(
                    torch, lambda: bev_pool_ext.bev_pool_backward(og, geom4, lengths, starts, 1, 1, 360, 360), 5)
            if rv is not None:
                def ref_hv():
                    vbuf.zero_(); cbuf.zero_(); nbuf.zero_()     # the caller's zero-fill (voxelize.py:51-53)
                    return rv.hard_voxelize(pts0, vbuf, cbuf, nbuf, [float(v) for v in cfg["voxel"]], NUS_RANGE, 10,
                                            cap, 3, True)

                def our_hv():
                    vbuf.zero_(); cbuf.zero_(); nbuf.zero_()
                    return voxel_layer.hard_voxelize(pts0, vbuf, cbuf, nbuf, cfg["voxel"], NUS_RANGE, 10, cap, 3, True)

                bf["hard_voxelize"]["reference_gpu_ms"] = event_ms(torch, ref_hv, 2, warm=1)
                bf["hard_voxelize"]["ms_eager_sync_form"] = event_ms(torch, our_hv, 5)
                bf["reference_gpu"] = ("the reference's own bev_pool_cuda.cu / voxelization_cuda.cu (deterministic "
                                       "hard_voxelize_gpu: O(N^2) scan) compiled unmodified for sm_100a, same inputs")
        except Exception as exc:
            print(f"[bench] reference-GPU timing unavailable: {exc}", file=sys.stderr)
        stages["boundary_forms"] = bf
    except Exception as exc:
        print(f"[bench] boundary-form timing unavailable: {exc}", file=sys.stderr)


def _depth_prep_stage(torch, stages, dev_b, cfg, pk, dev):
    try:
        from bevfusion_3d_object_detection_b200 import ops as _ops
        from bevfusion_3d_object_detection_b200 import synthetic

        IMAGE, FEAT, N_CAMS = cfg["image"], cfg["feat"], cfg["n_cams"]
        D_BINS = n_depth_bins(cfg)
        rig_np = synthetic.camera_rig(n_cams=N_CAMS, image_size=IMAGE)
        l2i, iaug, laug = (torch.from_numpy(a).to(dev) for a in synthetic.camera_matrices(rig_np))
        linv = torch.inverse(laug)
        pts0 = dev_b[0]["points"][0]
        dimg = torch.empty((1, N_CAMS, 1) + tuple(IMAGE), device=dev)
        ms_di = graph_ms(torch, lambda i: _ops.lidar_depth_image([pts0], l2i, iaug, laug, IMAGE, linv, out=dimg), 4, 3)
        ms_dh = graph_ms(torch, lambda i: _ops.depth_histogram(dimg, FEAT, cfg["dbound"]), 4, 3)
        px = N_CAMS * IMAGE[0] * IMAGE[1]
        by_di = 4 * cfg["dims"] * int(pts0.shape[0]) + 4 * px
        by_dh = 4 * px + 2 * 4 * N_CAMS * FEAT[0] * FEAT[1] * D_BINS
        stages["depth_prep"] = dict(
            lidar_depth_image=dict(ms=ms_di, bytes=by_di, gbs=by_di / ms_di / 1e6, frac=by_di / ms_di / 1e6 / pk["hbm"]),
            depth_histogram=dict(ms=ms_dh, bytes=by_dh, gbs=by_dh / ms_dh / 1e6, frac=by_dh / ms_dh / 1e6 / pk["hbm"]),
            timing="CUDA graph of 4 calls each; not part of the timed frame step")
    except Exception as exc:
        print(f"[bench] depth-prep timing unavailable: {exc}", file=sys.stderr)


def _tables_stage(torch, stages, model, rig, tables):
    try:
        vt = model.view_transform
        geom = vt.get_geometry(**rig)

        def wall_ms(fn, reps):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3 / reps

        ms_tb_dev = wall_ms(lambda: vt.build_tables(geom, device_build=True), 10)
        ms_tb_aux = wall_ms(lambda: vt.build_tables(geom, device_build=False), 3)
        vt._tables = tables
        stages["pool_tables"] = dict(device_build_ms=ms_tb_dev, bev_pool_aux_route_ms=ms_tb_aux,
                                     points=int(geom.numel() // 3),
                                     timing="host wall clock per build, including its size read-back")
    except Exception as exc:
        print(f"[bench] table-build timing unavailable: {exc}", file=sys.stderr)


# ---------------------------------------------------------------------------------------------------------------
# B200 arm: training (configs[2])
# ---------------------------------------------------------------------------------------------------------------
def _train_measure(torch, dist, cfg, precision, dev, rank, world, steps, warmup, e2e=True):
    from bevfusion_3d_object_detection_b200 import _lib, parallel, synthetic
    from bevfusion_3d_object_detection_b200.spconv import functional as Fsp
    from bevfusion_3d_object_detection_b200.training import TrainStep

    L = _lib.lib()
    B = cfg["batch"]
    model = build_model(cfg, precision, dev)
    rig = {k: torch.from_numpy(v).to(dev) for k, v in synthetic.camera_rig(cfg["n_cams"], cfg["image"], B).items()}
    tables = model.set_calibration(rig)
    ring = 2
    batches = batches_of(make_frames(cfg, ring * B, seed0=100), B)     # 2 x 4 frames x 19.8 MB = 158 MB > L2
    dev_b = [dict(points=[torch.from_numpy(p).to(dev) for p in b["points"]], depth=torch.from_numpy(b["depth"]).to(dev),
                  ctx=torch.from_numpy(b["ctx"]).to(dev)) for b in batches]
    step = TrainStep(model, tables, lr=1e-5)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(i):
        b = dev_b[i % ring]
        return step(b["points"], b["depth"], b["ctx"])

    def timed(fn, k, w):
        for i in range(w):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = L.bevf_launch_count()
        e0.record()
        for i in range(k):
            fn(w + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = L.bevf_launch_count() - n0
        _, ms, _ = parallel.job_throughput(k, ms, device=dev)
        return ms, launches

    ms, launches = timed(run, steps, warmup)
    out = dict(ms=ms, launches=int(launches), batch=B)
    # the same step with the collective switched off: what the all-reduce costs when it is NOT hidden shows here
    ms_nored = None
    if world > 1:
        step.reducer.active = False
        ms_nored, _ = timed(run, steps, 1)
        step.reducer.active = True
    # forward only (train-mode modules, no autograd graph kept)
    def fwd_only(i):
        b = dev_b[i % ring]
        with torch.no_grad():
            f, c, _ = model.voxelize(b["points"])
            model.pts_middle_encoder(f, c, B)
            model.view_transform.pool_fused(b["depth"], b["ctx"], tables)

    ms_fwd, _ = timed(fwd_only, steps, 1)
    # end to end from pinned host buffers: H2D of the batch, step, loss read back
    ms_e2e = None
    h2d = 0
    if e2e:
        pin_b = [dict(points=[torch.from_numpy(p).pin_memory() for p in b["points"]],
                      depth=torch.from_numpy(b["depth"]).pin_memory(), ctx=torch.from_numpy(b["ctx"]).pin_memory())
                 for b in batches]
        h2d = sum(int(p.numel() * 4) for p in pin_b[0]["points"]) + int(pin_b[0]["depth"].numel() * 4 +
                                                                        pin_b[0]["ctx"].numel() * 4)
        losses = []

        def run_e2e(i):
            b = pin_b[i % ring]
            pts = [p.to(dev, non_blocking=True) for p in b["points"]]
            d, c = b["depth"].to(dev, non_blocking=True), b["ctx"].to(dev, non_blocking=True)
            losses.append(float(step(pts, d, c).item()))       # the step's result comes back to the host

        ms_e2e, _ = timed(run_e2e, steps, 1)
        assert all(np.isfinite(v) for v in losses)
    # GEMM-family accounting of one step (CUDA events around every conv forward / dgrad / wgrad launch)
    Fsp.GEMM_TIMING = []
    run(0)
    torch.cuda.synchronize()
    fam = {}
    for ev in Fsp.GEMM_TIMING:
        a, b_, fl = ev[0], ev[1], ev[2]
        tag = ev[3] if len(ev) > 3 else "forward"
        d = fam.setdefault(tag, dict(ms=0.0, flops=0.0, launches=0))
        d["ms"] += a.elapsed_time(b_)
        d["flops"] += fl
        d["launches"] += 1
    Fsp.GEMM_TIMING = None
    pk = peaks()
    for d in fam.values():
        d["tflops"] = d["flops"] / max(d["ms"], 1e-9) / 1e9
        d["frac"] = d["tflops"] / pk["tc"]
    n_params = sum(p.numel() for p in step.params)
    fr0 = batches[0]
    frame_bytes = (sum(int(p.nbytes) for p in fr0["points"]) + int(fr0["depth"].nbytes) + int(fr0["ctx"].nbytes)) // B
    out.update(ms_no_allreduce=ms_nored, ms_forward_only=ms_fwd, ms_e2e=ms_e2e, h2d=h2d, families=fam, frame_bytes=frame_bytes,
               n_params=n_params, buckets=len(step.reducer.buckets), allreduce_launched=step.reducer.launched)
    out["summary"] = dict(
        frames_per_s=world * B * steps / (ms / 1e3), ms_per_step=ms / steps, batch_per_gpu=B, n_gpus=world,
        ms_per_step_forward_only=ms_fwd / steps, fwd_bwd_over_fwd=ms / ms_fwd,
        ms_per_step_without_allreduce=(ms_nored / steps) if ms_nored else None,
        e2e_frames_per_s=(world * B * steps / (ms_e2e / 1e3)) if ms_e2e else None,
        gemm_families=fam, grad_buckets=len(step.reducer.buckets), grad_bytes=4 * n_params,
        timing="eager module path (train-mode BatchNorm1d, tcgen05 forward / data gradient / weight gradient), batch %d "
               "per GPU, CUDA events, max over ranks; all-reduce %s" % (B, "active" if world > 1 else "not needed at N=1"))
    del step, model
    torch.cuda.empty_cache()
    return out


def run_train(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS["train"]
    B = cfg["batch"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    r = _train_measure(torch, dist, cfg, args.precision, dev, rank, world, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    fam = r["families"]
    tot_ms = sum(d["ms"] for d in fam.values())
    tot_fl = sum(d["flops"] for d in fam.values())
    roof = dict(kernel="sparse-conv tensor-core kernels of one step: forward / data gradient (spconv_ts_kernel) + weight "
                       "gradient (spconv_wgrad_tc_kernel), aggregate", bound="tensor",
                achieved=tot_fl / max(tot_ms, 1e-9) / 1e9, peak=pk["tc"], unit="TFLOP/s",
                frac=tot_fl / max(tot_ms, 1e-9) / 1e9 / pk["tc"], traffic=None, peak_source=pk["src"], families=fam)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        import torch as _t

        _t.set_num_threads(os.cpu_count() or 1)
        cpu = CpuFrontEnd(cfg)
        b1 = batches_of(make_frames(cfg, 1, seed0=100), 1)[0]
        t0 = time.perf_counter()
        cpu.train_step(b1)
        dt = time.perf_counter() - t0
        cpu_baseline = dict(value=1.0 / dt, unit="frames/s", cores=cpu.cores, kind=cpu.kind,
                            sample="1 frame forward + backward (torch autograd through the formulation), no warm-up; "
                                   + cpu.describe())
    s = r["summary"]
    line = dict(metric=METRIC, value=s["frames_per_s"], unit="frames/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=s["ms_per_step"], higher_is_better=True, scaling="weak",
                vs_baseline=None,
                dtype=("bf16 sparse conv operands (fp32 accumulate, fp32 master weights and gradients) + fp32 "
                       "voxelize/bev_pool/BatchNorm" if args.precision == "bf16" else "f32"),
                data="synthetic",
                config=arm_config(cfg, args, r["frame_bytes"]),
                e2e=dict(value=s["e2e_frames_per_s"], unit="frames/s", h2d_bytes_per_step=r["h2d"], d2h_bytes_per_step=4),
                gpu_launches=r["launches"], roofline=roof, stages=dict(training=s), cpu_baseline=cpu_baseline,
                clocks=clocks)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("BEVFRONT_BENCH_CONFIG", "infer"), choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("BEVFRONT_BENCH_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-parity block of the default line")
    ap.add_argument("--no-train-stage", action="store_true", help="skip stages.training of the default line")
    ap.add_argument("--repeats", type=int, default=3, help="timed regions of K steps each (median reported beside the first)")
    ap.add_argument("--compact-output", action="store_true",
                    help="e2e: copy the BEV maps to the host as bf16 instead of fp32 (opt-in, lossy; = --output bf16)")
    ap.add_argument("--output", default=None, choices=["dense", "rows", "bf16"],
                    help="form in which e2e returns the BEV maps to the host: dense fp32 maps (default, the reference's "
                         "tensors), lossless rows (active rows + coordinates, 2.4x fewer bytes), or bf16 dense maps")
    ap.add_argument("--inflight", type=int, default=3, help="batches in flight per GPU: plans of the host pipeline")
    ap.add_argument("--inflight-device", type=int, default=2,
                    help="batches in flight for the device-resident measurement (<= --inflight; measured best at 2)")
    args = ap.parse_args()
    if args.output is None:
        args.output = "bf16" if args.compact_output else "dense"
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config == "train":
        run_train(args, rank, world, local_rank)
    else:
        run_frontend(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
