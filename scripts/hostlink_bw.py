"""Raw host-link bandwidth of the box with N ranks copying at once: every rank moves, per "frame", one
pinned H2D copy of the front end's input bytes and one pinned D2H copy of its output bytes on two streams -- exactly what
`bench.py`'s end-to-end pipeline moves -- and rank 0 prints the aggregate.  This is the ceiling of `e2e`; bench.py runs the
same measurement (`bench.hostlink_gbs`) inside every run and reports it as `e2e.hostlink`.

    python scripts/hostlink_bw.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/hostlink_bw.py
    ... [--h2d BYTES] [--d2h BYTES]     default: 19 782 792 in, 74 649 600 out (config A, fp32 dense BEV maps)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--h2d", type=int, default=19782792)
ap.add_argument("--d2h", type=int, default=74649600)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
gbs, ms = bench.hostlink_gbs(torch, dev, args.h2d, args.d2h, reps=10)
tot = torch.tensor([gbs], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tot)
if rank == 0:
    print(json.dumps(dict(n_gpus=world, h2d_bytes=args.h2d, d2h_bytes=args.d2h, gbs_all_ranks=float(tot.item()),
                          gbs_rank0=gbs, ms_per_frame_copies_only=ms,
                          ceiling_frames_per_s=world * 1e3 / ms)))
if world > 1:
    dist.destroy_process_group()
