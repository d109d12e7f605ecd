"""world_size-2 gloo test of the frame-parallel bookkeeping (no GPU): every frame has exactly one owner, and the
job-level throughput is total units over the slowest rank's time."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bevfusion_3d_object_detection_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_frames(n_frames, rank, world)
    # "process" the frames: a checksum per frame so the union can be verified, and a rank-dependent time
    owned = torch.zeros(n_frames, dtype=torch.int64)
    owned[mine] = 1
    dist.all_reduce(owned)
    local_ms = 10.0 * (rank + 1)
    units, ms, per_s = parallel.job_throughput(len(mine), local_ms)
    torch.save(dict(owned=owned, units=units, ms=ms, per_s=per_s, mine=mine), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_sharding_and_throughput(tmp_path):
    world, n_frames = 2, 17
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for r in res:
        assert bool((r["owned"] == 1).all())            # every frame processed exactly once
        assert r["units"] == n_frames
        assert r["ms"] == 20.0                            # slowest rank
        assert abs(r["per_s"] - n_frames / 0.020) < 1e-6
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(n_frames))
    assert len(res[0]["mine"]) - len(res[1]["mine"]) in (0, 1)


def test_single_process_is_identity():
    assert parallel.shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]
    u, ms, per_s = parallel.job_throughput(8, 4.0)
    assert (u, ms, per_s) == (8, 4.0, 2000.0)
