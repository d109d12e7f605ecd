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


def _reducer_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                       # same parameters on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3))
    red = parallel.GradBucketReducer(net.parameters(), bucket_bytes=600)
    assert len(red.buckets) >= 2
    g = torch.Generator().manual_seed(100 + rank)
    grads = []
    for step in range(2):                      # second step: buckets are re-zeroed, not accumulated across steps
        x = torch.randn(5, 6, generator=g)
        red.prepare()
        net(x).pow(2).sum().backward()
        red.finish()
        grads.append([p.grad.clone() for p in net.parameters()])
        local = torch.autograd.grad(net(x).pow(2).sum(), list(net.parameters()))
        grads.append([t.clone() for t in local])
    torch.save(dict(grads=grads, launched=red.launched, n_buckets=len(red.buckets),
                    views=all(p.grad.data_ptr() >= red.buckets[red._bucket_of[p]]["flat"].data_ptr()
                              for p in net.parameters())), os.path.join(out_dir, f"g{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_bucketed_gradient_average(tmp_path):
    """GradBucketReducer (the training split's all-reduce, SURVEY 8e / configs[2]) on gloo, world_size 2: after
    finish() every rank holds the mean of the ranks' local gradients, one collective per bucket per step."""
    world = 2
    mp.spawn(_reducer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, f"g{r}.pt")) for r in range(world)]
    for step in range(2):
        reduced = [r["grads"][2 * step] for r in res]
        local = [r["grads"][2 * step + 1] for r in res]
        for i in range(len(reduced[0])):
            want = (local[0][i] + local[1][i]) / 2
            assert torch.allclose(reduced[0][i], want, rtol=1e-6, atol=1e-7)
            assert torch.equal(reduced[0][i], reduced[1][i])
    assert res[0]["launched"] == 2 * res[0]["n_buckets"] and res[0]["views"]


def test_reducer_single_process_keeps_local_gradients():
    net = torch.nn.Linear(4, 3)
    red = parallel.GradBucketReducer(net.parameters())
    x = torch.randn(2, 4)
    red.prepare()
    net(x).sum().backward()
    red.finish()
    want = torch.autograd.grad(net(x).sum(), list(net.parameters()))
    for p, w in zip(net.parameters(), want):
        assert torch.allclose(p.grad, w)
    assert red.launched == 0
