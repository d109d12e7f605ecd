"""Frame-parallel sharding of the front end across GPUs.

Every operator on the path is per-sample (voxelization loops over samples, bev_pool ranks carry the batch index,
sparse conv hashes on (b, x, y, z)), so the multi-GPU form is: one process per GPU, each owning a disjoint subset of
the frames, NO data-path collective.  The only collective the reference has here is DDP's gradient all-reduce of the
sparse encoder's parameters in training, which stays with torch.distributed (NCCL).  What is left for this module is
the bookkeeping: which frames a rank owns, and the job-level throughput (units of all ranks / slowest rank's time).
"""
import torch
import torch.distributed as dist


def shard_frames(n_frames, rank, world_size):
    """Indices of the frames rank `rank` processes: a strided split, so a stream of frames arriving in order is
    spread evenly and every frame has exactly one owner."""
    assert 0 <= rank < world_size
    return list(range(rank, n_frames, world_size))


def job_throughput(local_units, local_ms, device=None):
    """(total units over all ranks, max time over ranks [ms], units per second for the whole job).
    Works with any initialised backend (nccl on GPUs, gloo on CPU); without a process group it is the local value."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_units, local_ms, local_units / (local_ms / 1e3)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()), float(t.item()), float(u.item()) / (float(t.item()) / 1e3)
