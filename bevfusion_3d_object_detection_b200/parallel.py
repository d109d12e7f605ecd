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


class GradBucketReducer:
    """The training split's one collective (BASELINE configs[2]): average the sparse encoder's parameter gradients
    over the data-parallel ranks, overlapped with the backward pass.  The reference gets this from mmengine's
    MMDistributedDataParallel (configs/_base_/default_runtime.py:14, tools/dist_train.sh:10-19 -> torch DDP); here it is
    a ~100-line reducer with the same semantics and no per-step bookkeeping on the host:

      * parameters are grouped, in REVERSE registration order (the order backward produces gradients: conv_out and the
        128-channel stage, 65 % of the bytes, come first), into flat fp32 buckets of about `bucket_bytes`;
      * every `param.grad` IS a view into its bucket (gradients are accumulated in place by autograd: no flatten copy);
      * a post-accumulate hook counts a bucket's ready parameters and launches ONE asynchronous all-reduce (NCCL: AVG on
        NCCL's own stream, over NVLink / NVSwitch) as soon as the bucket is complete, while backward keeps running;
      * `finish()` makes the current stream wait for the collectives.

    With world_size 1 (or no process group) nothing is launched; buckets and views are still used."""

    def __init__(self, params, bucket_bytes=2 << 20, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.world = dist.get_world_size(process_group) if self.active else 1
        self.avg_op = None
        if self.active:
            self.avg_op = dist.ReduceOp.AVG if dist.get_backend(process_group) == "nccl" else dist.ReduceOp.SUM
        self.buckets = []            # dict(flat, params, pending)
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
        if cur:
            self._close(cur)
        self._bucket_of = {}
        self._hooks = []
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._bucket_of[p] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._work = []
        self.launched = 0

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=torch.float32, device=plist[0].device)
        off = 0
        for p in plist:
            assert p.dtype == torch.float32, "gradient buckets are fp32"
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.buckets.append(dict(flat=flat, params=list(plist), pending=len(plist)))

    def prepare(self):
        """Before every backward: zero the buckets (gradients accumulate in place), re-arm the counters."""
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            for p in b["params"]:      # an optimizer / zero_grad(set_to_none) may have dropped the views
                if p.grad is None or p.grad.data_ptr() < b["flat"].data_ptr() or \
                        p.grad.data_ptr() >= b["flat"].data_ptr() + b["flat"].numel() * 4:
                    self._rebind(b)
                    break
        self._work = []

    def _rebind(self, b):
        off = 0
        for p in b["params"]:
            p.grad = b["flat"][off:off + p.numel()].view_as(p)
            off += p.numel()

    def _on_grad(self, p):
        b = self.buckets[self._bucket_of[p]]
        b["pending"] -= 1
        if b["pending"] == 0 and self.active:
            self._work.append((dist.all_reduce(b["flat"], op=self.avg_op, group=self.group, async_op=True), b))
            self.launched += 1

    def finish(self):
        """After backward: the current stream waits for every bucket's all-reduce; gloo (tests) divides by the world."""
        for w, b in self._work:
            w.wait()
            if self.avg_op == dist.ReduceOp.SUM:
                b["flat"].div_(self.world)
        self._work = []

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    @property
    def flats(self):
        return [b["flat"] for b in self.buckets]
