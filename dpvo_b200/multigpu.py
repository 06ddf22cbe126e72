"""One process per GPU, one independent video stream per process.

The update hot path has no intra-update exchange step, so multi-GPU inference is pure sharding by
stream: there is NO collective on the data path.  The only communication is the reduction of timing
scalars (max over ranks) and of per-rank unit counts for the whole-job throughput, done with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend, device=None):
    """Initialise the process group from the torchrun environment (no-op for a single process)."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = torch.device(device)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def streams_of_rank(n_streams, rank, world):
    """Stream ids handled by `rank`: round robin, every stream exactly once."""
    return list(range(rank, n_streams, world))


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (identity for one process)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_rate(units_this_rank, seconds_this_rank, device="cpu"):
    """units processed by all ranks / slowest rank's time -- the `value` of bench.py."""
    total = sum_over_ranks(units_this_rank, device)
    tmax, = max_over_ranks([seconds_this_rank], device)
    return total / tmax, tmax


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def finalize():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


class GradReducer:
    """Data-parallel gradient averaging for the training step (train.py:121 in a DDP setting):
    the parameters are packed into buckets of ~`bucket_mb` MB in reverse registration order (the order in which
    backward finishes them); a post-accumulate-grad hook per parameter counts its bucket down and, when the bucket is
    complete, copies its gradients into a flat fp32 buffer and starts an asynchronous all-reduce (NCCL on its own
    stream: it overlaps the rest of the backward pass).  `finish()` waits, divides by the world size and scatters
    the averages back into `.grad`.  With one process it does nothing."""

    def __init__(self, params, bucket_mb=4.0):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.buckets, cur, cur_bytes = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            cur_bytes += p.numel() * 4
            if cur_bytes >= bucket_mb * 2 ** 20:
                self.buckets.append(cur); cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device) for b in self.buckets]
        self.bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self.pending, self.work, self.launched = [], [], []
        self.bytes_per_step = sum(f.numel() * 4 for f in self.flat)
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._ready)

    def begin(self):
        """arm the hooks for the backward pass that follows (earlier, un-armed backward passes only accumulate)"""
        self.pending = [len(b) for b in self.buckets]
        self.work, self.launched = [], [False] * len(self.buckets)
        self.armed = True

    def _launch(self, i):
        torch._foreach_copy_(list(self.flat[i].split([p.numel() for p in self.buckets[i]])),
                             [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.buckets[i]])
        self.work.append((i, dist.all_reduce(self.flat[i], op=dist.ReduceOp.SUM, async_op=True)))
        self.launched[i] = True

    def _ready(self, p):
        if not getattr(self, "armed", False):
            return
        i = self.bucket_of[id(p)]
        self.pending[i] -= 1
        if self.pending[i] == 0:
            self._launch(i)

    def finish(self):
        self.armed = False
        if self.world == 1:
            return
        for i in range(len(self.buckets)):          # parameters that received no gradient this step never fire a hook
            if not self.launched[i]:
                self._launch(i)
        for i, w in self.work:
            w.wait()
            self.flat[i].div_(self.world)
            for p, chunk in zip(self.buckets[i], self.flat[i].split([p.numel() for p in self.buckets[i]])):
                if p.grad is None:
                    p.grad = chunk.view_as(p).clone()
                else:
                    p.grad.copy_(chunk.view_as(p))
