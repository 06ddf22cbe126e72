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
