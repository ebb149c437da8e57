"""Data-parallel plumbing for the hot path: the work shards by image pair (SURVEY.md section 8e:
every kernel is independent per batch item, correlation_cuda_kernel.cu:52; nothing on the path
communicates), so multi-GPU = one process per GPU, a contiguous split of the pairs, and no
data-path collective.  torch.distributed is used only for the barrier and for reducing the timing
(max over ranks), on whatever backend the process group has (nccl on GPUs, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [start, stop) of `total` image pairs owned by `rank`; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world):
    """Slices every tensor of a (f1, f2, flow, ...) tuple along dim 0 for this rank."""
    start, stop = shard_range(tensors[0].shape[0], rank, world)
    return tuple(t[start:stop] for t in tensors)


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def barrier():
    if _active():
        dist.barrier()


def max_over_ranks(value, device=None):
    """max over ranks of a python float (the job's step time is the slowest rank's)."""
    if not _active():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    if not _active():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank, elapsed_ms_this_rank, device=None):
    """Whole-job units/s: all ranks' units divided by the slowest rank's time."""
    total_units = sum_over_ranks(units_this_rank, device)
    worst_ms = max_over_ranks(elapsed_ms_this_rank, device)
    return total_units / (worst_ms * 1e-3), worst_ms
