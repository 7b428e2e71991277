"""Multi-GPU plumbing: the problems of a batch are independent, so the batch is cut into
contiguous blocks, one per rank (one process per GPU), and nothing is exchanged while solving.
torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests) is used only to combine the
per-rank counters and, on request, to gather the solutions."""
import torch
import torch.distributed as dist


def shard_range(n_problems, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: block sizes differ by at most one."""
    base, rem = divmod(int(n_problems), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counters(time_ms, converged, iterations, problems, device="cpu"):
    """(max over ranks of time_ms, sums of the counters).  Works without an initialised group."""
    t = torch.tensor([float(time_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(converged), float(iterations), float(problems)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return t.item(), c[0].item(), c[1].item(), c[2].item()


def gather_solutions(local, n_total):
    """all_gather_into_tensor of per-rank result blocks (equal block sizes are padded by the caller).
    `local` is a [b, ...] tensor; returns the [world * b, ...] tensor cut to n_total rows."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local[:n_total]
    world = dist.get_world_size()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out[:n_total]
