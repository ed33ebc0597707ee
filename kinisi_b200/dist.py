"""Multi-GPU plumbing: one process per GPU, atoms sharded across ranks, per-dt moment sums and the
per-frame framework / collective sums joined by a sum all-reduce (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  Everything is a no-op when torch.distributed is not initialised."""
import torch
import torch.distributed as td


def active():
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def world_size():
    return td.get_world_size() if active() else 1


def rank():
    return td.get_rank() if active() else 0


def allreduce_sum_(t):
    """In-place sum all-reduce of a tensor (float64 moment sums / frame sums)."""
    if active():
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def allreduce_sum_scalar(x, device=None):
    """Sum of a host scalar over ranks (atom counts, charge sums)."""
    if not active():
        return x
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else 'cpu')
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return float(t.item())


def shard_bounds(n, r=None, w=None):
    """Contiguous block of `n` atoms owned by rank r of w."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)
