"""User-sharded multi-GPU plumbing (one process per GPU, torch.distributed).

Test users are independent units (evaluate.h:64-70 evaluates each row in isolation; the only
cross-user step is the mean, evaluator.py:208), so the path shards by users with no data-path
collective: every rank evaluates a contiguous slice against a replicated item table and only
`n_metrics * max_top` float64 sums plus one user count are all-reduced (NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
import numpy as np


def rank_world(group=None):
    """(rank, world_size) of `group`, or (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as td
    except ImportError:  # pragma: no cover
        return 0, 1
    if not (td.is_available() and td.is_initialized()):
        return 0, 1
    return td.get_rank(group), td.get_world_size(group)


def shard_range(n, rank, world):
    """Contiguous near-equal slice [lo, hi) of n units for `rank`: every user costs the same
    2*I*d flops, so equal counts balance the ranks.  The first n % world ranks get one more."""
    if world <= 1:
        return 0, n
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def allreduce_sums(sums, group=None):
    """In-place SUM all-reduce of the float64 [n_metrics*max_top + 1] vector (sums | user count)."""
    import torch.distributed as td
    td.all_reduce(sums, op=td.ReduceOp.SUM, group=group)
    return sums


def allgather_keys(keys, group=None):
    """Item-sharded evaluation: every rank contributes its [n, K] per-shard rank keys (int64 view of
    uint64) and receives all of them as [world, n, K], shard order = rank order (NCCL all-gather over
    NVLink on GPUs, gloo on CPU tensors in the tests)."""
    import torch
    import torch.distributed as td
    world = td.get_world_size(group)
    out = torch.empty((world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
    try:
        td.all_gather_into_tensor(out, keys.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):  # backends without the fused form
        td.all_gather([out[r] for r in range(world)], keys.contiguous(), group=group)
    return out


def finalize_means(col_sums, n_users):
    """float64 column sums / user count, rounded once to float32 (what MetricReport carries)."""
    n = float(n_users)
    if n <= 0:
        return np.zeros(len(col_sums), dtype=np.float32)
    return (np.asarray(col_sums, dtype=np.float64) / n).astype(np.float32)
