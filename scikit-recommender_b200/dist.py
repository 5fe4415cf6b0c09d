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


_comms = {}  # (group id, device index) -> _native.Comm, or False when the one-shot path is not available there


def nvlink_comm(device, group=None):
    """The one-shot NVLink all-reduce communicator of (device, group), set up on first use: every rank exports its inbox
    with cudaIpc, the 64-byte handles travel through `all_gather_object`, every rank maps its peers' inboxes.  COLLECTIVE
    on first use.  -> _native.Comm, or None when the group is not NCCL, spans more than 16 ranks, or the peers' memory
    cannot be mapped (different nodes, no peer access) -- the caller then uses `torch.distributed.all_reduce`."""
    import os
    import torch.distributed as td
    key = (id(group) if group is not None else 0, int(device))
    c = _comms.get(key)
    if c is not None:
        return c or None
    c = False
    rank, world = td.get_rank(group), td.get_world_size(group)
    usable = td.get_backend(group) == "nccl" and 1 < world <= 16 and os.environ.get("SKR_NVLINK_ALLREDUCE", "1") != "0"
    if usable:
        from . import _native
        comm, handle = None, b""
        try:
            comm = _native.Comm(device, rank, world)
            handle = comm.handle()
        except _native.NativeError:
            comm = None
        handles = [None] * world
        td.all_gather_object(handles, handle, group=group)
        ok = comm is not None and all(isinstance(h, (bytes, bytearray)) and len(h) == 64 for h in handles)
        if ok:
            try:
                comm.connect([bytes(h) for h in handles])
            except _native.NativeError:
                ok = False
        flags = [None] * world
        td.all_gather_object(flags, bool(ok), group=group)  # all or nobody: a mixed choice would deadlock
        if all(flags):
            c = comm
        elif comm is not None:
            comm.close()
    _comms[key] = c
    return c or None


def allreduce_sums(sums, group=None):
    """In-place SUM all-reduce of the float64 [n_metrics*max_top + 1] vector (sums | user count): on CUDA tensors of
    an NCCL group inside one node, one kernel over NVLink peer memory (`nvlink_comm`), ranks added in rank order --
    identical bits on every rank; otherwise `torch.distributed.all_reduce` (NCCL across nodes, gloo in the CPU tests)."""
    import torch.distributed as td
    if sums.is_cuda and sums.numel() <= 4096:
        comm = nvlink_comm(sums.device.index, group)
        if comm is not None:
            comm.allreduce(sums)
            return sums
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


def gather_host_table(host, dev, rank, world, group=None):
    """A float32 table [n, d] (or vector [n]) that every rank holds in HOST memory (the replicated item table of
    user-sharded evaluation) -> the whole table on this rank's device.  Instead of `world` identical uploads competing
    for the host's memory and PCIe links, each rank uploads only rows [rank * per, (rank + 1) * per), per = ceil(n /
    world), and the slices are all-gathered in place (NCCL over NVLink / NVSwitch on GPUs, gloo on CPU tensors in the
    tests).  COLLECTIVE.  The all-gather runs asynchronously: -> (table view [n, ...], work handle or None); the caller
    issues whatever else it has to upload, then `work.wait()`s (the current stream waits, not the host)."""
    import torch
    import torch.distributed as td
    if isinstance(host, np.ndarray):
        host = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32))
    host = host.detach()
    if host.dtype != torch.float32:
        host = host.float()
    if not host.is_contiguous():
        host = host.contiguous()
    n = int(host.shape[0])
    per = -(-n // world)
    full = torch.empty((per * world,) + tuple(host.shape[1:]), dtype=torch.float32, device=dev)
    lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
    mine = full[rank * per:(rank + 1) * per]
    if hi > lo:
        mine[:hi - lo].copy_(host[lo:hi])  # blocking for the host: the caller may change its table as soon as we return
    try:
        work = td.all_gather_into_tensor(full, mine, group=group, async_op=True)
    except (RuntimeError, NotImplementedError):  # backends without the fused form
        parts = [full[r * per:(r + 1) * per] for r in range(world)]
        work = td.all_gather(parts, mine.clone(), group=group, async_op=True)
    return full[:n], work


def finalize_means(col_sums, n_users):
    """float64 column sums / user count, rounded once to float32 (what MetricReport carries)."""
    n = float(n_users)
    if n <= 0:
        return np.zeros(len(col_sums), dtype=np.float32)
    return (np.asarray(col_sums, dtype=np.float64) / n).astype(np.float32)
