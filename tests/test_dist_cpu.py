"""User-sharded evaluation, host logic only (gloo, world_size 2, CPU): slice arithmetic, the
all-reduce of [column sums | user count] and the final mean equal the single-process result.
The per-rank sums are produced by the oracle here -- the CUDA kernels are covered by -m gpu."""
import os
import socket

import numpy as np
import torch
import torch.distributed as td
import torch.multiprocessing as mp

import oracle
from skrec_b200 import dist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 128, 29858, 1000003):
        for world in (1, 2, 3, 4, 8):
            cuts = [dist.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_rank_world_without_init():
    assert dist.rank_world() == (0, 1)


def _workload():
    g = np.random.default_rng(123)
    U, I, K = 301, 400, 10
    s = np.stack([g.permutation(I).astype(np.float32) / I for _ in range(U)])
    sizes = g.integers(1, 9, size=U)
    indptr = np.zeros(U + 1, np.int64)
    np.cumsum(sizes, out=indptr[1:])
    indices = np.concatenate([g.choice(I, size=int(n), replace=False) for n in sizes]).astype(np.int32)
    return s, indptr, indices, [1, 2, 3, 4, 5], K


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    s, indptr, indices, metric, K = _workload()
    lo, hi = dist.shard_range(s.shape[0], *dist.rank_world())
    local_ptr = indptr[lo:hi + 1] - indptr[lo]
    local_idx = indices[indptr[lo]:indptr[hi]]
    per = oracle.eval_scores(s[lo:hi], local_ptr, local_idx, metric, K)
    vec = torch.zeros(len(metric) * K + 1, dtype=torch.float64)
    vec[:-1] = torch.from_numpy(oracle.sums_f64(per))
    vec[-1] = hi - lo
    dist.allreduce_sums(vec)
    means = dist.finalize_means(vec[:-1].numpy(), vec[-1].item())
    if rank == 0:
        np.save(out_path, means)
    td.destroy_process_group()


def test_two_rank_gloo_equals_single_process(tmp_path):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "means.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    s, indptr, indices, metric, K = _workload()
    per = oracle.eval_scores(s, indptr, indices, metric, K)
    single = dist.finalize_means(oracle.sums_f64(per), s.shape[0])
    got = np.load(out)
    assert np.max(np.abs(got - single)) <= 1e-7
    assert np.max(np.abs(got - oracle.mean_f32(per))) < 1e-5


# ---- item-sharded evaluation, host logic (gloo, world_size 2): gather layout + column partition -------
def _pack_keys(scores, items):
    """rank keys as the kernels build them: ord(score) << 32 | ~item (int64 view)"""
    b = scores.astype(np.float32).view(np.uint32).astype(np.uint64)
    ordv = np.where(b & np.uint64(0x80000000), (~b) & np.uint64(0xffffffff), b | np.uint64(0x80000000))
    return ((ordv << np.uint64(32)) | ((~items.astype(np.uint64)) & np.uint64(0xffffffff))).view(np.int64)


def _items_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    s, indptr, indices, metric, K = _workload()
    U, I = s.shape
    lo, hi = dist.shard_range(I, rank, world)
    # this rank's exact top-K over its item range (descending score; scores are tie-free)
    order = np.argsort(-s[:, lo:hi], axis=1, kind="stable")[:, :K]
    keys = _pack_keys(np.take_along_axis(s[:, lo:hi], order, 1), order + lo)
    gathered = dist.allgather_keys(torch.from_numpy(keys)).numpy()
    if rank == 0:
        np.save(out_path, gathered)
    td.destroy_process_group()


def test_two_rank_gloo_item_shard_gather_and_merge(tmp_path):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = str(tmp_path / "keys.npy")
    mp.spawn(_items_worker, args=(2, port, out), nprocs=2, join=True)
    s, indptr, indices, metric, K = _workload()
    g = np.load(out).view(np.uint64)  # [world, U, K], shard order = rank order
    assert g.shape == (2, s.shape[0], K)
    merged = np.sort(np.concatenate([g[0], g[1]], axis=1), axis=1)[:, ::-1][:, :K]  # larger key = ranked earlier
    items = (~merged).astype(np.uint32).astype(np.int64)
    expect = np.argsort(-s, axis=1, kind="stable")[:, :K]
    assert np.array_equal(items, expect)
    for r in range(2):
        lo, hi = dist.shard_range(s.shape[1], r, 2)
        it = (~g[r]).astype(np.uint32)
        assert it.min() >= lo and it.max() < hi


def test_column_partition_of_interactions():
    import importlib
    ev = importlib.import_module("skrec_b200.evaluator")
    d = {3: np.array([5, 1, 9, 7], np.int32), 4: np.array([], np.int32), 8: np.array([2, 8], np.int32)}
    users = [8, 3, 4, 99]
    ptr, idx = ev._dict_to_csr(users, d)
    assert ptr.tolist() == [0, 2, 6, 6, 6] and idx.tolist() == [2, 8, 5, 1, 9, 7]
    ptr, idx = ev._dict_to_csr(users, d, (5, 9))
    assert ptr.tolist() == [0, 1, 3, 3, 3] and idx.tolist() == [3, 0, 2]
    ptr, idx = ev._dict_to_csr(users, d, (100, 200))
    assert ptr.tolist() == [0, 0, 0, 0, 0] and idx.size == 0


# ---- replicated host item table: 1/world uploaded per rank + all-gather (gloo here, NCCL over NVLink on GPUs) -----
def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for n, d in ((7, 5), (1, 3), (64, 4), (0, 4), (9, None)):  # ragged last slice, fewer rows than ranks, even, empty, vector
        g = np.random.default_rng(n)
        table = g.standard_normal((n, d) if d is not None else (n,)).astype(np.float32)  # identical on every rank
        host = table if (n % 2) else torch.from_numpy(table)
        full, work = dist.gather_host_table(host, torch.device("cpu"), rank, world)
        if work is not None:
            work.wait()
        ok = ok and tuple(full.shape) == table.shape and np.array_equal(full.numpy(), table)
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([ok]))
    td.destroy_process_group()


def test_two_rank_gloo_host_table_is_uploaded_in_slices_and_gathered(tmp_path):
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(bool(np.load(str(tmp_path / ("ok%d.npy" % r)))[0]) for r in range(2))
