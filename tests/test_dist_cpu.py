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
