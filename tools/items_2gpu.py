"""Item-sharded evaluation through the public API on 2+ GPUs at a c5-like per-rank size (NCCL all-gather included):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/items_2gpu.py [users=131072] [items_per_rank=1250000]

Every rank holds `items_per_rank` item rows (d=128) and the column partition of the train CSR;
`RankingEvaluator.from_csr(..., shard="items").evaluate(model)` = per-shard top-100 -> all-gather of the rank keys ->
merge + metrics of the rank's slice -> all-reduce of the sums.  The same users are then evaluated user-sharded
(every rank needs the whole item table for that) and the two reports must agree.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402
from skrec_b200 import RankingEvaluator, dist  # noqa: E402

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
U = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
I = (int(sys.argv[2]) if len(sys.argv) > 2 else 1_250_000) * world
d, K = 128, 100
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
td.init_process_group("nccl", device_id=dev)

# the same tables on every rank (same seed); a rank only keeps what its mode needs
g = torch.Generator(device=dev).manual_seed(2026)
uv = torch.randn((U, d), generator=g, device=dev) * 0.1
iv = torch.randn((I, d), generator=g, device=dev) * 0.1
rng = np.random.default_rng(2026)
n_tr, n_te = 12, 10
tr = (np.arange(U + 1, dtype=np.int64) * n_tr, rng.integers(0, I, size=U * n_tr, dtype=np.int32))
# test items: half random, half from a 4,096-item block the user scores highly in, so that the metrics are not ~0
blk = iv[:4096]
top = torch.topk(uv @ blk.T, n_te // 2, dim=1).indices.cpu().numpy().astype(np.int32)
te_idx = np.concatenate([top, rng.integers(4096, I, size=(U, n_te - n_te // 2), dtype=np.int32)], axis=1).ravel()
te = (np.arange(U + 1, dtype=np.int64) * n_te, te_idx)


class Model(object):
    def predict(self, users):
        raise NotImplementedError

    def eval_embeddings(self, users, item_shard=None):
        rows = uv if len(users) == U else uv[torch.as_tensor(np.asarray(users, dtype=np.int64), device=dev)]
        if item_shard is None:
            return rows, iv, None
        lo, hi = dist.shard_range(I, item_shard[0], item_shard[1])
        return rows, iv[lo:hi], None, I


def run(shard, reps):
    ev = RankingEvaluator.from_csr(tr, te, metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], top_k=[K], device=local, shard=shard, shard_users=True)
    model, times = Model(), []
    for r in range(reps + 1):  # first call: CSR upload, workspace, work plan
        torch.cuda.synchronize()
        td.barrier()
        t0 = time.perf_counter()
        rep = ev.evaluate(model)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        if r > 0:
            times.append(float(t.item()))
    return np.array(list(rep.values()), np.float64), float(np.median(times)), ev.last_stats["path"]


v_items, t_items, p_items = run("items", 3)
v_users, t_users, p_users = run("users", 2)
if rank == 0:
    print(json.dumps({
        "workload": "%d users x %d items (%d per rank), d=%d, top-%d, %d ranks" % (U, I, I // world, d, K, world),
        "items": {"path": p_items, "ms_per_evaluate": t_items * 1e3, "users_per_s": U / t_items,
                  "allgather_bytes_received_per_rank": U * K * 8 * world},
        "users": {"path": p_users, "ms_per_evaluate": t_users * 1e3, "users_per_s": U / t_users},
        "max_abs_metric_diff": float(np.max(np.abs(v_items - v_users))), "NDCG@100": float(v_items[3]),
        "timing": "wall clock around evaluate() with a barrier before and a device synchronisation after, max over ranks, median"}))
td.destroy_process_group()
