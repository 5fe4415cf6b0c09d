"""One rank's work per user chunk of the item-sharded config c5 (2M users x 10M items, d=128, top-100, 8 shards),
measured on ONE GPU: the per-shard fused top-K over this rank's 1.25M item rows for a chunk of users
(`skr_topk_fused`), and the merge + metrics of this rank's slice of the chunk from 8 gathered lists
(`skr_eval_merged_topk`).  The all-gather between the two needs the 8 GPUs and is not part of this measurement.

    python tools/c5_shard.py [chunk_users=262144] [repeats=3]

The seven foreign lists are this shard's own list re-labelled with the other shards' id ranges (same scores, other
global ids): the merge does the same work as on real lists, ties across shards included.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, dist, synth  # noqa: E402

cfg = synth.CONFIGS["c5"]
WORLD = 8
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
I, d, K = cfg["items"], cfg["d"], max(cfg["top_k"])
ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
MK = len(ids) * K
rank = 3
ilo, ihi = dist.shard_range(I, rank, WORLD)
n_loc = ihi - ilo

g = torch.Generator(device="cuda").manual_seed(cfg["seed"])
uv = torch.randn((n, d), generator=g, device="cuda") * 0.1
iv = torch.randn((n_loc, d), generator=g, device="cuda") * 0.1
rng = np.random.default_rng(cfg["seed"])
# column partition of the train CSR: ~50 train items per user over 8 shards, shard-local ids; 10 test items, global ids
per_tr, per_te = max(1, cfg["nnz_train"] // cfg["users"] // WORLD), cfg["nnz_test"] // cfg["users"]
tr_ptr = np.arange(n + 1, dtype=np.int64) * per_tr
tr_idx = rng.integers(0, n_loc, size=n * per_tr, dtype=np.int32)
te_ptr = np.arange(n + 1, dtype=np.int64) * per_te
te_idx = rng.integers(0, I, size=n * per_te, dtype=np.int32)

ctx = _native.Context(0)
ctx.set_train_csr(tr_ptr, tr_idx, n_loc)
ctx.set_test_csr(te_ptr, te_idx, I)
keys_all = torch.empty((WORLD, n, K), dtype=torch.int64, device="cuda")
sums = torch.zeros(MK, dtype=torch.float64, device="cuda")
lo, hi = dist.shard_range(n, rank, WORLD)


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


t_topk, t_kernel, t_merge = [], [], []
for r in range(reps + 1):  # first round: warm-up (workspace allocation, work plan)
    for s in range(WORLD):
        slo, _ = dist.shard_range(I, s, WORLD)
        ms = timed(lambda: ctx.topk_fused(uv, iv, None, 0, slo, K, keys_all[s], precision="auto"))
        if s == rank and r > 0:
            t_topk.append(ms)
            t_kernel.append(ctx.fused_kernel_ms(0))
    sums.zero_()
    ms = timed(lambda: ctx.eval_merged_topk(keys_all, lo, hi - lo, lo, ids, K, sums=sums))
    if r > 0:
        t_merge.append(ms)

ka = keys_all[:, :4096].cpu().numpy().view(np.uint64)
items = (~ka).astype(np.uint32)
for s in range(WORLD):
    slo, shi = dist.shard_range(I, s, WORLD)
    assert items[s].min() >= slo and items[s].max() < shi, "per-shard lists carry global ids of their own range"
    assert np.all(ka[s][:, :-1] > ka[s][:, 1:]), "per-shard lists are strictly descending rank keys"
topk, kern, merge = float(np.median(t_topk)), float(np.median(t_kernel)), float(np.median(t_merge))
flops = 2.0 * n * n_loc * d
ag_bytes = n * K * 8
print(json.dumps({
    "workload": "c5 shard: %d-user chunk x %d of %d items (rank %d of %d), d=%d, top-%d" % (n, n_loc, I, rank, WORLD, d, K),
    "path": ctx.last_fused_kernel, "plan": ctx.fused_stats(),
    "topk_fused_ms": topk, "scoring_kernel_ms": kern, "scoring_tflops": flops / (kern * 1e-3) / 1e12,
    "merge_metrics_ms": merge, "merge_rows": hi - lo,
    "allgather_bytes_sent_per_rank": ag_bytes, "allgather_bytes_received_per_rank": ag_bytes * WORLD,
    "users_per_s_8gpu_without_allgather": n / ((topk + merge) * 1e-3),
    "note": "one GPU; the all-gather of the [n, K] uint64 lists between the two calls is not measured here"}))
