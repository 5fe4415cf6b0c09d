"""ncu target: a few launches of the fused scoring kernel on a bench workload (c2 default; c3a, c3b, c4 = 125,000-user share) (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfgname = sys.argv[3] if len(sys.argv) > 3 else "c2"
if cfgname in ("c4", "c5"):  # one row chunk (131,072 users) of the large configs against the full catalogue, generated on the GPU
    cfg = dict(synth.CONFIGS[cfgname])
    f = 131072.0 / cfg["users"]
    cfg.update(users=131072, nnz_train=int(cfg["nnz_train"] * f), nnz_test=int(cfg["nnz_test"] * f))
    if cfgname == "c5":
        cfg["items"] = cfg["items"] // 8  # one rank's item shard of the 8-way item sharding
    d = synth.make_large(device="cuda", item_seed=cfg["seed"] + 7, **cfg)
    d["user_emb"], d["item_emb"] = d["user_emb"].cpu().numpy(), d["item_emb"].cpu().numpy()
    d["bias"] = None if d["bias"] is None else d["bias"].cpu().numpy()
else:
    d = synth.make_config(cfgname, device="cuda")
    cfg = d["config"]
ctx = _native.Context(0)
ctx.set_option("chunks", int(os.environ.get("SKR_CHUNKS", "0")))  # item-range chunks per user tile (0 = planner)
ctx.set_train_csr(d["train_indptr"], d["train_indices"], d["items"])
ctx.set_test_csr(d["test_indptr"], d["test_indices"], d["items"])
n_use = int(os.environ.get("SKR_USERS", "0")) or d["users"]  # evaluate only the first rows (a rank's share of a strong-scaled run)
ue, ie = torch.from_numpy(d["user_emb"][:n_use]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
b = None if d["bias"] is None else torch.from_numpy(d["bias"]).cuda()
ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
K = max(cfg["top_k"])
sums = torch.zeros(len(ids) * K, dtype=torch.float64, device="cuda")
ms = []
import time
wall = []
for kv in filter(None, os.environ.get("SKR_OPTS", "").split(",")):  # e.g. SKR_OPTS=no_aug=1,retry_min=1
    ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
ab = os.environ.get("SKR_AB")  # A/B an option under the same thermal / power state: blocks of n evaluates, value 0 and 1 alternating
if ab:
    for blk in range(6):
        ctx.set_option(ab, blk & 1)
        ks, ws = [], []
        for _ in range(n):
            sums.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.eval_fused(ue, ie, b, 0, ids, K, precision=prec, sums=sums)
            torch.cuda.synchronize()
            ws.append((time.perf_counter() - t0) * 1e3)
            ks.append(ctx.fused_kernel_ms(0))
        print("%s=%d: evaluate ms %s | kernel ms %s | prepass %.3f" % (ab, blk & 1, " ".join("%.2f" % w for w in ws), " ".join("%.2f" % k for k in ks),
                                                                     ctx.fused_prepass_ms(0)), flush=True)
    sys.exit(0)
for _ in range(n):
    sums.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.eval_fused(ue, ie, b, 0, ids, K, precision=prec, sums=sums)
    torch.cuda.synchronize()
    wall.append((time.perf_counter() - t0) * 1e3)
    ms.append(ctx.fused_kernel_ms(0))
print("%s %s rows %d wall ms:" % (cfgname, prec, n_use), ["%.3f" % m for m in wall], "kernel ms:", ["%.3f" % m for m in ms], "prepass %.3f" % ctx.fused_prepass_ms(0),
      "NDCG@%d=%.6f" % (K, float(sums[-1]) / d["users"]), ctx.fused_stats())
