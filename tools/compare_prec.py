"""Whole-evaluate device time of the fused path per precision on several shapes (development aid).
    python tools/compare_prec.py [shape ...]   shapes: c2 c3a c3b big128 (16384 x 262144, d=128, K=100) c4slice (8192 x 1M, d=128)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

SHAPES = {"big128": dict(users=16384, items=262144, d=128, bias=True, nnz_train=16384 * 50, nnz_test=16384 * 10, top_k=[10, 20, 50, 100],
                         metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], seed=77, name="16K x 262K, d=128"),
          # one GPU's share of c4 (1M x 1M, d=128, bias, 50 train / 10 test items per user, K up to 100): 8,192 of its 125,000 users
          "c4slice": dict(users=8192, items=1_000_000, d=128, bias=True, nnz_train=8192 * 50, nnz_test=8192 * 10, top_k=[10, 20, 50, 100],
                          metric=["Precision", "Recall", "MAP", "NDCG", "MRR"], seed=2025, name="c4 slice: 8,192 users x 1M items, d=128")}
for name in (sys.argv[1:] or ["c2", "c3b", "big128"]):
    cfg = dict(SHAPES[name]) if name in SHAPES else dict(synth.CONFIGS[name])
    d = synth.make(device="cuda", **cfg)
    ctx = _native.Context(0)
    ctx.set_train_csr(d["train_indptr"], d["train_indices"], d["items"])
    ctx.set_test_csr(d["test_indptr"], d["test_indices"], d["items"])
    ue, ie = torch.from_numpy(d["user_emb"]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
    b = None if d["bias"] is None else torch.from_numpy(d["bias"]).cuda()
    ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
    K = max(cfg["top_k"])
    res = {}
    for prec in os.environ.get("SKR_PRECS", "3xtf32,tf32r,f16r").split(","):
        sums = torch.zeros(len(ids) * K, dtype=torch.float64, device="cuda")
        ts = []
        for it in range(6):
            sums.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.eval_fused(ue, ie, b, 0, ids, K, precision=prec, sums=sums)
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        res[prec] = (min(ts[2:]), ctx.fused_kernel_ms(0), ctx.fused_prepass_ms(0), ctx.fused_stats(), (sums / d["users"]).cpu().numpy())
        print("%-7s %-7s evaluate %.3f ms (main kernel %.3f, prepass %.3f) -> %.2f M users/s; %s" % (
            name, prec, res[prec][0], res[prec][1], res[prec][2], d["users"] / res[prec][0] / 1e3, res[prec][3]), flush=True)
    ks = list(res)
    print("%-7s max |mean metric diff| against %s: %s" % (name, ks[0], ", ".join("%s %.2e" % (k, float(np.max(np.abs(res[ks[0]][4] - res[k][4])))) for k in ks[1:])), flush=True)
    ctx.close()
