"""c2 with heavy-tailed item norms, a few fused evaluates (ncu target): python tools/heavy_c2.py [retry_min=-1] [n=3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

cfg = dict(synth.CONFIGS["c2"], norms="heavy")
d = synth.make(device="cuda", **cfg)
ctx = _native.Context(0)
ctx.set_train_csr(d["train_indptr"], d["train_indices"], d["items"])
ctx.set_test_csr(d["test_indptr"], d["test_indices"], d["items"])
ctx.set_option("retry_min", int(sys.argv[1]) if len(sys.argv) > 1 else -1)
ue, ie = torch.from_numpy(d["user_emb"]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
K = max(cfg["top_k"])
sums = torch.zeros(len(ids) * K, dtype=torch.float64, device="cuda")
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    sums.zero_()
    ctx.eval_fused(ue, ie, None, 0, ids, K, precision="tf32r", sums=sums)
    torch.cuda.synchronize()
print(ctx.fused_stats())
