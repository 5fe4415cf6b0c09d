"""Tile timeline of one CTA of the fused tcgen05 main pass (development aid).
    python tools/trace_tiles.py [precision] [cta] [config]
Prints per tile, relative to the tile's 'issuer: accumulator free' stamp of the first traced tile, the
clock64 stamps of every pipeline role (slots documented in k_fused_tc.cuh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "3xtf32"
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 300
cfgname = sys.argv[3] if len(sys.argv) > 3 else "c2"
d = synth.make_config(cfgname, device="cuda")
cfg = d["config"]
ctx = _native.Context(0)
ctx.set_train_csr(d["train_indptr"], d["train_indices"], d["items"])
ctx.set_test_csr(d["test_indptr"], d["test_indices"], d["items"])
ue, ie = torch.from_numpy(d["user_emb"]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
b = None if d["bias"] is None else torch.from_numpy(d["bias"]).cuda()
ids = [synth.METRIC_IDS[m] for m in cfg["metric"]]
K = max(cfg["top_k"])
sums = torch.zeros(len(ids) * K, dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.eval_fused(ue, ie, b, 0, ids, K, precision=prec, sums=sums)
torch.cuda.synchronize()
ctx.set_option("trace_cta", cta)
ctx.set_option("dbg", int(os.environ.get("SKR_DBG", "0")))  # ablation switches (results invalid), see TcArgs::dbg
ctx.eval_fused(ue, ie, b, 0, ids, K, precision=prec, sums=sums)
torch.cuda.synchronize()
plan = ctx.fused_stats()
n_ct = (d["items"] + 127) // 128
tiles = (n_ct + plan["chunks"] - 1) // plan["chunks"]
tr = ctx.fused_trace(tiles)
ctx.set_option("trace_cta", -1)
names = ["prodGot", "prodTMA", "issFree", "issFull", "issCommit", "mskFree", "mskRdy", "e0Full", "e0Rel", "e0Done", "e15Full", "e15Rel", "e15Done"]
t0 = tr[0, 2]
print("kernel %.3f ms, plan %s, tiles/CTA %d" % (ctx.fused_kernel_ms(0), plan, tiles))
print("tile " + " ".join("%9s" % n for n in names))
for i in list(range(0, 12)) + list(range(40, 52)):
    if i >= tiles:
        break
    print("%4d " % i + " ".join("%9d" % (tr[i, s] - t0 if tr[i, s] else -1) for s in range(13)))
steady = tr[10:tiles - 2]
for s, n in enumerate(names):
    dt = np.diff(steady[:, s])
    print("%-10s per-tile period: mean %.0f  min %d  max %d" % (n, dt.mean(), dt.min(), dt.max()))
print("mask builder: free->start %.0f | zero %.0f | keys %.0f | arrive %.0f" % ((steady[:, 13] - steady[:, 5]).mean(), (steady[:, 14] - steady[:, 13]).mean(),
      (steady[:, 15] - steady[:, 14]).mean(), (steady[:, 6] - steady[:, 15]).mean()))
print("issFree->issCommit %.0f | issCommit->e0Full %.0f | e0Full->e0Rel %.0f | e0Rel->e0Done %.0f | e15Full->e15Rel %.0f | e15Rel->issFree(+2) %.0f | mskFree->mskRdy %.0f | mskRdy->e0Full %.0f" % (
    (steady[:, 4] - steady[:, 2]).mean(), (steady[:, 7] - steady[:, 4]).mean(), (steady[:, 8] - steady[:, 7]).mean(),
    (steady[:, 9] - steady[:, 8]).mean(), (steady[:, 11] - steady[:, 10]).mean(),
    (steady[2:, 2] - np.maximum(steady[:-2, 8], steady[:-2, 11])).mean(), (steady[:, 6] - steady[:, 5]).mean(),
    (steady[:, 7] - steady[:, 6]).mean()))
