"""Per-CTA overhead of the fused main pass: 148 x 128 users of c2's catalogue so that chunks = 1, 2, 3, ...
give exactly 1, 2, 3 full waves of equal CTAs (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from skrec_b200 import _native, synth  # noqa: E402

cfg = dict(synth.CONFIGS["c2"])
U = 148 * 128
cfg.update(users=U, nnz_train=int(810128 * U / 29858), nnz_test=int(217242 * U / 29858))
d = synth.make(device="cuda", **cfg)
ctx = _native.Context(0)
ctx.set_train_csr(d["train_indptr"], d["train_indices"], d["items"])
ctx.set_test_csr(d["test_indptr"], d["test_indices"], d["items"])
ue, ie = torch.from_numpy(d["user_emb"]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
sums = torch.zeros(150, dtype=torch.float64, device="cuda")
for prec in ("3xtf32", "1xtf32"):
    for chunks in (1, 2, 3, 4, 6, 8):
        ctx.set_option("chunks", chunks)
        ms = []
        for _ in range(4):
            ctx.eval_fused(ue, ie, None, 0, [1, 2, 4], 50, precision=prec, sums=sums)
            torch.cuda.synchronize()
            ms.append(ctx.fused_kernel_ms(0))
        plan = ctx.fused_stats()
        tiles = (321 + plan["chunks"] - 1) // plan["chunks"]
        print("%s chunks=%d (%d tiles/CTA, %d waves): collect %.3f ms = %.0f cycles/tile-slot; prepass %.3f ms" % (
            prec, plan["chunks"], tiles, plan["chunks"], min(ms), min(ms) * 1e-3 * 1.965e9 / (tiles * plan["chunks"]), ctx.fused_prepass_ms(0)), flush=True)
