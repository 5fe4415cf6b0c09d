"""Timing experiments on the fused tcgen05 kernel with parts disabled (results are invalid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from skrec_b200 import _native, synth
d = synth.make_config("c2", device="cuda")
ctx = _native.Context(0)
ctx.set_train_csr(d["train_indptr"], d["train_indices"], d["items"])
ctx.set_test_csr(d["test_indptr"], d["test_indices"], d["items"])
ue, ie = torch.from_numpy(d["user_emb"]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
sums = torch.zeros(150, dtype=torch.float64, device="cuda")
for prec in ("1xtf32",):
    for stages in (0,):
        for dbg, name in ((0, "full"), (1, "no epilogue work"), (4, "ld+mask only, no appends"), (32, "detection, no slow path"), (64, "slow path, no global store"), (2, "no MMA"), (3, "no MMA, no epilogue"), (19, "no MMA, no epilogue, no bitmaps"), (16, "no bitmaps"), (9, "no TMA, no epilogue (MMA only)"), (25, "MMA only, no bitmaps")):
            ctx.set_option("dbg", dbg)
            ms = []
            for _ in range(3):
                ctx.eval_fused(ue, ie, None, 0, [1, 2, 4], 50, precision=prec, sums=sums)
                torch.cuda.synchronize()
                ms.append(ctx.fused_kernel_ms(0))
            print("%s stages=%d dbg=%3d %-32s collect %.3f ms  prepass %.3f ms" % (prec, stages, dbg, name, min(ms), ctx.fused_prepass_ms(0)), flush=True)
