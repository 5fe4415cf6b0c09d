"""Many rows on the fail list (weak thresholds): exercises the segmented exact fallback (development aid)."""
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
res = {}
seg = int(sys.argv[1]) if len(sys.argv) > 1 else -1
ctx.set_option("exact_seg_rows", seg)
for st, rk, prec in ((0, 0, "3xtf32"), (0, 16, "3xtf32"), (16, 0, "3xtf32"), (16, 0, "tf32r"), (4, 0, "3xtf32"), (4, 0, "tf32r"), (16, 0, "3xtf32")):
    ctx.set_option("sample_tiles", st)
    ctx.set_option("rank", rk)
    sums = torch.zeros(150, dtype=torch.float64, device="cuda")
    t = []
    for _ in range(3):
        sums.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.eval_fused(ue, ie, None, 0, [1, 2, 4], 50, precision=prec, sums=sums)
        e.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(e))
    res[(st, rk, prec)] = sums.cpu().numpy()
    print("sample_tiles=%d rank=%d %s: evaluate %.3f ms, %s" % (st, rk, prec, min(t), ctx.fused_stats()), flush=True)
base = res[(0, 0, "3xtf32")]
for k, v in res.items():
    print(k, "max |mean diff| vs default plan: %.2e" % (np.max(np.abs(v - base)) / d["users"]))
