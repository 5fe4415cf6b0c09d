"""Small shapes with deliberately useless thresholds (rank 1): nearly every row goes through the exact fallback."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import oracle
from skrec_b200 import _native, synth
for U in (int(sys.argv[1]) if len(sys.argv) > 1 else 1500,):
    d = synth.make(users=U, items=int(sys.argv[2]) if len(sys.argv) > 2 else 5000, d=64, nnz_train=U * 20, nnz_test=U * 5, seed=3, device="cuda")
    ctx = _native.Context(0)
    ctx.set_train_csr(d["train_indptr"], d["train_indices"], d["items"])
    ctx.set_test_csr(d["test_indptr"], d["test_indices"], d["items"])
    ue, ie = torch.from_numpy(d["user_emb"]).cuda(), torch.from_numpy(d["item_emb"]).cuda()
    ctx.set_option("rank", 1)
    ctx.set_option("sample_tiles", 2)
    for prec in ("3xtf32", "tf32r"):
        sums = torch.zeros(100, dtype=torch.float64, device="cuda")
        idx = torch.empty((U, 20), dtype=torch.int32, device="cuda")
        ctx.eval_fused(ue, ie, None, 0, [1, 2, 3, 4, 5], 20, precision=prec, sums=sums, topk_idx=idx)
        torch.cuda.synchronize()
        S = oracle.scores(d["user_emb"], d["item_emb"], None)
        oracle.mask_rows(S, d["train_indptr"], d["train_indices"])
        eper, etop = oracle.eval_scores(S, d["test_indptr"], d["test_indices"], [1, 2, 3, 4, 5], 20, return_topk=True)
        print("U=%d %s: %s idx mismatch %.4f%% max mean diff %.2e" % (U, prec, ctx.fused_stats(), 100 * (idx.cpu().numpy() != etop).mean(),
              float(np.max(np.abs(sums.cpu().numpy() / U - oracle.sums_f64(eper) / U)))), flush=True)
